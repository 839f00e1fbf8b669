// f3 — host-side serialisation of one frame's label JSON (§8a row R7).
//
// The reference writes label_%06d.json with json.dump(label, f, indent=2, ensure_ascii=False)
// (gcd.py:608-613; frame dict gcd.py:2056-2064, object dict gcd.py:1938-1946).  With indent set
// Python runs its pure-Python encoder: 23 ms for a 100-object frame (dict building included),
// 250x the cost of annotating the frame on the GPU including the PCIe copy.  This is the same text
// produced straight from the D2H record buffer: plain C++, no CUDA, all pointers are HOST pointers.
//
// Byte-identical to the Python path (formats.reference_label + json.dump): floats are written as
// Python's repr() — shortest round-trip digits (std::to_chars), fixed notation for
// 1e-4 <= |x| < 1e16 with a trailing ".0" on integers, d[.ddd]e±XX otherwise — non-finite label
// values become null exactly where formats.object_entry maps them to None, and NaN / Infinity where
// json.dump would let them through.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "cspe.h"
#include "repr6.h"

namespace cspe {
void set_error(const char* fmt, ...);
}

namespace {

// Bounds are checked per object (see `room`), not per character: one record needs at most
// kRecordBytes plus its two strings plus kJointBytes per joint.
constexpr int64_t kFrameBytes = 4096;   // frame-level keys, camera pose, mask shape
constexpr int64_t kRecordBytes = 4096;  // ~65 numbers of <= 40 bytes per line + ~20 keys
constexpr int64_t kJointBytes = 128;    // three lines per joint
const char kSpaces[] = "                                ";

struct Out {
  char* p;
  char* end;

  bool room(int64_t n) const { return end - p >= n; }
  void raw(const char* s, size_t n) {
    memcpy(p, s, n);
    p += n;
  }
  void lit(const char* s) { raw(s, strlen(s)); }
  void ch(char c) { *p++ = c; }
  void newline(int level) {
    *p++ = '\n';
    memcpy(p, kSpaces, 16);  // level <= 5: at most ten spaces are kept
    p += 2 * level;
  }
  void integer(long long v) {
    char b[24];
    auto r = std::to_chars(b, b + sizeof(b), v);
    raw(b, static_cast<size_t>(r.ptr - b));
  }
  // repr(round(float(v), 6)) of a float32 — the COCO ratio fields (repr6.h)
  void repr6(float v) {
    if (!(std::fabs(v) < 1048576.0f)) return repr(static_cast<double>(v));  // NaN / inf / >= 2^20: rounding changes nothing
    const unsigned long long q = static_cast<unsigned long long>(std::nearbyint(std::fabs(static_cast<double>(v)) * 1e6));
    p += cspe::repr_units6(std::signbit(v), q, p);
  }
  // Python float.__repr__
  void repr(double v) {
    if (std::isnan(v)) return lit("NaN");  // what json.dump(allow_nan=True) writes
    if (std::isinf(v)) return lit(v < 0 ? "-Infinity" : "Infinity");
    if (v == 0.0) return lit(std::signbit(v) ? "-0.0" : "0.0");
    char b[40];
    auto r = std::to_chars(b, b + sizeof(b), v, std::chars_format::scientific);  // [-]d[.ddd]e±XX, shortest digits
    const char* s = b;
    if (*s == '-') {
      ch('-');
      ++s;
    }
    const char* e = static_cast<const char*>(memchr(s, 'e', static_cast<size_t>(r.ptr - s)));
    char digits[24];
    int nd = 0;
    for (const char* q = s; q < e; ++q)
      if (*q != '.') digits[nd++] = *q;
    int exp10 = 0;
    std::from_chars(e + (e[1] == '+' ? 2 : 1), r.ptr, exp10);
    const int decpt = exp10 + 1;  // value = 0.d1d2... * 10^decpt
    if (decpt > 16 || decpt < -3) {  // float_repr_style 'short', repr: exponent form outside [1e-4, 1e16)
      ch(digits[0]);
      if (nd > 1) {
        ch('.');
        raw(digits + 1, static_cast<size_t>(nd - 1));
      }
      ch('e');
      ch(exp10 < 0 ? '-' : '+');
      const int a = exp10 < 0 ? -exp10 : exp10;
      if (a < 10) ch('0');
      integer(a);
    } else if (decpt <= 0) {
      raw("0.", 2);
      for (int i = 0; i < -decpt; ++i) ch('0');
      raw(digits, static_cast<size_t>(nd));
    } else if (decpt >= nd) {
      raw(digits, static_cast<size_t>(nd));
      for (int i = nd; i < decpt; ++i) ch('0');
      raw(".0", 2);
    } else {
      raw(digits, static_cast<size_t>(decpt));
      ch('.');
      raw(digits + decpt, static_cast<size_t>(nd - decpt));
    }
  }
  void finite_or_null(double v) {
    if (std::isfinite(v))
      repr(v);
    else
      lit("null");
  }
  void key(int level, const char* name, bool first = false) {
    if (!first) ch(',');
    newline(level);
    ch('"');
    lit(name);
    raw("\": ", 3);
  }
  // [v0, v1, ...] one element per line; non-finite -> null (formats._finite_list)
  void finite_list(int level, const double* v, int n) {
    ch('[');
    for (int i = 0; i < n; ++i) {
      if (i) ch(',');
      newline(level + 1);
      finite_or_null(v[i]);
    }
    newline(level);
    ch(']');
  }
  void int_list(int level, const int32_t* v, int n) {
    ch('[');
    for (int i = 0; i < n; ++i) {
      if (i) ch(',');
      newline(level + 1);
      integer(v[i]);
    }
    newline(level);
    ch(']');
  }
};

int64_t too_small(int64_t capacity) {
  cspe::set_error("cspe_format_label_json_host: output buffer of %lld bytes is too small (needs 4096 + fragments + "
                  "per record 4096 + strings + 128 per joint)",
                  static_cast<long long>(capacity));
  return CSPE_ERR_INVALID_ARGUMENT;
}

}  // namespace

extern "C" int64_t cspe_format_label_json_host(const cspe_record* records_host, int n, int64_t frame_id,
                                               const double* camera_pose7, const char* camera_params_json,
                                               const char* class_mapping_json, const char* slot_strings,
                                               const int32_t* slot_string_offsets, int num_slots, int height,
                                               int width, const double* keypoints, const uint8_t* visibility,
                                               const int32_t* person_of_slot, int num_people, int num_joints,
                                               char* out_host, int64_t capacity) {
  if (n < 0 || num_slots < 0 || capacity < 0 || !camera_pose7 || !camera_params_json || !class_mapping_json ||
      (n > 0 && (!records_host || !slot_strings || !slot_string_offsets)) || (capacity > 0 && !out_host) ||
      (person_of_slot && (!keypoints || !visibility || num_people < 0 || num_joints < 0))) {
    cspe::set_error("cspe_format_label_json_host: invalid argument");
    return CSPE_ERR_INVALID_ARGUMENT;
  }
  const int64_t fragments = static_cast<int64_t>(strlen(camera_params_json) + strlen(class_mapping_json));
  Out o{out_host, out_host + capacity};
  if (!o.room(kFrameBytes + fragments)) return too_small(capacity);
  o.ch('{');
  o.key(1, "frame_id", true);
  o.integer(frame_id);
  o.key(1, "camera_pose");  // [float(v) for v in camera_pose]: json.dump lets NaN / Infinity through here
  o.ch('[');
  for (int i = 0; i < 7; ++i) {
    if (i) o.ch(',');
    o.newline(2);
    o.repr(camera_pose7[i]);
  }
  o.newline(1);
  o.ch(']');
  o.key(1, "camera_params");
  o.lit(camera_params_json);
  o.key(1, "objects");
  if (n == 0) {
    o.lit("[]");
  } else {
    o.ch('[');
    for (int i = 0; i < n; ++i) {
      const cspe_record& r = records_host[i];
      if (r.inst_idx < 0 || r.inst_idx >= num_slots) {
        cspe::set_error("cspe_format_label_json_host: record %d has inst_idx %d outside [0, %d)", i, r.inst_idx,
                        num_slots);
        return CSPE_ERR_INVALID_ARGUMENT;
      }
      const bool valid = (r.flags & CSPE_OBJ_POSE_VALID) != 0;
      const int32_t* so = slot_string_offsets + 2 * r.inst_idx;
      if (so[2] < so[1] || so[1] < so[0]) {
        cspe::set_error("cspe_format_label_json_host: slot_string_offsets not ascending at slot %d", r.inst_idx);
        return CSPE_ERR_INVALID_ARGUMENT;
      }
      if (!o.room(kFrameBytes + fragments + kRecordBytes + (so[2] - so[0]) + kJointBytes * num_joints))
        return too_small(capacity);
      if (i) o.ch(',');
      o.newline(2);
      o.ch('{');
      o.key(3, "inst_idx", true);
      o.integer(r.inst_idx);
      o.key(3, "class_id");
      o.integer(r.class_id);
      o.key(3, "class_name");
      o.raw(slot_strings + so[0], static_cast<size_t>(so[1] - so[0]));
      o.key(3, "center");
      o.finite_list(3, r.pose + 7, 3);
      o.key(3, "size");
      o.finite_list(3, r.pose + 10, 3);
      o.key(3, "rotation");
      if (valid)
        o.finite_list(3, r.pose + 13, 3);
      else
        o.lit("null");
      o.key(3, "prim_path");
      o.raw(slot_strings + so[1], static_cast<size_t>(so[2] - so[1]));
      // ---- additions ([SPEC] stages) ----
      o.key(3, "pixel_count");
      o.integer(r.count);
      o.key(3, "bbox_2d_tight");
      const int32_t tight[4] = {r.x_min, r.y_min, r.x_max, r.y_max};
      o.int_list(3, tight, 4);
      o.key(3, "bbox_2d_loose");
      o.int_list(3, r.loose, 4);
      o.key(3, "occlusion");
      o.repr(static_cast<double>(r.occlusion));
      o.key(3, "truncation");
      o.repr(static_cast<double>(r.truncation));
      o.key(3, "fill");
      o.repr(static_cast<double>(r.fill));
      o.key(3, "bbox_3d_projected");
      o.ch('[');
      for (int k = 0; k < 8; ++k) {
        if (k) o.ch(',');
        o.newline(4);
        o.finite_list(4, r.uv + 2 * k, 2);
      }
      o.newline(3);
      o.ch(']');
      o.key(3, "bbox_3d_depth");
      o.finite_list(3, r.z, 8);
      o.key(3, "pose_in_camera");
      o.ch('{');
      o.key(4, "translation", true);
      o.finite_list(4, r.pose, 3);
      o.key(4, "quaternion_xyzw");
      if (valid)
        o.finite_list(4, r.pose + 3, 4);
      else
        o.lit("null");
      o.newline(3);
      o.ch('}');
      o.key(3, "flags");
      o.integer(r.flags);
      const int person = person_of_slot ? person_of_slot[r.inst_idx] : -1;
      if (person >= 0 && person < num_people) {  // formats.coco_keypoint_block
        const double* kp = keypoints + static_cast<int64_t>(person) * num_joints * 2;
        const uint8_t* vis = visibility + static_cast<int64_t>(person) * num_joints;
        o.key(3, "keypoints");
        o.ch('{');
        o.key(4, "keypoints", true);
        int seen = 0;
        if (num_joints == 0) {
          o.lit("[]");
        } else {
          o.ch('[');
          for (int j = 0; j < num_joints; ++j) {
            const bool shown = vis[j] != 0 && std::isfinite(kp[2 * j]) && std::isfinite(kp[2 * j + 1]);
            seen += vis[j] > 0;
            if (j) o.ch(',');
            o.newline(5);
            o.repr(shown ? kp[2 * j] : 0.0);
            o.ch(',');
            o.newline(5);
            o.repr(shown ? kp[2 * j + 1] : 0.0);
            o.ch(',');
            o.newline(5);
            o.integer(shown ? vis[j] : 0);
          }
          o.newline(4);
          o.ch(']');
        }
        o.key(4, "num_keypoints");
        o.integer(seen);
        o.newline(3);
        o.ch('}');
      }
      o.newline(2);
      o.ch('}');
    }
    o.newline(1);
    o.ch(']');
  }
  o.key(1, "instance_mask_shape");
  const int32_t shape[2] = {height, width};
  o.int_list(1, shape, 2);
  o.key(1, "num_objects");
  o.integer(n);
  o.key(1, "class_mapping");
  o.lit(class_mapping_json);
  o.newline(0);
  o.ch('}');
  return o.p - out_host;
}

// COCO annotations of a batch: for every kept record of frames [0, frames) one
//   {"id": i, "image_id": f, "category_id": c, "bbox": [x, y, w, h], "area": a, "iscrowd": 0,
//    "occlusion": o, "truncation": t[, "keypoints": [x, y, v, ...], "num_keypoints": k]}
// joined by ", " — the text json.dump(list_of_annotations) puts between its brackets (default
// separators), byte for byte what formats.coco_annotations + json.dump produce.
extern "C" int64_t cspe_format_coco_host(const cspe_record* records_host, const int32_t* n_out_host, int B, int N,
                                         int frames, const int64_t* image_ids, int64_t first_annotation_id,
                                         const double* keypoints, const uint8_t* visibility,
                                         const int32_t* person_of_slot, int num_people, int num_joints,
                                         char* out_host, int64_t capacity) {
  if (B < 0 || N < 0 || frames < 0 || frames > B || capacity < 0 || (frames > 0 && (!records_host || !n_out_host || !image_ids)) ||
      (capacity > 0 && !out_host) || (person_of_slot && (!keypoints || !visibility || num_people < 0 || num_joints < 0))) {
    cspe::set_error("cspe_format_coco_host: invalid argument");
    return CSPE_ERR_INVALID_ARGUMENT;
  }
  Out o{out_host, out_host + capacity};
  int64_t ann_id = first_annotation_id;
  bool first = true;
  for (int f = 0; f < frames; ++f) {
    const int n = n_out_host[f];
    if (n < 0 || n > N) {
      cspe::set_error("cspe_format_coco_host: n_out[%d] = %d outside [0, %d]", f, n, N);
      return CSPE_ERR_INVALID_ARGUMENT;
    }
    const cspe_record* r = records_host + static_cast<int64_t>(f) * N;
    for (int i = 0; i < n; ++i, ++r, ++ann_id) {
      if (!o.room(512 + kJointBytes * num_joints)) {
        cspe::set_error("cspe_format_coco_host: output buffer of %lld bytes is too small (512 + 128 per joint per record)",
                        static_cast<long long>(capacity));
        return CSPE_ERR_INVALID_ARGUMENT;
      }
      if (!first) o.raw(", ", 2);
      first = false;
      o.lit("{\"id\": ");
      o.integer(ann_id);
      o.lit(", \"image_id\": ");
      o.integer(image_ids[f]);
      o.lit(", \"category_id\": ");
      o.integer(r->class_id);
      o.lit(", \"bbox\": [");
      if (r->count > 0) {
        o.integer(r->x_min);
        o.raw(", ", 2);
        o.integer(r->y_min);
        o.raw(", ", 2);
        o.integer(static_cast<long long>(r->x_max) - r->x_min + 1);
        o.raw(", ", 2);
        o.integer(static_cast<long long>(r->y_max) - r->y_min + 1);
      } else {
        o.lit("0, 0, 0, 0");
      }
      o.lit("], \"area\": ");
      o.integer(r->count);
      o.lit(", \"iscrowd\": 0, \"occlusion\": ");
      o.repr6(r->occlusion);
      o.lit(", \"truncation\": ");
      o.repr6(r->truncation);
      int person = -1;
      if (person_of_slot && r->inst_idx >= 0 && r->inst_idx < N)
        person = person_of_slot[static_cast<int64_t>(f) * N + r->inst_idx];
      if (person >= 0 && person < num_people) {
        const double* kp = keypoints + (static_cast<int64_t>(f) * num_people + person) * num_joints * 2;
        const uint8_t* vis = visibility + (static_cast<int64_t>(f) * num_people + person) * num_joints;
        o.lit(", \"keypoints\": [");
        int seen = 0;
        for (int j = 0; j < num_joints; ++j) {
          const bool shown = vis[j] != 0 && std::isfinite(kp[2 * j]) && std::isfinite(kp[2 * j + 1]);
          seen += vis[j] > 0;
          if (j) o.raw(", ", 2);
          o.repr(shown ? kp[2 * j] : 0.0);
          o.raw(", ", 2);
          o.repr(shown ? kp[2 * j + 1] : 0.0);
          o.raw(", ", 2);
          o.integer(shown ? vis[j] : 0);
        }
        o.lit("], \"num_keypoints\": ");
        o.integer(seen);
      }
      o.ch('}');
    }
  }
  return o.p - out_host;
}
