// S6 / f3 — YOLO label text on the device (SURVEY §8a row S6 "YOLO line `class cx cy w h` normalised",
// §8f row f3 "host serialisation is the end-to-end limiter at 100 k frames").
//
// A 100 k-frame sweep moves 26 KB of records per frame to the host only to print 38 bytes per object
// from them.  This kernel prints the lines where the records are: one CTA per frame turns the frame's
// kept records (K4's output) into
//     f"{class_id} {cx:.6f} {cy:.6f} {w:.6f} {h:.6f}\n"
// so that D2H carries ~2 KB of text per frame instead, and the host only writes files.  The bytes are
// identical to the Python formatter (formats.yolo_lines) and to cspe_format_yolo_host: for a float32 v
// with |v| < 2^20 the product v * 1e6 is exact in double (24-bit significand times 2^6 * 15625), so
// rounding it to nearest-even IS the correctly rounded 6-decimal value printf / Python print.
//
// Each frame's text is contiguous at text + f * frame_stride; n_bytes[f] is its full size (bytes past
// frame_stride are dropped but counted; -1 = a box value outside |v| < 2^20 or not finite).
#include <math.h>

#include "cspe_common.cuh"
#include "repr6.h"

namespace cspe {
namespace {

constexpr int kYoloThreads = 256;

__device__ __forceinline__ int dec_digits(unsigned long long v) {
  int n = 1;
  while (v >= 10ull) {
    v /= 10ull;
    ++n;
  }
  return n;
}

// q = round_half_even(|v| * 1e6); false when v is outside the exact domain
__device__ __forceinline__ bool fixed6_units(float v, unsigned long long* q) {
  if (!(fabsf(v) < 1048576.0f)) return false;
  *q = static_cast<unsigned long long>(__double2ll_rn(fabs(static_cast<double>(v)) * 1e6));
  return true;
}

__device__ __forceinline__ int fixed6_len(float v, unsigned long long q) {
  return (signbit(v) ? 1 : 0) + dec_digits(q / 1000000ull) + 7;
}

// bounded byte sink: counts everything, stores what fits
struct Sink {
  char* p;
  long long pos, cap;
  __device__ __forceinline__ void put(char c) {
    if (pos < cap) p[pos] = c;
    ++pos;
  }
  __device__ __forceinline__ void put_uint(unsigned long long v) {
    char tmp[20];
    int n = 0;
    do {
      tmp[n++] = static_cast<char>('0' + static_cast<int>(v % 10ull));
      v /= 10ull;
    } while (v);
    while (n) put(tmp[--n]);
  }
};

__global__ void __launch_bounds__(kYoloThreads)
    yolo_text_kernel(const cspe_record* __restrict__ records, const int32_t* __restrict__ n_out, int N, char* text,
                     long long frame_stride, int32_t* __restrict__ n_bytes) {
  __shared__ int warp_sums[kYoloThreads / 32];
  __shared__ long long base_s;
  __shared__ int bad_s;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  pdl_launch_dependents();  // whatever is queued behind may get placed; it waits for us before it reads
  if (tid == 0) {
    base_s = 0;
    bad_s = 0;
  }
  __syncthreads();
  pdl_wait();  // K4's records and n_out are complete after this
  // ... but they were written while this kernel was already resident: explicit L2 loads, never ld.global.nc
  // (cspe_common.cuh, PDL rule)
  int n = __ldcg(n_out + f);
  n = n < 0 ? 0 : (n > N ? N : n);
  const cspe_record* rec = records + static_cast<long long>(f) * N;
  char* out = text + static_cast<long long>(f) * frame_stride;
  for (int r0 = 0; r0 < n; r0 += kYoloThreads) {
    const int r = r0 + tid;
    int len = 0;
    int cls = 0;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned long long q[4] = {0, 0, 0, 0};
    if (r < n) {
      cls = __ldcg(&rec[r].class_id);
      len = (cls < 0 ? 1 : 0) + dec_digits(static_cast<unsigned long long>(cls < 0 ? -static_cast<long long>(cls) : cls)) + 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = __ldcg(&rec[r].yolo[k]);
        if (!fixed6_units(v[k], &q[k])) bad_s = 1;
        len += 1 + fixed6_len(v[k], q[k]);
      }
    }
    // exclusive prefix of the line lengths inside the chunk
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kYoloThreads / 32; ++w) {
      const int s = warp_sums[w];
      if (w < wid) before += s;
      total += s;
    }
    if (r < n) {
      Sink sk{out, base_s + before + incl - len, frame_stride};
      if (cls < 0) sk.put('-');
      sk.put_uint(static_cast<unsigned long long>(cls < 0 ? -static_cast<long long>(cls) : cls));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sk.put(' ');
        if (signbit(v[k])) sk.put('-');
        sk.put_uint(q[k] / 1000000ull);
        sk.put('.');
        unsigned int frac = static_cast<unsigned int>(q[k] % 1000000ull);
        char d[6];
#pragma unroll
        for (int i = 5; i >= 0; --i) {
          d[i] = static_cast<char>('0' + frac % 10u);
          frac /= 10u;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) sk.put(d[i]);
      }
      sk.put('\n');
    }
    __syncthreads();
    if (tid == 0) base_s += total;
    __syncthreads();
  }
  if (tid == 0) n_bytes[f] = bad_s ? -1 : static_cast<int32_t>(base_s);
}

// ---- COCO annotations on the device -----------------------------------------------------------------------------
// One CTA per frame prints, for every kept record,
//   {"id": i, "image_id": frame, "category_id": c, "bbox": [x, y, w, h], "area": a, "iscrowd": 0,
//    "occlusion": o, "truncation": t}
// each preceded by ", " unless it is annotation 1 — so the frames' texts, concatenated in order, are byte for byte
// what cspe_format_coco_host / json.dump(formats.coco_annotations(...)) write for records without keypoints.  Annotation ids count up across frames and batches: ann_state[0] holds the number of
// annotations printed so far (the caller zeroes it at sweep start), every CTA adds the counts of the frames before it
// in the batch, and the LAST CTA to finish (ticket in ann_state[1]) adds the batch total.
constexpr int kCocoMaxRecord = 224;   // longest record: 10-digit ids / coordinates, two 13-character ratios

// Two writers with one interface: the first pass only counts, the second stores at the final position (bounded by the
// frame's stride), so a record is never staged in local memory.
struct CountWriter {
  int n = 0;
  __device__ __forceinline__ void put(char) { ++n; }
};
struct TextWriter {
  char* p;
  long long pos, cap;
  __device__ __forceinline__ void put(char c) {
    if (pos < cap) p[pos] = c;
    ++pos;
  }
};

template <class W>
__device__ __forceinline__ void w_lit(W& w, const char* s) {
  for (int i = 0; s[i]; ++i) w.put(s[i]);
}

template <class W>
__device__ __forceinline__ void w_int(W& w, long long v) {
  if (v < 0) {
    w.put('-');
    v = -v;
  }
  char tmp[20];
  const int n = put_decimal(static_cast<unsigned long long>(v), tmp);
  for (int i = 0; i < n; ++i) w.put(tmp[i]);
}

template <class W>
__device__ __forceinline__ void w_ratio(W& w, float v) {
  char tmp[28];
  const int n = repr_units6(signbit(v), static_cast<unsigned long long>(__double2ll_rn(fabs(static_cast<double>(v)) * 1e6)), tmp);
  for (int i = 0; i < n; ++i) w.put(tmp[i]);
}

struct CocoFields {
  long long ann_id;
  int frame, cls, cnt, x0, y0, x1, y1;
  float occ, trunc;
  bool sep;   // ", " in front (every annotation but the sweep's first)
};

template <class W>
__device__ __forceinline__ void coco_record(W& w, const CocoFields& r) {
  if (r.sep) w_lit(w, ", ");
  w_lit(w, "{\"id\": ");
  w_int(w, r.ann_id);
  w_lit(w, ", \"image_id\": ");
  w_int(w, r.frame);
  w_lit(w, ", \"category_id\": ");
  w_int(w, r.cls);
  w_lit(w, ", \"bbox\": [");
  if (r.cnt > 0) {
    w_int(w, r.x0);
    w_lit(w, ", ");
    w_int(w, r.y0);
    w_lit(w, ", ");
    w_int(w, static_cast<long long>(r.x1) - r.x0 + 1);
    w_lit(w, ", ");
    w_int(w, static_cast<long long>(r.y1) - r.y0 + 1);
  } else {
    w_lit(w, "0, 0, 0, 0");
  }
  w_lit(w, "], \"area\": ");
  w_int(w, r.cnt);
  w_lit(w, ", \"iscrowd\": 0, \"occlusion\": ");
  w_ratio(w, r.occ);
  w_lit(w, ", \"truncation\": ");
  w_ratio(w, r.trunc);
  w.put('}');
}

__global__ void __launch_bounds__(kYoloThreads)
    coco_text_kernel(const cspe_record* records, const int32_t* n_out, int B, int N, unsigned long long* ann_state,
                     char* text, long long frame_stride, int32_t* n_bytes) {
  __shared__ int warp_sums[kYoloThreads / 32];
  __shared__ long long base_s, first_id_s;
  __shared__ long long part_s[kYoloThreads / 32];
  __shared__ int bad_s;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  pdl_launch_dependents();
  if (tid == 0) {
    base_s = 0;
    bad_s = 0;
  }
  __syncthreads();
  pdl_wait();  // K4's records and n_out are complete after this; read through L2 (cspe_common.cuh, PDL rule)
  int n = __ldcg(n_out + f);
  n = n < 0 ? 0 : (n > N ? N : n);
  // annotation id of this frame's first record: everything printed before this batch + the frames before this one
  long long before_frames = 0;
  for (int g = tid; g < f; g += kYoloThreads) {
    const int m = __ldcg(n_out + g);
    before_frames += m < 0 ? 0 : (m > N ? N : m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before_frames += __shfl_xor_sync(0xffffffffu, before_frames, o);
  if (lane == 0) part_s[wid] = before_frames;
  __syncthreads();
  if (tid == 0) {
    long long s = 0;
#pragma unroll
    for (int w = 0; w < kYoloThreads / 32; ++w) s += part_s[w];
    first_id_s = static_cast<long long>(__ldcg(ann_state)) + s + 1;
  }
  __syncthreads();
  const cspe_record* rec = records + static_cast<long long>(f) * N;
  char* out = text + static_cast<long long>(f) * frame_stride;
  for (int r0 = 0; r0 < n; r0 += kYoloThreads) {
    const int r = r0 + tid;
    CocoFields fld{};
    int len = 0;
    if (r < n) {
      const cspe_record* q = rec + r;
      fld.ann_id = first_id_s + r;
      fld.frame = __ldcg(&q->frame);
      fld.cls = __ldcg(&q->class_id);
      fld.cnt = __ldcg(&q->count);
      fld.x0 = __ldcg(&q->x_min);
      fld.y0 = __ldcg(&q->y_min);
      fld.x1 = __ldcg(&q->x_max);
      fld.y1 = __ldcg(&q->y_max);
      fld.occ = __ldcg(&q->occlusion);
      fld.trunc = __ldcg(&q->truncation);
      fld.sep = fld.ann_id > 1;
      if (!(fabsf(fld.occ) < 1048576.0f) || !(fabsf(fld.trunc) < 1048576.0f)) {
        bad_s = 1;   // an unprintable ratio: the frame is flagged, nothing is written for the record
      } else {
        CountWriter cw;
        coco_record(cw, fld);
        len = cw.n;
      }
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kYoloThreads / 32; ++w) {
      const int sw = warp_sums[w];
      if (w < wid) before += sw;
      total += sw;
    }
    if (len > 0) {
      TextWriter tw{out, base_s + before + incl - len, frame_stride};
      coco_record(tw, fld);
    }
    __syncthreads();
    if (tid == 0) base_s += total;
    __syncthreads();
  }
  if (tid == 0) {
    n_bytes[f] = bad_s ? -1 : static_cast<int32_t>(base_s);
    // last CTA of the batch: every CTA has read ann_state[0] by now -> add the batch total, reset the ticket
    __threadfence();
    if (atomicAdd(ann_state + 1, 1ull) == static_cast<unsigned long long>(B) - 1) {
      long long total = 0;
      for (int g = 0; g < B; ++g) {
        const int m = __ldcg(n_out + g);
        total += m < 0 ? 0 : (m > N ? N : m);
      }
      ann_state[0] = __ldcg(ann_state) + static_cast<unsigned long long>(total);
      ann_state[1] = 0ull;
    }
  }
}

// ---- frames' text back to back ------------------------------------------------------------------------------------
// The formatters above leave every frame's text at a fixed stride (no cross-frame dependency while printing).  A
// sweep that keeps the text of a whole batch wants ONE chunk: CTA f copies row f behind the rows before it; the host
// then takes packed[0 .. total) with a single memcpy instead of B slices.
__global__ void __launch_bounds__(kYoloThreads)
    pack_rows_kernel(const char* text, long long stride, const int32_t* n_bytes, int B, char* packed, long long capacity,
                     long long* total) {
  __shared__ long long part_s[kYoloThreads / 32];
  __shared__ long long off_s;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  pdl_launch_dependents();
  pdl_wait();   // the formatter's text and sizes, read through L2 (cspe_common.cuh, PDL rule)
  long long before = 0;
  for (int g = tid; g < f; g += kYoloThreads) {
    const long long m = __ldcg(n_bytes + g);
    before += m < 0 ? 0 : (m > stride ? stride : m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0) part_s[wid] = before;
  __syncthreads();
  if (tid == 0) {
    long long s = 0;
#pragma unroll
    for (int w = 0; w < kYoloThreads / 32; ++w) s += part_s[w];
    off_s = s;
  }
  __syncthreads();
  long long n = __ldcg(n_bytes + f);
  n = n < 0 ? 0 : (n > stride ? stride : n);
  const long long off = off_s;
  const char* src = text + static_cast<long long>(f) * stride;
  if (((reinterpret_cast<uintptr_t>(src)) & 15) == 0) {   // rows at a 16-byte stride: one 16-byte load per 16 bytes
    for (long long i = static_cast<long long>(tid) * 16; i < n; i += kYoloThreads * 16) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src + i));
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (i + k < n && off + i + k < capacity) packed[off + i + k] = static_cast<char>((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
    }
  } else {
    for (long long i = tid; i < n; i += kYoloThreads)
      if (off + i < capacity) packed[off + i] = __ldcg(src + i);
  }
  if (f == B - 1 && tid == 0) *total = off + n;
}

}  // namespace
}  // namespace cspe

using namespace cspe;

extern "C" int cspe_format_yolo(const cspe_record* records, const int32_t* n_out, int B, int N, char* text,
                                int64_t frame_stride, int32_t* n_bytes, void* stream) {
  CSPE_REQUIRE(B >= 0 && N >= 0 && frame_stride >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_yolo: negative size (B=%d N=%d frame_stride=%lld)", B, N, static_cast<long long>(frame_stride));
  if (B == 0) return CSPE_OK;
  CSPE_REQUIRE(n_out && n_bytes && (N == 0 || records) && (frame_stride == 0 || text), CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_yolo: null pointer");
  CSPE_REQUIRE(static_cast<long long>(N) * 96 < (1ll << 31), CSPE_ERR_UNSUPPORTED,
               "cspe_format_yolo: %d slots per frame overflow the int32 byte count", N);
  CSPE_CUDA_OK(launch_pdl(yolo_text_kernel, dim3(static_cast<unsigned>(B)), dim3(kYoloThreads), 0,
                          static_cast<cudaStream_t>(stream), records, n_out, N, text,
                          static_cast<long long>(frame_stride), n_bytes));
  return CSPE_OK;
}

extern "C" int cspe_format_coco(const cspe_record* records, const int32_t* n_out, int B, int N, int64_t* ann_state,
                                char* text, int64_t frame_stride, int32_t* n_bytes, void* stream) {
  CSPE_REQUIRE(B >= 0 && N >= 0 && frame_stride >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_coco: negative size (B=%d N=%d frame_stride=%lld)", B, N, static_cast<long long>(frame_stride));
  if (B == 0) return CSPE_OK;
  CSPE_REQUIRE(n_out && n_bytes && ann_state && (N == 0 || records) && (frame_stride == 0 || text), CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_coco: null pointer");
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(ann_state) & 7) == 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_coco: ann_state must be 8-byte aligned");
  CSPE_REQUIRE(static_cast<long long>(N) * (kCocoMaxRecord + 2) < (1ll << 31), CSPE_ERR_UNSUPPORTED,
               "cspe_format_coco: %d slots per frame overflow the int32 byte count", N);
  CSPE_CUDA_OK(launch_pdl(coco_text_kernel, dim3(static_cast<unsigned>(B)), dim3(kYoloThreads), 0,
                          static_cast<cudaStream_t>(stream), records, n_out, B, N,
                          reinterpret_cast<unsigned long long*>(ann_state), text, static_cast<long long>(frame_stride),
                          n_bytes));
  return CSPE_OK;
}

extern "C" int cspe_pack_rows(const char* text, int64_t frame_stride, const int32_t* n_bytes, int B, char* packed,
                              int64_t capacity, int64_t* total_bytes, void* stream) {
  CSPE_REQUIRE(B >= 0 && frame_stride >= 0 && capacity >= 0, CSPE_ERR_INVALID_ARGUMENT, "cspe_pack_rows: negative size");
  if (B == 0) return CSPE_OK;
  CSPE_REQUIRE(n_bytes && total_bytes && (frame_stride == 0 || text) && (capacity == 0 || packed), CSPE_ERR_INVALID_ARGUMENT,
               "cspe_pack_rows: null pointer");
  CSPE_CUDA_OK(launch_pdl(pack_rows_kernel, dim3(static_cast<unsigned>(B)), dim3(kYoloThreads), 0,
                          static_cast<cudaStream_t>(stream), text, static_cast<long long>(frame_stride), n_bytes, B, packed,
                          static_cast<long long>(capacity), reinterpret_cast<long long*>(total_bytes)));
  return CSPE_OK;
}
