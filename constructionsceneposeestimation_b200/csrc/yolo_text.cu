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
#include <stdlib.h>

#include "cspe_common.cuh"
#include "repr6.h"

namespace cspe {
namespace {

constexpr int kYoloThreads = 256;

// The first version printed every record byte by byte into global memory through local-memory digit buffers with 64-bit
// divisions and took 32 us per 64-frame batch (ncu) — a third of the mask scan, and its resident CTAs held registers the
// next scan's CTAs were waiting for.  Now: lengths from compare ladders (no division), digits written right to left
// straight at their final place in a SHARED-memory stage (32-bit arithmetic; ids beyond 2^32 take a 64-bit loop), and
// the stage leaves as aligned 16-byte stores (it is laid out at the same offset modulo 16 as its destination).
__device__ __forceinline__ int digits_u32(unsigned v) {
  return 1 + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) +
         (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
}
__device__ __forceinline__ int digits_u64(unsigned long long v) {
  if (v <= 0xffffffffull) return digits_u32(static_cast<unsigned>(v));
  int n = 10;
  v /= 10000000000ull;
  while (v) {
    v /= 10ull;
    ++n;
  }
  return n;
}
__device__ __forceinline__ int len_i32(int v) { return v < 0 ? 1 + digits_u32(0u - static_cast<unsigned>(v)) : digits_u32(static_cast<unsigned>(v)); }
__device__ __forceinline__ int len_i64(long long v) {
  return v < 0 ? 1 + digits_u64(0ull - static_cast<unsigned long long>(v)) : digits_u64(static_cast<unsigned long long>(v));
}

// length of repr(round(x, 6)) for x = +-q * 1e-6 (repr6.h: repr_units6), without printing it
__device__ __forceinline__ int len_units6(bool neg, unsigned long long q) {
  const int s = neg ? 1 : 0;
  if (q == 0) return s + 3;                                    // 0.0
  if (q < 100ull) {                                            // exponent form
    const unsigned tens = static_cast<unsigned>(q) / 10u, ones = static_cast<unsigned>(q) % 10u;
    return s + (tens == 0 ? 5 : (ones ? 7 : 5));               // 5e-06 | 1.2e-05 | 1e-05
  }
  unsigned frac = static_cast<unsigned>(q % 1000000ull);
  int fd = 6;                                                  // fractional digits kept: trailing zeros dropped, >= 1
  while (fd > 1 && frac % 10u == 0u) {
    frac /= 10u;
    --fd;
  }
  return s + digits_u64(q / 1000000ull) + 1 + fd;
}

// cursor into the shared-memory stage
struct Stage {
  char* p;
  __device__ __forceinline__ void put(char c) { *p++ = c; }
  template <int kLen>
  __device__ __forceinline__ void lit(const char (&s)[kLen]) {   // kLen counts the terminating NUL
#pragma unroll
    for (int i = 0; i < kLen - 1; ++i) p[i] = s[i];
    p += kLen - 1;
  }
  __device__ __forceinline__ void u32(unsigned v) {
    const int n = digits_u32(v);
    for (int i = n - 1; i >= 0; --i) {
      const unsigned q = v / 10u;
      p[i] = static_cast<char>('0' + (v - q * 10u));
      v = q;
    }
    p += n;
  }
  __device__ __forceinline__ void u64(unsigned long long v) {
    if (v <= 0xffffffffull) {
      u32(static_cast<unsigned>(v));
      return;
    }
    const int n = digits_u64(v);
    for (int i = n - 1; i >= 0; --i) {
      const unsigned long long q = v / 10ull;
      p[i] = static_cast<char>('0' + static_cast<int>(v - q * 10ull));
      v = q;
    }
    p += n;
  }
  __device__ __forceinline__ void i32(int v) {
    if (v < 0) {
      put('-');
      u32(0u - static_cast<unsigned>(v));
    } else {
      u32(static_cast<unsigned>(v));
    }
  }
  __device__ __forceinline__ void i64(long long v) {
    if (v < 0) {
      put('-');
      u64(0ull - static_cast<unsigned long long>(v));
    } else {
      u64(static_cast<unsigned long long>(v));
    }
  }
  // repr(round(x, 6)), the bytes of repr_units6 (repr6.h)
  __device__ __forceinline__ void units6(bool neg, unsigned long long q) {
    if (neg) put('-');
    if (q == 0) {
      lit("0.0");
      return;
    }
    if (q < 100ull) {
      const unsigned tens = static_cast<unsigned>(q) / 10u, ones = static_cast<unsigned>(q) % 10u;
      if (tens == 0) {
        put(static_cast<char>('0' + ones));
        lit("e-06");
      } else {
        put(static_cast<char>('0' + tens));
        if (ones) {
          put('.');
          put(static_cast<char>('0' + ones));
        }
        lit("e-05");
      }
      return;
    }
    u64(q / 1000000ull);
    put('.');
    unsigned frac = static_cast<unsigned>(q % 1000000ull);
    int fd = 6;
    while (fd > 1 && frac % 10u == 0u) {
      frac /= 10u;
      --fd;
    }
    for (int i = fd - 1; i >= 0; --i) {   // fd digits, leading zeros included
      const unsigned d = frac / 10u;
      p[i] = static_cast<char>('0' + (frac - d * 10u));
      frac = d;
    }
    p += fd;
  }
};

// bytes [0, n) of a shared-memory stage -> dst, where stage[0] corresponds to dst rounded DOWN to 16 bytes (the text
// starts at stage + (dst & 15)): whole 16-byte chunks leave as one aligned store, the ragged ends byte by byte
template <int kThreads>
__device__ __forceinline__ void stage_to_global(const char* stage, char* dst, long long n, int tid) {
  const int a = static_cast<int>(reinterpret_cast<uintptr_t>(dst) & 15);
  char* g0 = dst - a;
  const long long end = a + n;   // stage / g0 relative
  for (long long c = static_cast<long long>(tid) * 16; c < end; c += kThreads * 16) {
    if (c >= a && c + 16 <= end) {
      *reinterpret_cast<uint4*>(g0 + c) = *reinterpret_cast<const uint4*>(stage + c);
    } else {
      const long long lo = c > a ? c : a, hi = c + 16 < end ? c + 16 : end;
      for (long long i = lo; i < hi; ++i) g0[i] = stage[i];
    }
  }
}

// q = round_half_even(|v| * 1e6); false when v is outside the exact domain
__device__ __forceinline__ bool fixed6_units(float v, unsigned long long* q) {
  if (!(fabsf(v) < 1048576.0f)) return false;
  *q = static_cast<unsigned long long>(__double2ll_rn(fabs(static_cast<double>(v)) * 1e6));
  return true;
}

// "%.6f" of +-q * 1e-6 with |value| < 2^20: q < 2^40, integer part < 2^20, fraction < 10^6 — all 32-bit after the split
__device__ __forceinline__ int fixed6_len(float v, unsigned long long q) {
  return (signbit(v) ? 1 : 0) + digits_u32(static_cast<unsigned>(q / 1000000ull)) + 7;
}

constexpr int kYoloMaxLine = 96;   // 11 (class) + 4 x (1 + 1 + 7 + 1 + 6) + 1 = 76 at most
constexpr int kYoloStage = kYoloThreads * kYoloMaxLine + 16;

__global__ void __launch_bounds__(kYoloThreads)
    yolo_text_kernel(const cspe_record* __restrict__ records, const int32_t* __restrict__ n_out, int N, char* text,
                     long long frame_stride, int32_t* __restrict__ n_bytes) {
  __shared__ int warp_sums[kYoloThreads / 32];
  __shared__ long long base_s;
  __shared__ int bad_s;
  __shared__ __align__(16) char stage_s[kYoloStage];
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
      len = len_i32(cls) + 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = __ldcg(&rec[r].yolo[k]);
        if (!fixed6_units(v[k], &q[k])) bad_s = 1;   // q stays 0: the line is printed with 0.000000, the frame flagged
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
    // the pass is staged in shared memory at the same offset modulo 16 as its destination, then leaves as aligned
    // 16-byte stores (the first version stored it byte by byte from every thread)
    const long long base = base_s;
    char* dst = out + base;
    const int a = static_cast<int>(reinterpret_cast<uintptr_t>(dst) & 15);
    if (r < n) {
      Stage st{stage_s + a + before + incl - len};
      st.i32(cls);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        st.put(' ');
        if (signbit(v[k])) st.put('-');
        st.u32(static_cast<unsigned>(q[k] / 1000000ull));
        st.put('.');
        unsigned frac = static_cast<unsigned>(q[k] % 1000000ull);
#pragma unroll
        for (int i = 5; i >= 0; --i) {
          const unsigned d = frac / 10u;
          st.p[i] = static_cast<char>('0' + (frac - d * 10u));
          frac = d;
        }
        st.p += 6;
      }
      st.put('\n');
    }
    __syncthreads();
    long long room = frame_stride - base;   // bytes past the frame's stride are dropped but counted
    room = room < 0 ? 0 : (room > total ? total : room);
    stage_to_global<kYoloThreads>(stage_s, dst, room, tid);
    __syncthreads();
    if (tid == 0) base_s = base + total;
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
constexpr int kCocoThreads = 128;     // records printed per pass; a frame of the configs here holds 50-100
// shared-memory stage of one pass: 240 bytes is the true bound of a record (19-digit id, seven signed 10-digit
// integers, two 15-character ratios) — kCocoMaxRecord is what realistic values need and what callers size strides by
constexpr int kCocoStage = kCocoThreads * 240 + 16;

struct CocoFields {
  long long ann_id;
  int frame, cls, cnt, x0, y0, x1, y1;
  unsigned long long q_occ, q_trunc;   // round-half-even(|ratio| * 1e6)
  bool neg_occ, neg_trunc;
  bool sep;   // ", " in front (every annotation but the sweep's first)
};

// the literal text of a record: ({"id": )(, "image_id": )(, "category_id": )(, "bbox": [)(], "area": )
// (, "iscrowd": 0, "occlusion": )(, "truncation": )(}) = 7 + 14 + 17 + 11 + 11 + 29 + 16 + 1 = 106 bytes
constexpr int kCocoLiteral = 106;

__device__ __forceinline__ int coco_len(const CocoFields& r) {
  int n = kCocoLiteral + (r.sep ? 2 : 0) + len_i64(r.ann_id) + len_i32(r.frame) + len_i32(r.cls) + len_i32(r.cnt);
  if (r.cnt > 0) {
    const long long w = static_cast<long long>(r.x1) - r.x0 + 1, h = static_cast<long long>(r.y1) - r.y0 + 1;
    n += len_i32(r.x0) + len_i32(r.y0) + len_i64(w) + len_i64(h) + 6;   // three ", "
  } else {
    n += 10;   // 0, 0, 0, 0
  }
  return n + len_units6(r.neg_occ, r.q_occ) + len_units6(r.neg_trunc, r.q_trunc);
}

__device__ __forceinline__ void coco_print(Stage& w, const CocoFields& r) {
  if (r.sep) w.lit(", ");
  w.lit("{\"id\": ");
  w.i64(r.ann_id);
  w.lit(", \"image_id\": ");
  w.i32(r.frame);
  w.lit(", \"category_id\": ");
  w.i32(r.cls);
  w.lit(", \"bbox\": [");
  if (r.cnt > 0) {
    w.i32(r.x0);
    w.lit(", ");
    w.i32(r.y0);
    w.lit(", ");
    w.i64(static_cast<long long>(r.x1) - r.x0 + 1);
    w.lit(", ");
    w.i64(static_cast<long long>(r.y1) - r.y0 + 1);
  } else {
    w.lit("0, 0, 0, 0");
  }
  w.lit("], \"area\": ");
  w.i32(r.cnt);
  w.lit(", \"iscrowd\": 0, \"occlusion\": ");
  w.units6(r.neg_occ, r.q_occ);
  w.lit(", \"truncation\": ");
  w.units6(r.neg_trunc, r.q_trunc);
  w.put('}');
}

__global__ void __launch_bounds__(kCocoThreads)
    coco_text_kernel(const cspe_record* records, const int32_t* n_out, int B, int N, unsigned long long* ann_state,
                     char* text, long long frame_stride, int32_t* n_bytes) {
  __shared__ int warp_sums[kCocoThreads / 32];
  __shared__ long long base_s, first_id_s;
  __shared__ long long part_s[kCocoThreads / 32];
  __shared__ int bad_s;
  __shared__ __align__(16) char stage_s[kCocoStage];
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
  for (int g = tid; g < f; g += kCocoThreads) {
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
    for (int w = 0; w < kCocoThreads / 32; ++w) s += part_s[w];
    first_id_s = static_cast<long long>(__ldcg(ann_state)) + s + 1;
  }
  __syncthreads();
  const cspe_record* rec = records + static_cast<long long>(f) * N;
  char* out = text + static_cast<long long>(f) * frame_stride;
  for (int r0 = 0; r0 < n; r0 += kCocoThreads) {
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
      const float occ = __ldcg(&q->occlusion), trunc = __ldcg(&q->truncation);
      fld.sep = fld.ann_id > 1;
      if (!(fabsf(occ) < 1048576.0f) || !(fabsf(trunc) < 1048576.0f)) {
        bad_s = 1;   // an unprintable ratio: the frame is flagged, nothing is written for the record
      } else {
        fld.neg_occ = signbit(occ);
        fld.neg_trunc = signbit(trunc);
        fld.q_occ = static_cast<unsigned long long>(__double2ll_rn(fabs(static_cast<double>(occ)) * 1e6));
        fld.q_trunc = static_cast<unsigned long long>(__double2ll_rn(fabs(static_cast<double>(trunc)) * 1e6));
        len = coco_len(fld);
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
    for (int w = 0; w < kCocoThreads / 32; ++w) {
      const int sw = warp_sums[w];
      if (w < wid) before += sw;
      total += sw;
    }
    // this pass's text goes to out + base_s; the stage mirrors the destination modulo 16
    const long long base = base_s;
    char* dst = out + base;
    const int a = static_cast<int>(reinterpret_cast<uintptr_t>(dst) & 15);
    if (len > 0) {
      Stage st{stage_s + a + before + incl - len};
      coco_print(st, fld);
    }
    __syncthreads();
    long long room = frame_stride - base;   // bytes past the frame's stride are dropped but counted
    room = room < 0 ? 0 : (room > total ? total : room);
    stage_to_global<kCocoThreads>(stage_s, dst, room, tid);
    __syncthreads();
    if (tid == 0) base_s = base + total;
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
                     long long* total, int variant) {
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
  const long long n_full = n;   // what the row holds; only what fits below `capacity` is stored
  if (off + n > capacity) n = capacity > off ? capacity - off : 0;
  char* dst = packed + off;
  if (variant == 2 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // rows start on a 16-byte boundary, their place in `packed` does not: every aligned 16-byte chunk of the destination
    // is cut out of TWO consecutive aligned 16-byte source chunks (the second one is the next thread's first: an L1 / L2
    // hit) with funnel shifts — the byte offset between the two grids is the same for the whole row — and leaves as
    // ONE store
    const int a = static_cast<int>(reinterpret_cast<uintptr_t>(dst) & 15);
    char* g0 = dst - a;
    const long long end = a + n;
    const uint4* src16 = reinterpret_cast<const uint4*>(src);
    const int o = (16 - a) & 15;                 // source byte offset of a destination chunk inside its source chunk
    const int wo = o >> 2;
    const unsigned sh = static_cast<unsigned>(o & 3) * 8u;
    const long long n_src16 = (n + 15) >> 4;     // source chunks that hold text
    for (long long c = static_cast<long long>(tid) * 16; c < end; c += kYoloThreads * 16) {
      if (c >= a && c + 16 <= end) {
        const long long k = (c - a) >> 4;        // (c - a) = 16 k + o
        const uint4 A = __ldcg(src16 + k);
        const uint4 Bq = (o != 0 && k + 1 < n_src16) ? __ldcg(src16 + k + 1) : make_uint4(0u, 0u, 0u, 0u);
        const unsigned w[9] = {A.x, A.y, A.z, A.w, Bq.x, Bq.y, Bq.z, Bq.w, 0u};
        uint4 v;
        switch (wo) {   // uniform over the CTA
          case 0: v = make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh)); break;
          case 1: v = make_uint4(__funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh), __funnelshift_r(w[4], w[5], sh)); break;
          case 2: v = make_uint4(__funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh), __funnelshift_r(w[4], w[5], sh), __funnelshift_r(w[5], w[6], sh)); break;
          default: v = make_uint4(__funnelshift_r(w[3], w[4], sh), __funnelshift_r(w[4], w[5], sh), __funnelshift_r(w[5], w[6], sh), __funnelshift_r(w[6], w[7], sh)); break;
        }
        *reinterpret_cast<uint4*>(g0 + c) = v;
      } else {
        const long long lo = c > a ? c : a, hi = c + 16 < end ? c + 16 : end;
        for (long long i = lo; i < hi; ++i) g0[i] = __ldcg(src + (i - a));
      }
    }
  } else if (variant >= 1 && (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (variant == 1 || (reinterpret_cast<uintptr_t>(src) & 15) != 0)) {
    // five consecutive source words per destination chunk (measured slower than the byte stores: 8.7 vs 6.9 us)
    const int a = static_cast<int>(reinterpret_cast<uintptr_t>(dst) & 15);
    char* g0 = dst - a;
    const long long end = a + n;
    const unsigned* src32 = reinterpret_cast<const unsigned*>(src);
    for (long long c = static_cast<long long>(tid) * 16; c < end; c += kYoloThreads * 16) {
      if (c >= a && c + 16 <= end) {
        const long long sidx = c - a;
        const long long w0 = sidx >> 2;
        const unsigned sh = static_cast<unsigned>(sidx & 3) * 8u;
        unsigned w[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = __ldcg(src32 + w0 + k);
        w[4] = (sh != 0u && (w0 + 4) * 4 < n) ? __ldcg(src32 + w0 + 4) : 0u;
        uint4 v;
        v.x = __funnelshift_r(w[0], w[1], sh);
        v.y = __funnelshift_r(w[1], w[2], sh);
        v.z = __funnelshift_r(w[2], w[3], sh);
        v.w = __funnelshift_r(w[3], w[4], sh);
        *reinterpret_cast<uint4*>(g0 + c) = v;
      } else {
        const long long lo = c > a ? c : a, hi = c + 16 < end ? c + 16 : end;
        for (long long i = lo; i < hi; ++i) g0[i] = __ldcg(src + (i - a));
      }
    }
  } else if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {   // one 16-byte load, sixteen single-byte stores
    for (long long i = static_cast<long long>(tid) * 16; i < n; i += kYoloThreads * 16) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src + i));
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (i + k < n) dst[i + k] = static_cast<char>((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
    }
  } else {
    for (long long i = tid; i < n; i += kYoloThreads) dst[i] = __ldcg(src + i);
  }
  if (f == B - 1 && tid == 0) *total = off + n_full;
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
  CSPE_CUDA_OK(launch_pdl(coco_text_kernel, dim3(static_cast<unsigned>(B)), dim3(kCocoThreads), 0,
                          static_cast<cudaStream_t>(stream), records, n_out, B, N,
                          reinterpret_cast<unsigned long long*>(ann_state), text, static_cast<long long>(frame_stride),
                          n_bytes));
  return CSPE_OK;
}

// CSPE_PACK_VARIANT (A/B, read per call): 0 = sixteen byte stores per chunk, 1 = five source words + funnel shift,
// 2 = two aligned 16-byte source chunks + funnel shift (default)
static int pack_variant() {
  const char* e = getenv("CSPE_PACK_VARIANT");
  return (e && *e) ? atoi(e) : 2;
}

extern "C" int cspe_pack_rows(const char* text, int64_t frame_stride, const int32_t* n_bytes, int B, char* packed,
                              int64_t capacity, int64_t* total_bytes, void* stream) {
  CSPE_REQUIRE(B >= 0 && frame_stride >= 0 && capacity >= 0, CSPE_ERR_INVALID_ARGUMENT, "cspe_pack_rows: negative size");
  if (B == 0) return CSPE_OK;
  CSPE_REQUIRE(n_bytes && total_bytes && (frame_stride == 0 || text) && (capacity == 0 || packed), CSPE_ERR_INVALID_ARGUMENT,
               "cspe_pack_rows: null pointer");
  CSPE_CUDA_OK(launch_pdl(pack_rows_kernel, dim3(static_cast<unsigned>(B)), dim3(kYoloThreads), 0,
                          static_cast<cudaStream_t>(stream), text, static_cast<long long>(frame_stride), n_bytes, B, packed,
                          static_cast<long long>(capacity), reinterpret_cast<long long*>(total_bytes), pack_variant()));
  return CSPE_OK;
}
