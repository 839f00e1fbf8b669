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
