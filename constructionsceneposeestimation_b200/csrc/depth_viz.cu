// f4 — depth visualisation and RGB(A) -> BGR (gcd.py:1691-1709 and 1671).
//
// depth_normalized[valid] = ((d - min) / (max - min + 1e-6) * 255).astype(uint8) over valid =
// isfinite & > 0 pixels (float32 arithmetic, truncating cast), 0 elsewhere, then a 256-entry BGR
// colour LUT (the reference uses cv2.COLORMAP_JET; the caller passes the table so the kernel does
// not hard-code OpenCV's values).  min / max come from cspe_depth_stats (gcd.py:1693-1694 takes
// them from the same valid set).  No valid pixel -> black image (gcd.py:1705-1709).
//
// Both kernels are HBM-bound elementwise passes (4 B read + 3 B written per pixel).  Fast path: a
// block takes 4096 pixels per step — warp-contiguous 16-byte loads, the 24-bit pixels packed with
// byte permutes into a shared-memory stage, warp-contiguous 16-byte streaming stores — when the
// frame geometry keeps every access 16-byte aligned; otherwise a scalar path.
#include "cspe_common.cuh"

namespace cspe {
namespace {

constexpr int kVizThreads = 256;

__device__ __forceinline__ unsigned depth_bin(float d, float mn, float den) {
  const unsigned b = __float_as_uint(d);
  if (!((b - 1u) < 0x7f7fffffu)) return 0u;  // not (finite and > 0)
  const float t = ((d - mn) / den) * 255.0f;
  return min(__float2uint_rz(t), 255u);
}

// The IEEE division costs ~25 instructions per pixel and made the kernel issue-bound.  One multiply
// by 255 / den gives t within 255 * 2^-22 < 1e-4 of the reference's ((d - min) / den) * 255 (two
// roundings each way), so whenever t is further than 1e-3 from an integer — 99.8 % of the pixels —
// its truncation IS the reference's bin; the rest (and anything outside [0, 255)) takes the exact
// formula.  Bit-identical to numpy by construction, checked on every float around every bin boundary.
__device__ __forceinline__ unsigned depth_bin_fast(float d, float mn, float den, float scale) {
  const unsigned b = __float_as_uint(d);
  if (!((b - 1u) < 0x7f7fffffu)) return 0u;  // not (finite and > 0)
  const float t = (d - mn) * scale;
  const float tr = truncf(t);
  const float f = t - tr;
  if (!(t >= 0.0f && t < 255.0f) || f < 1e-3f || f > 0.999f) return depth_bin(d, mn, den);
  return static_cast<unsigned>(tr);
}

// four 24-bit pixels c0..c3 (colour in the low 3 bytes) -> three packed 32-bit words
__device__ __forceinline__ void pack4(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned& w0, unsigned& w1,
                                      unsigned& w2) {
  w0 = __byte_perm(c0, c1, 0x4210);
  w1 = __byte_perm(c1, c2, 0x5421);
  w2 = __byte_perm(c2, c3, 0x6542);
}

__device__ __forceinline__ void st_stream(uint4* p, const uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

constexpr int kVizChunkPx = kVizThreads * 16;         // pixels per block iteration (4096)
constexpr int kVizChunkWords = kVizChunkPx * 3 / 4;    // packed BGR words per iteration (3072)

// The block's packed output for one 4096-pixel chunk sits in shared memory; stream it out with
// warp-contiguous 16-byte stores.  (Storing each thread's 12 / 48 bytes directly is 3x the sector
// traffic: the lanes of one store instruction land 12 or 48 bytes apart.)
__device__ __forceinline__ void flush_chunk(const unsigned* stage, uint8_t* dst) {
  __syncthreads();
  const uint4* s4 = reinterpret_cast<const uint4*>(stage);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < kVizChunkWords / 4 / kVizThreads; ++j) st_stream(d4 + threadIdx.x + j * kVizThreads, s4[threadIdx.x + j * kVizThreads]);
  __syncthreads();
}

__global__ void __launch_bounds__(kVizThreads)
    depth_colormap_kernel(const float* __restrict__ depth, long long hw, const cspe_depth_stats_t* __restrict__ stats,
                          const uint8_t* __restrict__ lut, uint8_t* __restrict__ out, int vec_ok) {
  __shared__ unsigned lut_s[256];  // BGR in the low three bytes
  __shared__ __align__(16) unsigned stage[kVizChunkWords];
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    lut_s[i] = lut[i * 3] | (lut[i * 3 + 1] << 8) | (static_cast<unsigned>(lut[i * 3 + 2]) << 16);
  const int b = blockIdx.y;
  const cspe_depth_stats_t st = stats[b];
  const float mn = st.depth_min;
  const float den = (st.depth_max - st.depth_min) + 1e-6f;
  const float scale = 255.0f / den;
  const bool any_valid = st.valid_pixels > 0;
  __syncthreads();
  const float* d = depth + static_cast<long long>(b) * hw;
  uint8_t* o = out + static_cast<long long>(b) * hw * 3;
  const long long nchunks = vec_ok ? hw / kVizChunkPx : 0;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(d + ch * kVizChunkPx);
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld_stream(src + threadIdx.x + k * kVizThreads);   // warp-contiguous loads
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned c[4];
      const float f[4] = {__uint_as_float(v[k].x), __uint_as_float(v[k].y), __uint_as_float(v[k].z),
                          __uint_as_float(v[k].w)};
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = any_valid ? lut_s[depth_bin_fast(f[j], mn, den, scale)] : 0u;
      unsigned w0, w1, w2;
      pack4(c[0], c[1], c[2], c[3], w0, w1, w2);
      unsigned* sp = stage + (threadIdx.x + k * kVizThreads) * 3;   // 3-word stride: bank-conflict free
      sp[0] = w0;
      sp[1] = w1;
      sp[2] = w2;
    }
    flush_chunk(stage, o + ch * (kVizChunkPx * 3));
  }
  // leftover pixels (and every pixel when the geometry is not 16-byte friendly)
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = nchunks * kVizChunkPx + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw;
       i += stride) {
    const unsigned c = any_valid ? lut_s[depth_bin_fast(d[i], mn, den, scale)] : 0u;
    o[i * 3 + 0] = static_cast<uint8_t>(c);
    o[i * 3 + 1] = static_cast<uint8_t>(c >> 8);
    o[i * 3 + 2] = static_cast<uint8_t>(c >> 16);
  }
}

// RGBA (4 channels): 4096 pixels = 16 KB in, 12 KB out per block iteration
__global__ void __launch_bounds__(kVizThreads)
    rgba_to_bgr_kernel(const uint8_t* __restrict__ rgb, long long n, uint8_t* __restrict__ bgr) {
  __shared__ __align__(16) unsigned stage[kVizChunkWords];
  const long long nchunks = n / kVizChunkPx;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(rgb + ch * (kVizChunkPx * 4));
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld_stream(src + threadIdx.x + k * kVizThreads);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // 0xAABBGGRR -> colour with B, G, R in bytes 0, 1, 2
      const unsigned c0 = __byte_perm(v[k].x, 0u, 0x4012), c1 = __byte_perm(v[k].y, 0u, 0x4012);
      const unsigned c2 = __byte_perm(v[k].z, 0u, 0x4012), c3 = __byte_perm(v[k].w, 0u, 0x4012);
      unsigned w0, w1, w2;
      pack4(c0, c1, c2, c3, w0, w1, w2);
      unsigned* sp = stage + (threadIdx.x + k * kVizThreads) * 3;
      sp[0] = w0;
      sp[1] = w1;
      sp[2] = w2;
    }
    flush_chunk(stage, bgr + ch * (kVizChunkPx * 3));
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = nchunks * kVizChunkPx + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    bgr[i * 3 + 0] = rgb[i * 4 + 2];
    bgr[i * 3 + 1] = rgb[i * 4 + 1];
    bgr[i * 3 + 2] = rgb[i * 4 + 0];
  }
}

// packed RGB (3 channels): same bytes in and out, R and B of every triple swapped; 12 bytes (4 pixels)
// per thread step keeps loads and stores warp-contiguous without staging
__global__ void __launch_bounds__(kVizThreads)
    rgb3_to_bgr_kernel(const uint8_t* __restrict__ rgb, long long n, uint8_t* __restrict__ bgr) {
  __shared__ __align__(16) unsigned stage[kVizChunkWords];
  const long long nchunks = n / kVizChunkPx;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(rgb + ch * (kVizChunkPx * 3));
    uint4* s4 = reinterpret_cast<uint4*>(stage);
#pragma unroll
    for (int j = 0; j < kVizChunkWords / 4 / kVizThreads; ++j) s4[threadIdx.x + j * kVizThreads] = ld_stream(src + threadIdx.x + j * kVizThreads);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // each thread swaps four 12-byte groups in place (3-word stride: conflict free)
      unsigned* sp = stage + (threadIdx.x + k * kVizThreads) * 3;
      const unsigned w0 = sp[0], w1 = sp[1], w2 = sp[2];
      sp[0] = __byte_perm(w0, w1, 0x5012);
      sp[1] = __byte_perm(__byte_perm(w1, w0, 0x3070), w2, 0x3410);
      sp[2] = __byte_perm(w1, w2, 0x5672);
    }
    flush_chunk(stage, bgr + ch * (kVizChunkPx * 3));
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = nchunks * kVizChunkPx + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    bgr[i * 3 + 0] = rgb[i * 3 + 2];
    bgr[i * 3 + 1] = rgb[i * 3 + 1];
    bgr[i * 3 + 2] = rgb[i * 3 + 0];
  }
}

// any other layout (more than 4 channels, unaligned buffers)
__global__ void __launch_bounds__(kVizThreads)
    rgb_to_bgr_generic_kernel(const uint8_t* __restrict__ rgb, int channels, long long n, uint8_t* __restrict__ bgr) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint8_t* s = rgb + i * channels;
    bgr[i * 3 + 0] = s[2];
    bgr[i * 3 + 1] = s[1];
    bgr[i * 3 + 2] = s[0];
  }
}

int resident_blocks(const void* kernel, int threads) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) occ = 4;
  return occ;
}

}  // namespace
}  // namespace cspe

using namespace cspe;

extern "C" int cspe_depth_colormap(const float* depth, int B, int H, int W, const cspe_depth_stats_t* stats,
                                   const uint8_t* lut_bgr, uint8_t* out, void* stream) {
  CSPE_REQUIRE(B >= 0 && H >= 0 && W >= 0, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_colormap: negative size");
  const long long hw = static_cast<long long>(H) * W;
  if (B == 0 || hw == 0) return CSPE_OK;
  CSPE_REQUIRE(depth && stats && lut_bgr && out, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_colormap: null pointer");
  CSPE_REQUIRE(B <= 65535, CSPE_ERR_UNSUPPORTED, "cspe_depth_colormap: B > 65535");
  const int sms = sm_count();
  CSPE_REQUIRE(sms > 0, CSPE_ERR_NO_DEVICE, "cspe_depth_colormap: no CUDA device");
  // one wave of resident blocks over the batch
  static const int occ = resident_blocks(reinterpret_cast<const void*>(depth_colormap_kernel), kVizThreads);
  long long per_frame = (static_cast<long long>(sms) * occ) / B;
  const long long max_useful = (hw + kVizChunkPx - 1) / kVizChunkPx;
  if (per_frame > max_useful) per_frame = max_useful;
  if (per_frame < 1) per_frame = 1;
  // 16-pixel path: every frame base must keep 16-byte alignment for loads (hw % 4) and stores (3*hw % 16)
  const int vec_ok = (reinterpret_cast<uintptr_t>(depth) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                     (B == 1 || hw % 16 == 0);
  dim3 grid(static_cast<unsigned>(per_frame), static_cast<unsigned>(B));
  depth_colormap_kernel<<<grid, kVizThreads, 0, static_cast<cudaStream_t>(stream)>>>(depth, hw, stats, lut_bgr, out,
                                                                                    vec_ok);
  CSPE_LAUNCH_OK("depth_colormap_kernel");
  return CSPE_OK;
}

extern "C" int cspe_rgb_to_bgr(const uint8_t* rgb, int channels, int64_t num_pixels, uint8_t* bgr, void* stream) {
  CSPE_REQUIRE(num_pixels >= 0, CSPE_ERR_INVALID_ARGUMENT, "cspe_rgb_to_bgr: negative size");
  CSPE_REQUIRE(channels >= 3, CSPE_ERR_INVALID_ARGUMENT, "cspe_rgb_to_bgr: needs >= 3 channels (got %d)", channels);
  if (num_pixels == 0) return CSPE_OK;
  CSPE_REQUIRE(rgb && bgr, CSPE_ERR_INVALID_ARGUMENT, "cspe_rgb_to_bgr: null pointer");
  const int sms = sm_count();
  CSPE_REQUIRE(sms > 0, CSPE_ERR_NO_DEVICE, "cspe_rgb_to_bgr: no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool aligned = (reinterpret_cast<uintptr_t>(rgb) % 16 == 0) && (reinterpret_cast<uintptr_t>(bgr) % 16 == 0);
  long long blocks = (num_pixels + kVizChunkPx - 1) / kVizChunkPx;
  if (blocks < 1) blocks = 1;
  if (aligned && channels == 4) {
    static const int occ = resident_blocks(reinterpret_cast<const void*>(rgba_to_bgr_kernel), kVizThreads);
    if (blocks > static_cast<long long>(sms) * occ) blocks = static_cast<long long>(sms) * occ;
    rgba_to_bgr_kernel<<<static_cast<unsigned>(blocks), kVizThreads, 0, st>>>(rgb, num_pixels, bgr);
  } else if (aligned && channels == 3) {
    static const int occ = resident_blocks(reinterpret_cast<const void*>(rgb3_to_bgr_kernel), kVizThreads);
    if (blocks > static_cast<long long>(sms) * occ) blocks = static_cast<long long>(sms) * occ;
    rgb3_to_bgr_kernel<<<static_cast<unsigned>(blocks), kVizThreads, 0, st>>>(rgb, num_pixels, bgr);
  } else {
    blocks = (num_pixels + kVizThreads * 8 - 1) / (kVizThreads * 8);
    if (blocks > static_cast<long long>(sms) * 16) blocks = static_cast<long long>(sms) * 16;
    rgb_to_bgr_generic_kernel<<<static_cast<unsigned>(blocks), kVizThreads, 0, st>>>(rgb, channels, num_pixels, bgr);
  }
  CSPE_LAUNCH_OK("rgb_to_bgr_kernel");
  return CSPE_OK;
}
