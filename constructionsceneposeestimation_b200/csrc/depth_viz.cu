// f4 — depth visualisation and RGB(A) -> BGR (gcd.py:1691-1709 and 1671).
//
// depth_normalized[valid] = ((d - min) / (max - min + 1e-6) * 255).astype(uint8) over valid =
// isfinite & > 0 pixels (float32 arithmetic, truncating cast), 0 elsewhere, then a 256-entry BGR
// colour LUT (the reference uses cv2.COLORMAP_JET; the caller passes the table so the kernel does
// not hard-code OpenCV's values).  min / max come from cspe_depth_stats (gcd.py:1693-1694 takes
// them from the same valid set).  No valid pixel -> black image (gcd.py:1705-1709).
// HBM-bound elementwise pass: 4 B read, 3 B written per pixel.
#include "cspe_common.cuh"

namespace cspe {
namespace {

__device__ __forceinline__ unsigned depth_bin(float d, float mn, float den) {
  const bool valid = (d > 0.0f) && (fabsf(d) != __int_as_float(0x7f800000));
  if (!valid) return 0u;
  const float t = ((d - mn) / den) * 255.0f;
  return min(__float2uint_rz(t), 255u);
}

// 4 pixels per thread when the frame base is 16-byte aligned: one LDG.128 in, three STG.32 out
__global__ void __launch_bounds__(256)
    depth_colormap_kernel(const float* __restrict__ depth, long long hw, const cspe_depth_stats_t* __restrict__ stats,
                          const uint8_t* __restrict__ lut, uint8_t* __restrict__ out, int vec_ok) {
  __shared__ uint8_t lut_s[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut_s[i] = lut[i];
  __syncthreads();
  const int b = blockIdx.y;
  const cspe_depth_stats_t st = stats[b];
  const float mn = st.depth_min;
  const float den = (st.depth_max - st.depth_min) + 1e-6f;
  const bool any_valid = st.valid_pixels > 0;
  const float* d = depth + static_cast<long long>(b) * hw;
  uint8_t* o = out + static_cast<long long>(b) * hw * 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long gtid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n4 = vec_ok ? hw / 4 : 0;
  for (long long i = gtid; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(d) + i);
    unsigned char px[12];
    const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned bin = any_valid ? depth_bin(vals[k], mn, den) : 0u;
#pragma unroll
      for (int c = 0; c < 3; ++c) px[k * 3 + c] = any_valid ? lut_s[bin * 3 + c] : 0;
    }
    uint32_t w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      w[k] = px[4 * k] | (px[4 * k + 1] << 8) | (px[4 * k + 2] << 16) | (static_cast<uint32_t>(px[4 * k + 3]) << 24);
    uint32_t* o32 = reinterpret_cast<uint32_t*>(o + i * 12);
    o32[0] = w[0];
    o32[1] = w[1];
    o32[2] = w[2];
  }
  for (long long i = n4 * 4 + gtid; i < hw; i += stride) {
    const unsigned bin = any_valid ? depth_bin(d[i], mn, den) : 0u;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[i * 3 + c] = any_valid ? lut_s[bin * 3 + c] : 0;
  }
}

__global__ void __launch_bounds__(256)
    rgb_to_bgr_kernel(const uint8_t* __restrict__ rgb, int channels, long long n, uint8_t* __restrict__ bgr) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint8_t* s = rgb + i * channels;
    bgr[i * 3 + 0] = s[2];
    bgr[i * 3 + 1] = s[1];
    bgr[i * 3 + 2] = s[0];
  }
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
  long long per_frame = (hw / 4 + 256 * 4 - 1) / (256 * 4);
  if (per_frame < 1) per_frame = 1;
  const long long want = (static_cast<long long>(sms) * 16 + B - 1) / B;
  if (per_frame > want) per_frame = want;
  dim3 grid(static_cast<unsigned>(per_frame), static_cast<unsigned>(B));
  // vector path: every frame base must be 16-byte aligned for the loads and 4-byte aligned for the stores
  const int vec_ok = (reinterpret_cast<uintptr_t>(depth) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 4 == 0) &&
                     (B == 1 || hw % 4 == 0);
  depth_colormap_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(depth, hw, stats, lut_bgr, out, vec_ok);
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
  long long blocks = (num_pixels + 256 * 8 - 1) / (256 * 8);
  if (blocks > static_cast<long long>(sms) * 16) blocks = static_cast<long long>(sms) * 16;
  rgb_to_bgr_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(rgb, channels,
                                                                                                 num_pixels, bgr);
  CSPE_LAUNCH_OK("rgb_to_bgr_kernel");
  return CSPE_OK;
}
