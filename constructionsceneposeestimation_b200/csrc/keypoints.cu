// K3 — skeleton keypoint projection + depth-buffer visibility.
//
// [SPEC] stage: the reference only knows people as a class keyword ("skelroot", gcd.py:105);
// the projection follows the same camera conventions as K2 (gcd.py:587-605 pose,
// gcd.py:646-649 intrinsics) and the depth test reads the distance_to_image_plane buffer the
// reference captures at gcd.py:1681 (inf = no hit, gcd.py:318-321).
//
// One thread per joint: 12 B in, one 4-byte depth gather, 25 B out — negligible traffic; the
// launch is amortised over the batch.  FP64, -fmad=false (bit-identical to the numpy oracle).
#include <math.h>

#include "cspe_common.cuh"

namespace cspe {
namespace {

__device__ __forceinline__ void keypoint_one(const float* __restrict__ joints, long long i, int per_frame,
                                             const float* __restrict__ depth, int H, int W,
                                             const double* __restrict__ cam, double tol, double* __restrict__ kp,
                                             double* __restrict__ kz, uint8_t* __restrict__ vis);

__global__ void __launch_bounds__(256)
    keypoints_kernel(const float* __restrict__ joints, long long total, int per_frame, const float* __restrict__ depth,
                     int H, int W, const double* __restrict__ cam, double tol, double* __restrict__ kp,
                     double* __restrict__ kz, uint8_t* __restrict__ vis, int overlap_previous) {
  // programmatic dependent launch, same protocol as K2 (project.cu): release whatever is queued
  // next, then either behave like a serialised launch (wait now) or — overlapped mode, inputs not
  // written by the previous kernel — run beside it and wait only before exiting
  pdl_launch_dependents();
  if (!overlap_previous) pdl_wait();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride)
    keypoint_one(joints, i, per_frame, depth, H, W, cam, tol, kp, kz, vis);
  if (overlap_previous) pdl_wait();
}

__device__ __forceinline__ void keypoint_one(const float* __restrict__ joints, long long i, int per_frame,
                                             const float* __restrict__ depth, int H, int W,
                                             const double* __restrict__ cam, double tol, double* __restrict__ kp,
                                             double* __restrict__ kz, uint8_t* __restrict__ vis) {
  const int frame = static_cast<int>(i / per_frame);
  // in overlapped mode everything here is read before the PDL wait: L1 bypass (cspe_common.cuh, PDL rule)
  double cm[17];
#pragma unroll
  for (int j = 0; j < 17; ++j) cm[j] = __ldcg(cam + static_cast<long long>(frame) * CSPE_CAM_STRIDE + j);
  const double d0 = static_cast<double>(__ldcg(joints + i * 3 + 0)) - cm[0];
  const double d1 = static_cast<double>(__ldcg(joints + i * 3 + 1)) - cm[1];
  const double d2 = static_cast<double>(__ldcg(joints + i * 3 + 2)) - cm[2];
  // p_c = Rcw^T d  (column i of Rcw)
  const double pc0 = (cm[3] * d0 + cm[6] * d1) + cm[9] * d2;
  const double pc1 = (cm[4] * d0 + cm[7] * d1) + cm[10] * d2;
  const double pc2 = (cm[5] * d0 + cm[8] * d1) + cm[11] * d2;
  const double z = -pc2;
  const double u = cm[14] + (cm[12] * pc0) / z;
  const double v = cm[15] - (cm[13] * pc1) / z;
  kp[i * 2 + 0] = u;
  kp[i * 2 + 1] = v;
  kz[i] = z;
  const bool in_view = (u >= 0.0) && (u < static_cast<double>(W)) && (v >= 0.0) && (v < static_cast<double>(H)) &&
                       (z > cm[16]);
  uint8_t flag = CSPE_KP_OUT;
  if (in_view) {
    const int ui = static_cast<int>(floor(u));
    const int vi = static_cast<int>(floor(v));
    const float dz = __ldcg(depth + (static_cast<long long>(frame) * H + vi) * W + ui);
    const bool visible = isfinite(dz) && (z <= static_cast<double>(dz) + tol);
    flag = visible ? CSPE_KP_VISIBLE : CSPE_KP_OCCLUDED;
  }
  vis[i] = flag;
}

}  // namespace
}  // namespace cspe

using namespace cspe;

static int keypoints_impl(const float* joints, int B, int P, int J, const float* depth, int H, int W, const double* cam,
                          double tol, double* kp, double* kz, uint8_t* vis, void* stream, int overlap_previous) {
  CSPE_REQUIRE(B >= 0 && P >= 0 && J >= 0 && H >= 0 && W >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_keypoints: negative size (B=%d P=%d J=%d H=%d W=%d)", B, P, J, H, W);
  const long long per_frame = static_cast<long long>(P) * J;
  const long long total = per_frame * B;
  if (total == 0) return CSPE_OK;
  CSPE_REQUIRE(per_frame < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_keypoints: P*J too large");
  CSPE_REQUIRE(joints && cam && kp && kz && vis, CSPE_ERR_INVALID_ARGUMENT, "cspe_keypoints: null pointer");
  CSPE_REQUIRE(depth != nullptr || H == 0 || W == 0, CSPE_ERR_INVALID_ARGUMENT, "cspe_keypoints: depth is null");
  // standalone: one joint per thread.  Overlapped: one small block per SM with a grid-stride loop —
  // its blocks stay resident beside the mask scan until that finishes (wait-at-exit), so they must
  // all fit in the few registers the scan leaves free; the loop is hidden behind the scan anyway.
  long long blocks = (total + 255) / 256;
  int threads = 256;
  if (overlap_previous) {
    const int sms = sm_count();
    CSPE_REQUIRE(sms > 0, CSPE_ERR_NO_DEVICE, "cspe_keypoints: no CUDA device");
    threads = 64;
    blocks = (total + threads - 1) / threads;
    if (blocks > sms) blocks = sms;
  }
  CSPE_REQUIRE(blocks < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_keypoints: too many joints");
  CSPE_CUDA_OK(launch_pdl(keypoints_kernel, dim3(static_cast<unsigned>(blocks)), dim3(threads), 0,
                          static_cast<cudaStream_t>(stream), joints, total, static_cast<int>(per_frame), depth, H, W, cam,
                          tol, kp, kz, vis, overlap_previous));
  return CSPE_OK;
}

extern "C" int cspe_keypoints(const float* joints, int B, int P, int J, const float* depth, int H, int W,
                              const double* cam, double tol, double* kp, double* kz, uint8_t* vis, void* stream) {
  return keypoints_impl(joints, B, P, J, depth, H, W, cam, tol, kp, kz, vis, stream, 0);
}

extern "C" int cspe_keypoints_overlapped(const float* joints, int B, int P, int J, const float* depth, int H, int W,
                                         const double* cam, double tol, double* kp, double* kz, uint8_t* vis,
                                         void* stream) {
  return keypoints_impl(joints, B, P, J, depth, H, W, cam, tol, kp, kz, vis, stream, 1);
}
