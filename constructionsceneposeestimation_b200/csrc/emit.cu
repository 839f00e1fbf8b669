// K4 — occlusion ratios, order-preserving compaction, record emission and per-class
// histogram.
//
// Replaces the reference's pose_list assembly (gcd.py:1938-1946): one record per kept
// object, in inst_idx order (gcd.py:1876-1886 numbers objects in first-seen order and
// K4 must keep that order, so compaction is a prefix sum, never an atomic append).  The
// per-class histogram generalises the reference's single object counter (gcd.py:361-372).
//
// One CTA per frame.  Bytes are negligible next to K1 (408 B per kept object).
#include <math.h>

#include "cspe_common.cuh"

static_assert(sizeof(cspe_record) == 408, "cspe_record layout changed: update the host dtype");

namespace cspe {
namespace {

constexpr int kEmitThreads = 256;

__global__ void __launch_bounds__(kEmitThreads)
    emit_kernel(const int32_t* __restrict__ scan, const double* __restrict__ uv, const double* __restrict__ z,
                const double* __restrict__ pose, const double* __restrict__ loose, const uint8_t* __restrict__ flags,
                const int32_t* __restrict__ slot_class, int N, int H, int W, int min_pixels, int frame_base,
                cspe_record* __restrict__ records, int32_t* __restrict__ n_out,
                unsigned long long* __restrict__ class_hist) {
  __shared__ int warp_sums[kEmitThreads / 32];
  __shared__ int base_s;
  __shared__ int hist_s[CSPE_NUM_CLASSES];
  const int f = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  if (tid < CSPE_NUM_CLASSES) hist_s[tid] = 0;
  if (tid == 0) base_s = 0;
  __syncthreads();

  for (int n0 = 0; n0 < N; n0 += kEmitThreads) {
    const int n = n0 + tid;
    const long long o = static_cast<long long>(f) * N + n;
    bool keep = false;
    int cls = -1, cnt = 0;
    uint8_t fl = 0;
    if (n < N) {
      cls = slot_class[o];
      cnt = scan[o * CSPE_SCAN_FIELDS + CSPE_SCAN_COUNT];
      fl = flags[o];
      keep = (cls >= 0) && (cnt >= min_pixels) && (fl & CSPE_OBJ_ANY_FRONT);
    }
    // stable rank = exclusive prefix sum of keep flags
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    int before = base_s;
    int chunk_total = 0;
#pragma unroll
    for (int w = 0; w < kEmitThreads / 32; ++w) {
      const int s = warp_sums[w];
      if (w < wid) before += s;
      chunk_total += s;
    }
    if (keep) {
      const int rank = before + in_warp;
      cspe_record* r = records + static_cast<long long>(f) * N + rank;
      const int x0 = scan[o * CSPE_SCAN_FIELDS + CSPE_SCAN_XMIN], y0 = scan[o * CSPE_SCAN_FIELDS + CSPE_SCAN_YMIN];
      const int x1 = scan[o * CSPE_SCAN_FIELDS + CSPE_SCAN_XMAX], y1 = scan[o * CSPE_SCAN_FIELDS + CSPE_SCAN_YMAX];
      r->frame = frame_base + f;
      r->inst_idx = n;
      r->class_id = cls;
      r->count = cnt;
      r->x_min = x0;
      r->y_min = y0;
      r->x_max = x1;
      r->y_max = y1;
      r->flags = fl;
      r->pad0 = 0;
      const int tw = x1 - x0 + 1, th = y1 - y0 + 1;
      const long long tight_area = static_cast<long long>(tw) * th;
      // loose box: projected 3D box clipped to the image and integerised
      const double umin = loose[o * 4 + 0], vmin = loose[o * 4 + 1], umax = loose[o * 4 + 2], vmax = loose[o * 4 + 3];
      const double dW = static_cast<double>(W), dH = static_cast<double>(H);
      const int lx0 = static_cast<int>(fmin(fmax(floor(umin), 0.0), dW));
      const int ly0 = static_cast<int>(fmin(fmax(floor(vmin), 0.0), dH));
      const int lx1 = static_cast<int>(fmax(fmin(ceil(umax) - 1.0, dW - 1.0), -1.0));
      const int ly1 = static_cast<int>(fmax(fmin(ceil(vmax) - 1.0, dH - 1.0), -1.0));
      const int lw = max(0, lx1 - lx0 + 1), lh = max(0, ly1 - ly0 + 1);
      const long long loose_area = static_cast<long long>(lw) * lh;
      if (loose_area > 0) {
        r->loose[0] = lx0;
        r->loose[1] = ly0;
        r->loose[2] = lx1;
        r->loose[3] = ly1;
      } else {
        r->loose[0] = 0;
        r->loose[1] = 0;
        r->loose[2] = -1;
        r->loose[3] = -1;
      }
      const float fcnt = static_cast<float>(cnt);
      const float vis = loose_area > 0 ? fminf(1.0f, fcnt / static_cast<float>(loose_area)) : 0.0f;
      r->visible_frac = vis;
      r->occlusion = 1.0f - vis;
      r->fill = cnt > 0 ? fcnt / static_cast<float>(tight_area) : 0.0f;  // min_pixels == 0 keeps unseen objects
      const double ua = (umax - umin) * (vmax - vmin);
      const double cwid = fmax(fmin(umax, dW) - fmax(umin, 0.0), 0.0);
      const double chei = fmax(fmin(vmax, dH) - fmax(vmin, 0.0), 0.0);
      const double ca = cwid * chei;
      r->truncation = ua > 0.0 ? 1.0f - static_cast<float>(ca) / static_cast<float>(ua) : 1.0f;
      const float fW = static_cast<float>(W), fH = static_cast<float>(H);
      r->yolo[0] = cnt > 0 ? (static_cast<float>(x0 + x1 + 1) * 0.5f) / fW : 0.0f;
      r->yolo[1] = cnt > 0 ? (static_cast<float>(y0 + y1 + 1) * 0.5f) / fH : 0.0f;
      r->yolo[2] = cnt > 0 ? static_cast<float>(tw) / fW : 0.0f;
      r->yolo[3] = cnt > 0 ? static_cast<float>(th) / fH : 0.0f;
#pragma unroll
      for (int k = 0; k < 16; ++k) r->uv[k] = uv[o * 16 + k];
#pragma unroll
      for (int k = 0; k < 8; ++k) r->z[k] = z[o * 8 + k];
#pragma unroll
      for (int k = 0; k < CSPE_POSE_STRIDE; ++k) r->pose[k] = pose[o * CSPE_POSE_STRIDE + k];
      if (cls < CSPE_NUM_CLASSES) atomicAdd(&hist_s[cls], 1);
    }
    __syncthreads();
    if (tid == 0) base_s += chunk_total;
    __syncthreads();
  }
  if (tid == 0) n_out[f] = base_s;
  if (tid < CSPE_NUM_CLASSES && hist_s[tid] > 0)
    atomicAdd(class_hist + tid, static_cast<unsigned long long>(hist_s[tid]));
}

}  // namespace
}  // namespace cspe

using namespace cspe;

extern "C" int cspe_emit(const int32_t* scan, const double* uv, const double* z, const double* pose,
                         const double* loose, const uint8_t* flags, const int32_t* slot_class, int B, int N, int H,
                         int W, int min_pixels, int frame_base, cspe_record* records, int32_t* n_out,
                         int64_t* class_hist, void* stream) {
  CSPE_REQUIRE(B >= 0 && N >= 0 && H >= 0 && W >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_emit: negative size (B=%d N=%d H=%d W=%d)", B, N, H, W);
  if (B == 0) return CSPE_OK;
  CSPE_REQUIRE(n_out && class_hist, CSPE_ERR_INVALID_ARGUMENT, "cspe_emit: n_out/class_hist is null");
  CSPE_REQUIRE(N == 0 || (scan && uv && z && pose && loose && flags && slot_class && records),
               CSPE_ERR_INVALID_ARGUMENT, "cspe_emit: null pointer");
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(records) & 7) == 0 && (reinterpret_cast<uintptr_t>(class_hist) & 7) == 0,
               CSPE_ERR_INVALID_ARGUMENT, "cspe_emit: records/class_hist must be 8-byte aligned");
  emit_kernel<<<static_cast<unsigned>(B), kEmitThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      scan, uv, z, pose, loose, flags, slot_class, N, H, W, min_pixels, frame_base, records, n_out,
      reinterpret_cast<unsigned long long*>(class_hist));
  CSPE_LAUNCH_OK("emit_kernel");
  return CSPE_OK;
}
