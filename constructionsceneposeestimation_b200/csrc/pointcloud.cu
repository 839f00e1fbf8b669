// f1 — depth -> coloured point cloud (gcd.py:616-711
// depth_to_pointcloud_with_rgb, the heaviest numeric function that exists in the reference).
//
// valid = isfinite & > 0 & < 250 (gcd.py:655); pinhole back-projection (gcd.py:664-666);
// rotate by the camera-pose quaternion's matrix and translate (gcd.py:677-685 — the reference
// applies the USD-axes pose directly to +Z-forward pinhole coordinates; kept as is);
// RGB gather with the "max <= 1 -> x255" rule (gcd.py:691-696); ROW-MAJOR STABLE compaction
// into (N, 6) float64.
//
// Three launches: (1) per-tile valid counts + max RGB over valid pixels, (2) one-CTA exclusive
// scan of tile counts, (3) recompute validity, in-tile ranks, write.  HBM-bound: reads
// 4*HW (depth) twice + C*HW (rgb), writes 48 B per point.
#include <math.h>

#include "cspe_common.cuh"

namespace cspe {
namespace {

constexpr int kPcThreads = 256;
constexpr int kPcPerThread = 4;
constexpr int kPcTile = kPcThreads * kPcPerThread;  // 1024 pixels

struct PcWorkspace {  // layout of the caller-provided scratch
  unsigned int rgb_max;
  unsigned int pad;
  long long total;
  // followed by: int32 tile_count[tiles]; int64 tile_offset[tiles]
};

__device__ __forceinline__ bool pc_valid(float d) {
  return (d > 0.0f) && (d < 250.0f);  // finite follows from < 250; NaN fails both
}

__global__ void __launch_bounds__(kPcThreads)
    pc_count_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int C, long long hw,
                    PcWorkspace* ws, int32_t* tile_count) {
  __shared__ int s_cnt[kPcThreads / 32];
  __shared__ unsigned s_max[kPcThreads / 32];
  const long long base = static_cast<long long>(blockIdx.x) * kPcTile + threadIdx.x * kPcPerThread;
  int cnt = 0;
  unsigned mx = 0;
#pragma unroll
  for (int k = 0; k < kPcPerThread; ++k) {
    const long long i = base + k;
    if (i < hw && pc_valid(__ldg(depth + i))) {
      ++cnt;
      if (rgb) {
        const uint8_t* c = rgb + i * C;
        mx = max(mx, max(static_cast<unsigned>(c[0]), max(static_cast<unsigned>(c[1]), static_cast<unsigned>(c[2]))));
      }
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  mx = __reduce_max_sync(0xffffffffu, mx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s_cnt[wid] = cnt;
    s_max[wid] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) {
      t += s_cnt[w];
      m = max(m, s_max[w]);
    }
    tile_count[blockIdx.x] = t;
    if (m) atomicMax(&ws->rgb_max, m);
  }
}

__global__ void __launch_bounds__(1024) pc_scan_kernel(const int32_t* __restrict__ tile_count, long long* tile_offset,
                                                      int tiles, PcWorkspace* ws, long long* n_points) {
  __shared__ long long s_warp[32];
  __shared__ long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < tiles; t0 += 1024) {
    const int t = t0 + tid;
    const long long v = t < tiles ? tile_count[t] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      long long w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const long long before = s_base + (wid ? s_warp[wid - 1] : 0) + inc - v;
    if (t < tiles) tile_offset[t] = before;
    __syncthreads();
    if (tid == 0) s_base += s_warp[31];
    __syncthreads();
  }
  if (tid == 0) {
    ws->total = s_base;
    *n_points = s_base;
  }
}

__global__ void __launch_bounds__(kPcThreads)
    pc_write_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int C, int W, long long hw,
                    const double* __restrict__ cam, const PcWorkspace* __restrict__ ws,
                    const long long* __restrict__ tile_offset, double* __restrict__ out, long long capacity) {
  __shared__ int s_warp[kPcThreads / 32];
  const long long base = static_cast<long long>(blockIdx.x) * kPcTile + threadIdx.x * kPcPerThread;
  float d[kPcPerThread];
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kPcPerThread; ++k) {
    const long long i = base + k;
    d[k] = i < hw ? __ldg(depth + i) : 0.0f;
    cnt += pc_valid(d[k]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int before = inc - cnt;
  for (int w = 0; w < wid; ++w) before += s_warp[w];
  if (cnt == 0) return;

  const double t0 = cam[0], t1 = cam[1], t2 = cam[2];
  const double fx = cam[12], fy = cam[13], cx = cam[14], cy = cam[15];
  const bool scale255 = ws->rgb_max <= 1u;  // gcd.py:693
  long long rank = tile_offset[blockIdx.x] + before;
#pragma unroll
  for (int k = 0; k < kPcPerThread; ++k) {
    if (!pc_valid(d[k])) continue;
    const long long i = base + k;
    if (rank < capacity) {
      const int v = static_cast<int>(i / W);
      const int u = static_cast<int>(i - static_cast<long long>(v) * W);
      const double zc = static_cast<double>(d[k]);
      const double xc = ((static_cast<double>(u) - cx) * zc) / fx;
      const double yc = ((static_cast<double>(v) - cy) * zc) / fy;
      double* o = out + rank * 6;
      o[0] = ((cam[3] * xc + cam[4] * yc) + cam[5] * zc) + t0;
      o[1] = ((cam[6] * xc + cam[7] * yc) + cam[8] * zc) + t1;
      o[2] = ((cam[9] * xc + cam[10] * yc) + cam[11] * zc) + t2;
      if (rgb) {
        const uint8_t* c = rgb + i * C;
        const unsigned m = scale255 ? 255u : 1u;
        o[3] = static_cast<double>((c[0] * m) & 0xffu);
        o[4] = static_cast<double>((c[1] * m) & 0xffu);
        o[5] = static_cast<double>((c[2] * m) & 0xffu);
      } else {
        o[3] = o[4] = o[5] = 255.0;  // gcd.py:698-700
      }
    }
    ++rank;
  }
}

}  // namespace
}  // namespace cspe

using namespace cspe;

static long long pc_tiles(long long hw) { return (hw + kPcTile - 1) / kPcTile; }

extern "C" size_t cspe_pointcloud_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return sizeof(PcWorkspace);
  const long long tiles = pc_tiles(static_cast<long long>(H) * W);
  // header | int64 tile_offset[tiles] | int32 tile_count[tiles]
  return sizeof(PcWorkspace) + static_cast<size_t>(tiles) * (8 + 4) + 16;
}

extern "C" int cspe_depth_to_pointcloud(const float* depth, const uint8_t* rgb, int rgb_channels, int H, int W,
                                        const double* cam, double* out, int64_t capacity, int64_t* n_points,
                                        void* workspace, void* stream) {
  CSPE_REQUIRE(H >= 0 && W >= 0 && capacity >= 0, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: negative size");
  CSPE_REQUIRE(n_points && workspace && cam, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: null pointer");
  CSPE_REQUIRE(rgb == nullptr || rgb_channels >= 3, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_depth_to_pointcloud: rgb needs >= 3 channels (got %d)", rgb_channels);
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
               CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: workspace/out must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long hw = static_cast<long long>(H) * W;
  PcWorkspace* ws = static_cast<PcWorkspace*>(workspace);
  CSPE_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(PcWorkspace), st));
  if (hw == 0) {
    CSPE_CUDA_OK(cudaMemsetAsync(n_points, 0, sizeof(int64_t), st));
    return CSPE_OK;
  }
  CSPE_REQUIRE(depth != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: depth is null");
  CSPE_REQUIRE(capacity == 0 || out != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: out is null");
  const long long tiles = pc_tiles(hw);
  CSPE_REQUIRE(tiles < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_depth_to_pointcloud: frame too large");
  long long* tile_offset = reinterpret_cast<long long*>(ws + 1);
  int32_t* tile_count = reinterpret_cast<int32_t*>(tile_offset + tiles);
  pc_count_kernel<<<static_cast<unsigned>(tiles), kPcThreads, 0, st>>>(depth, rgb, rgb_channels, hw, ws, tile_count);
  CSPE_LAUNCH_OK("pc_count_kernel");
  pc_scan_kernel<<<1, 1024, 0, st>>>(tile_count, tile_offset, static_cast<int>(tiles), ws,
                                     reinterpret_cast<long long*>(n_points));
  CSPE_LAUNCH_OK("pc_scan_kernel");
  pc_write_kernel<<<static_cast<unsigned>(tiles), kPcThreads, 0, st>>>(depth, rgb, rgb_channels, W, hw, cam, ws,
                                                                      tile_offset, out, capacity);
  CSPE_LAUNCH_OK("pc_write_kernel");
  return CSPE_OK;
}
