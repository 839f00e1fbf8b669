// f1 — depth -> coloured point cloud (gcd.py:616-711
// depth_to_pointcloud_with_rgb, the heaviest numeric function that exists in the reference).
//
// valid = isfinite & > 0 & < 250 (gcd.py:655); pinhole back-projection (gcd.py:664-666);
// rotate by the camera-pose quaternion's matrix and translate (gcd.py:677-685 — the reference
// applies the USD-axes pose directly to +Z-forward pinhole coordinates; kept as is);
// RGB gather with the "max <= 1 -> x255" rule (gcd.py:691-696); ROW-MAJOR STABLE compaction
// into (N, 6) float64.
//
// HBM-bound on the OUTPUT: 48 bytes written per valid pixel against 4 (depth) + C (rgb) bytes read,
// both read twice.  Two launches chained by programmatic dependent launch (a single 1080p frame is
// latency-bound, every launch costs):
//   1. per tile (1024 px): valid count and the maximum colour over valid pixels; the LAST CTA to
//      finish (atomic ticket) scans the tile counts into tile offsets and the total;
//   2. per tile: recompute validity, in-tile ranks, build the tile's points in shared memory and
//      stream them out as one contiguous run of 16-byte stores (a thread writing its own 48-byte
//      points directly costs ~6x the sector traffic); colours are scaled by 255 when the maximum
//      found by pass 1 is <= 1 (the reference's [0,1]-image rule, gcd.py:693).
#include <math.h>

#include "cspe_common.cuh"

namespace cspe {
namespace {

constexpr int kPcThreads = 256;
constexpr int kPcPerThread = 4;
constexpr int kPcTile = kPcThreads * kPcPerThread;  // 1024 pixels
constexpr int kPcStageBytes = kPcTile * (3 * 8 + 4);  // a tile's points as x[], y[], z[] (f64) + packed colour: 28 KB

struct PcWorkspace {  // layout of the caller-provided scratch
  unsigned int rgb_max;  // maximum colour byte over valid pixels
  unsigned int done;     // CTAs of pass 1 that have published their count (ticket for the last-CTA scan)
  long long total;
  // followed by: int64 tile_offset[tiles]; int32 tile_count[tiles]
};

__device__ __forceinline__ bool pc_valid(float d) {
  return (d > 0.0f) && (d < 250.0f);  // finite follows from < 250; NaN fails both
}

__device__ __forceinline__ void load4(const float* __restrict__ depth, long long base, long long hw, bool vec, float* d) {
  if (vec && base + 3 < hw) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(depth + base));
    d[0] = v.x;
    d[1] = v.y;
    d[2] = v.z;
    d[3] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) d[k] = base + k < hw ? __ldg(depth + base + k) : 0.0f;
  }
}

// colours of the thread's four pixels: one 16-byte load for RGBA, byte loads otherwise
__device__ __forceinline__ void load_rgb4(const uint8_t* __restrict__ rgb, int C, long long base, long long hw, bool vec,
                                          uint32_t* px) {
  if (vec && base + 3 < hw) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rgb + base * 4));
    px[0] = v.x, px[1] = v.y, px[2] = v.z, px[3] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) {
      px[k] = 0;
      if (base + k < hw) {
        const uint8_t* c = rgb + (base + k) * C;
        px[k] = static_cast<uint32_t>(c[0]) | (static_cast<uint32_t>(c[1]) << 8) | (static_cast<uint32_t>(c[2]) << 16);
      }
    }
  }
}

__global__ void __launch_bounds__(kPcThreads)
    pc_count_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int C, long long hw, int vec,
                    int rgb_vec, int32_t* tile_count, long long* tile_offset, int tiles, PcWorkspace* ws,
                    long long* n_points) {
  pdl_launch_dependents();
  __shared__ int s_cnt[kPcThreads / 32];
  __shared__ unsigned s_max[kPcThreads / 32];
  __shared__ long long s_scan[kPcThreads / 32];
  __shared__ long long s_base;
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long base = static_cast<long long>(blockIdx.x) * kPcTile + threadIdx.x * kPcPerThread;
  float d[kPcPerThread];
  load4(depth, base, hw, vec, d);
  int cnt = 0;
  unsigned mx = 0;
#pragma unroll
  for (int k = 0; k < kPcPerThread; ++k) cnt += pc_valid(d[k]);
  if (rgb != nullptr && cnt) {
    uint32_t px[kPcPerThread];
    load_rgb4(rgb, C, base, hw, rgb_vec, px);
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k)
      if (pc_valid(d[k])) mx = max(mx, max(px[k] & 255u, max((px[k] >> 8) & 255u, (px[k] >> 16) & 255u)));
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  mx = __reduce_max_sync(0xffffffffu, mx);
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
    __threadfence();  // publish the count before taking the ticket
    s_last = atomicAdd(&ws->done, 1u) == gridDim.x - 1;
    s_base = 0;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA: exclusive scan of every tile's count (all of them are published) ----
  __threadfence();
  for (int t0 = 0; t0 < tiles; t0 += kPcThreads) {
    const int t = t0 + threadIdx.x;
    const long long v = t < tiles ? __ldcg(tile_count + t) : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) s_scan[wid] = inc;
    __syncthreads();
    long long before = s_base + inc - v, chunk = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) {
      if (w < wid) before += s_scan[w];
      chunk += s_scan[w];
    }
    if (t < tiles) tile_offset[t] = before;
    __syncthreads();
    if (threadIdx.x == 0) s_base += chunk;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ws->total = s_base;
    *n_points = s_base;
    ws->done = 0;  // ready for the next call on this workspace
  }
}

__global__ void __launch_bounds__(kPcThreads)
    pc_write_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int C, int W, long long hw, int vec,
                    int rgb_vec, const double* __restrict__ cam, const PcWorkspace* __restrict__ ws,
                    const long long* __restrict__ tile_offset, double* __restrict__ out, long long capacity) {
  pdl_launch_dependents();
  // structure of arrays: 28 bytes per point instead of 48 doubles the resident CTAs per SM, and
  // consecutive ranks hit consecutive banks
  extern __shared__ __align__(16) double stage[];
  double* sx = stage;
  double* sy = stage + kPcTile;
  double* sz = stage + 2 * kPcTile;
  uint32_t* sc = reinterpret_cast<uint32_t*>(stage + 3 * kPcTile);
  __shared__ int s_warp[kPcThreads / 32];
  const long long base = static_cast<long long>(blockIdx.x) * kPcTile + threadIdx.x * kPcPerThread;
  float d[kPcPerThread];
  load4(depth, base, hw, vec, d);   // depth / rgb are inputs of the chain: no need to wait for pass 1 yet
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kPcPerThread; ++k) cnt += pc_valid(d[k]);
  uint32_t px[kPcPerThread] = {0, 0, 0, 0};
  if (rgb != nullptr && cnt) load_rgb4(rgb, C, base, hw, rgb_vec, px);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int before = inc - cnt, tile_total = 0;
#pragma unroll
  for (int w = 0; w < kPcThreads / 32; ++w) {
    if (w < wid) before += s_warp[w];
    tile_total += s_warp[w];
  }

  const double t0 = cam[0], t1 = cam[1], t2 = cam[2];
  const double fx = cam[12], fy = cam[13], cx = cam[14], cy = cam[15];
  int r = before;
#pragma unroll
  for (int k = 0; k < kPcPerThread; ++k) {
    if (!pc_valid(d[k])) continue;
    const long long i = base + k;
    const int v = static_cast<int>(i / W);
    const int u = static_cast<int>(i - static_cast<long long>(v) * W);
    const double zc = static_cast<double>(d[k]);
    const double xc = ((static_cast<double>(u) - cx) * zc) / fx;
    const double yc = ((static_cast<double>(v) - cy) * zc) / fy;
    sx[r] = ((cam[3] * xc + cam[4] * yc) + cam[5] * zc) + t0;
    sy[r] = ((cam[6] * xc + cam[7] * yc) + cam[8] * zc) + t1;
    sz[r] = ((cam[9] * xc + cam[10] * yc) + cam[11] * zc) + t2;
    sc[r] = rgb ? px[k] : 0x00ffffffu;  // no image: white, gcd.py:698-700
    ++r;
  }
  __syncthreads();  // stage complete

  pdl_wait();       // tile offsets and the colour maximum come from pass 1
  // gcd.py:693: rgb.max() <= 1.0 over the valid pixels -> the colours were a [0,1] image: x255
  const bool scale = rgb != nullptr && ws->rgb_max <= 1u;
  // stream the tile's points out: ranks are consecutive, so it is one contiguous run
  const long long first = tile_offset[blockIdx.x];
  long long keep = capacity - first;  // points beyond `capacity` are dropped but were counted
  if (keep > tile_total) keep = tile_total;
  if (keep <= 0) return;
  const double cs = scale ? 255.0 : 1.0;
  const int n2 = static_cast<int>(keep) * 3;  // 16-byte pairs: (x,y) (z,r) (g,b)
  double* dst = out + first * 6;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    double2* d2 = reinterpret_cast<double2*>(dst);
    for (int j = threadIdx.x; j < n2; j += kPcThreads) {
      const int pt = j / 3, m = j - pt * 3;
      const uint32_t c = sc[pt];
      double2 vv;
      if (m == 0) {
        vv.x = sx[pt];
        vv.y = sy[pt];
      } else if (m == 1) {
        vv.x = sz[pt];
        vv.y = static_cast<double>(c & 255u) * cs;
      } else {
        vv.x = static_cast<double>((c >> 8) & 255u) * cs;
        vv.y = static_cast<double>((c >> 16) & 255u) * cs;
      }
      asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(d2 + j), "d"(vv.x), "d"(vv.y) : "memory");
    }
  } else {
    for (int j = threadIdx.x; j < n2 * 2; j += kPcThreads) {
      const int pt = j / 6, m = j - pt * 6;
      dst[j] = m == 0 ? sx[pt] : m == 1 ? sy[pt] : m == 2 ? sz[pt] : static_cast<double>((sc[pt] >> (8 * (m - 3))) & 255u) * cs;
    }
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
  if (hw == 0) {
    CSPE_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(PcWorkspace), st));
    CSPE_CUDA_OK(cudaMemsetAsync(n_points, 0, sizeof(int64_t), st));
    return CSPE_OK;
  }
  CSPE_REQUIRE(depth != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: depth is null");
  CSPE_REQUIRE(capacity == 0 || out != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: out is null");
  const long long tiles = pc_tiles(hw);
  CSPE_REQUIRE(tiles < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_depth_to_pointcloud: frame too large");
  long long* tile_offset = reinterpret_cast<long long*>(ws + 1);
  int32_t* tile_count = reinterpret_cast<int32_t*>(tile_offset + tiles);
  const int vec = (reinterpret_cast<uintptr_t>(depth) & 15) == 0;
  static const cudaError_t smem_attr =
      cudaFuncSetAttribute(pc_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPcStageBytes);
  (void)smem_attr;
  const int rgb_vec = rgb != nullptr && rgb_channels == 4 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0;
  CSPE_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(PcWorkspace), st));
  // plain launch first (serialised behind whatever produced depth / rgb), then the PDL-chained writer
  pc_count_kernel<<<static_cast<unsigned>(tiles), kPcThreads, 0, st>>>(depth, rgb, rgb_channels, hw, vec, rgb_vec, tile_count,
                                                                      tile_offset, static_cast<int>(tiles), ws,
                                                                      reinterpret_cast<long long*>(n_points));
  CSPE_LAUNCH_OK("pc_count_kernel");
  if (capacity > 0)
    CSPE_CUDA_OK(launch_pdl(pc_write_kernel, dim3(static_cast<unsigned>(tiles)), dim3(kPcThreads), kPcStageBytes, st, depth,
                            rgb, rgb_channels, W, hw, vec, rgb_vec, cam, static_cast<const PcWorkspace*>(ws),
                            static_cast<const long long*>(tile_offset), out, static_cast<long long>(capacity)));
  return CSPE_OK;
}
