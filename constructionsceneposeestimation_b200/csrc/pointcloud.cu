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
// both read twice.  A whole BATCH of frames goes through two launches chained by programmatic dependent
// launch (one frame at a time is latency-bound: 2 025 short-lived CTAs per pass never fill the machine, and its
// 64 MB of points stay in L2); a CTA owns kPcGroup consecutive 1024-pixel tiles of one frame and keeps their
// loads in flight together:
//   1. per tile: valid count and the maximum colour over valid pixels; the LAST CTA of each frame (atomic
//      ticket) scans that frame's tile counts into tile offsets, and the last of those scans the frame totals
//      into offsets[0 .. B] — frames are compacted back to back, each in row-major order;
//   2. per tile: recompute validity, in-tile ranks, build the tile's points in shared memory and
//      stream them out as one contiguous run of 16-byte stores (a thread writing its own 48-byte
//      points directly costs ~6x the sector traffic); colours are scaled by 255 when the frame's maximum
//      found by pass 1 is <= 1 (the reference's [0,1]-image rule, gcd.py:693, per frame as it is per call there).
#include <math.h>
#include <stdlib.h>

#include "cspe_common.cuh"

namespace cspe {
namespace {

constexpr int kPcThreads = 256;
constexpr int kPcPerThread = 4;
constexpr int kPcTile = kPcThreads * kPcPerThread;  // 1024 pixels
constexpr int kPcGroupBatch = 4;   // consecutive tiles of one frame per CTA when the batch fills the machine ...
constexpr int kPcGroupSmall = 1;   // ... and one tile per CTA (more, shorter CTAs) when it does not: a single frame
constexpr int kPcStageBytes = kPcTile * (3 * 8 + 4);  // a tile's points as x[], y[], z[] (f64) + packed colour: 28 KB

// Caller-provided scratch, one batch: per-frame headers, then per-(frame, tile) counts and offsets.
struct PcFrameHeader {
  unsigned int rgb_max;  // maximum colour byte over the frame's valid pixels
  unsigned int done;     // CTAs of pass 1 that have published their counts (ticket for the frame's scan)
  long long total;       // valid pixels of the frame
};
struct PcBatchHeader {
  unsigned int frames_done;  // frames whose scan is complete (ticket for the scan over frames)
  unsigned int pad;
};

__host__ __device__ inline size_t pc_align16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }

struct PcLayout {
  PcBatchHeader* batch;
  PcFrameHeader* frame;     // [B]
  int32_t* tile_count;      // [B][tiles]
  int32_t* tile_offset;     // [B][tiles]  first point of the tile, relative to its frame
  size_t header_bytes;      // what a call has to zero
  size_t bytes;
};

__host__ __device__ inline PcLayout pc_layout(void* ws, int B, long long tiles) {
  PcLayout l;
  unsigned char* p = static_cast<unsigned char*>(ws);
  l.batch = reinterpret_cast<PcBatchHeader*>(p);
  size_t off = pc_align16(sizeof(PcBatchHeader));
  l.frame = reinterpret_cast<PcFrameHeader*>(p + off);
  off = pc_align16(off + static_cast<size_t>(B) * sizeof(PcFrameHeader));
  l.header_bytes = off;
  l.tile_count = reinterpret_cast<int32_t*>(p + off);
  off = pc_align16(off + static_cast<size_t>(B) * tiles * 4);
  l.tile_offset = reinterpret_cast<int32_t*>(p + off);
  off = pc_align16(off + static_cast<size_t>(B) * tiles * 4);
  l.bytes = off;
  return l;
}

__device__ __forceinline__ bool pc_valid(float d) {
  return (d > 0.0f) && (d < 250.0f);  // finite follows from < 250; NaN fails both
}

__device__ __forceinline__ void load4(const float* __restrict__ depth, long long base, long long hw, bool vec, float* d) {
  if (vec && base + 3 < hw) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(depth + base));   // L1 bypass: pass 2 reads before its PDL wait
    d[0] = v.x;
    d[1] = v.y;
    d[2] = v.z;
    d[3] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) d[k] = base + k < hw ? __ldcg(depth + base + k) : 0.0f;
  }
}

// colours of the thread's four pixels: one 16-byte load for RGBA, byte loads otherwise
__device__ __forceinline__ void load_rgb4(const uint8_t* __restrict__ rgb, int C, long long base, long long hw, bool vec,
                                          uint32_t* px) {
  if (vec && base + 3 < hw) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(rgb + base * 4));
    px[0] = v.x, px[1] = v.y, px[2] = v.z, px[3] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) {
      px[k] = 0;
      if (base + k < hw) {
        const uint8_t* c = rgb + (base + k) * C;
        px[k] = static_cast<uint32_t>(__ldcg(c)) | (static_cast<uint32_t>(__ldcg(c + 1)) << 8) |
                (static_cast<uint32_t>(__ldcg(c + 2)) << 16);
      }
    }
  }
}

// exclusive scan of n int64 values by one CTA (values read through `get`, results through `put`); returns the total
template <typename Get, typename Put>
__device__ __forceinline__ long long cta_exclusive_scan(int n, long long* s_scan, long long* s_base, Get get, Put put) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) *s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n; t0 += kPcThreads) {
    const int t = t0 + threadIdx.x;
    const long long v = t < n ? get(t) : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long m = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += m;
    }
    if (lane == 31) s_scan[wid] = inc;
    __syncthreads();
    long long before = *s_base + inc - v, chunk = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) {
      if (w < wid) before += s_scan[w];
      chunk += s_scan[w];
    }
    if (t < n) put(t, before);
    __syncthreads();
    if (threadIdx.x == 0) *s_base += chunk;
    __syncthreads();
  }
  return *s_base;
}

// Pass 1.  Grid = B x groups; CTA (f, g) counts the valid pixels of tiles g*kPcGroup .. of frame f and folds the
// frame's colour maximum.  The last CTA of a frame (atomic ticket) scans the frame's tile counts; the last of those
// scans the frame totals into `offsets`.
template <int kPcGroup>
__global__ void __launch_bounds__(kPcThreads)
    pc_count_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int C, long long hw, int vec,
                    int rgb_vec, int B, int tiles, int groups, void* ws_raw, long long* offsets, long long* total_out) {
  pdl_launch_dependents();
  __shared__ int s_cnt[kPcGroup][kPcThreads / 32];
  __shared__ unsigned s_max[kPcThreads / 32];
  __shared__ long long s_scan[kPcThreads / 32];
  __shared__ long long s_base;
  __shared__ bool s_last;
  const PcLayout L = pc_layout(ws_raw, B, tiles);
  const int f = blockIdx.x / groups, g = blockIdx.x - f * groups;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* dep = depth + static_cast<long long>(f) * hw;
  const uint8_t* col = rgb ? rgb + static_cast<long long>(f) * hw * C : nullptr;
  float d[kPcGroup][kPcPerThread];
#pragma unroll
  for (int i = 0; i < kPcGroup; ++i) {
    const long long base = (static_cast<long long>(g) * kPcGroup + i) * kPcTile + threadIdx.x * kPcPerThread;
    if (g * kPcGroup + i < tiles) load4(dep, base, hw, vec, d[i]);
    else d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.0f;
  }
  unsigned mx = 0;
#pragma unroll
  for (int i = 0; i < kPcGroup; ++i) {
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) cnt += pc_valid(d[i][k]);
    if (col != nullptr && cnt) {
      uint32_t px[kPcPerThread];
      const long long base = (static_cast<long long>(g) * kPcGroup + i) * kPcTile + threadIdx.x * kPcPerThread;
      load_rgb4(col, C, base, hw, rgb_vec, px);
#pragma unroll
      for (int k = 0; k < kPcPerThread; ++k)
        if (pc_valid(d[i][k])) mx = max(mx, max(px[k] & 255u, max((px[k] >> 8) & 255u, (px[k] >> 16) & 255u)));
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) s_cnt[i][wid] = cnt;
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) s_max[wid] = mx;
  __syncthreads();
  if (threadIdx.x < kPcGroup && g * kPcGroup + threadIdx.x < tiles) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) t += s_cnt[threadIdx.x][w];
    L.tile_count[static_cast<long long>(f) * tiles + g * kPcGroup + threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) m = max(m, s_max[w]);
    if (m) atomicMax(&L.frame[f].rgb_max, m);
    __threadfence();  // publish the counts before taking the ticket
    s_last = atomicAdd(&L.frame[f].done, 1u) == static_cast<unsigned>(groups) - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA of frame f: exclusive scan of the frame's tile counts (all of them are published) ----
  __threadfence();
  const int32_t* tc = L.tile_count + static_cast<long long>(f) * tiles;
  int32_t* to = L.tile_offset + static_cast<long long>(f) * tiles;
  const long long total = cta_exclusive_scan(
      tiles, s_scan, &s_base, [&](int t) { return static_cast<long long>(__ldcg(tc + t)); },
      [&](int t, long long v) { to[t] = static_cast<int32_t>(v); });
  if (threadIdx.x == 0) {
    L.frame[f].total = total;
    L.frame[f].done = 0;  // ready for the next call on this workspace
    __threadfence();
    s_last = atomicAdd(&L.batch->frames_done, 1u) == static_cast<unsigned>(B) - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last frame to finish: exclusive scan of the frame totals -> offsets[0 .. B] ----
  __threadfence();
  const long long all = cta_exclusive_scan(
      B, s_scan, &s_base, [&](int t) { return *reinterpret_cast<volatile long long*>(&L.frame[t].total); },
      [&](int t, long long v) { offsets[t] = v; });
  if (threadIdx.x == 0) {
    offsets[B] = all;
    if (total_out) *total_out = all;   // the one-frame entry point's n_points
    L.batch->frames_done = 0;
  }
}

// ---- bulk (TMA) stores -------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {   // at most kPending groups still READING shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kPcBulkBufBytes = kPcTile * 48;          // a tile's points as the (N, 6) float64 rows they become: 48 KB
// kBufs = 2: the bulk store of tile i drains under tile i+1 of the same CTA (2 CTAs per SM);
// kBufs = 1: one buffer, more CTAs per SM overlap each other instead

// Pass 2.  Same grid; every tile's points are built in shared memory exactly as they lie in `out` — rows of six
// doubles, written as three conflict-free 16-byte stores per point — and leave with ONE bulk copy
// (cp.async.bulk shared -> global, SASS UBLKCP) per tile: no thread spends instructions on the copy, and the store of
// tile i drains while the CTA projects tile i+1 into the other buffer.
template <int kPcGroup, int kBufs>
__global__ void __launch_bounds__(kPcThreads, kBufs == 1 ? 4 : 2)
    pc_write_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int C, int W, long long hw, int vec,
                    int rgb_vec, const double* __restrict__ cam_all, int B, int tiles, int groups, void* ws_raw,
                    const long long* __restrict__ offsets, double* __restrict__ out, long long capacity) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char pc_smem[];
  __shared__ int s_warp[kPcGroup][kPcThreads / 32];
  __shared__ long long s_first[kPcGroup];
  const PcLayout L = pc_layout(ws_raw, B, tiles);
  const int f = blockIdx.x / groups, g = blockIdx.x - f * groups;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* dep = depth + static_cast<long long>(f) * hw;
  const uint8_t* col = rgb ? rgb + static_cast<long long>(f) * hw * C : nullptr;
  const double* cam = cam_all + static_cast<long long>(f) * CSPE_CAM_STRIDE;
  // the camera block is read before the PDL wait: L1 bypass; issued first so it travels with the pixel loads
  double cm[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) cm[j] = __ldcg(cam + j);
  // depth / rgb are inputs of the chain: everything up to the first store runs before waiting for pass 1
  float d[kPcGroup][kPcPerThread];
  uint32_t px[kPcGroup][kPcPerThread];
  int cnt[kPcGroup], inc[kPcGroup];
#pragma unroll
  for (int i = 0; i < kPcGroup; ++i) {
    const long long base = (static_cast<long long>(g) * kPcGroup + i) * kPcTile + threadIdx.x * kPcPerThread;
    if (g * kPcGroup + i < tiles) load4(dep, base, hw, vec, d[i]);
    else d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.0f;
  }
#pragma unroll
  for (int i = 0; i < kPcGroup; ++i) {
    cnt[i] = 0;
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) cnt[i] += pc_valid(d[i][k]);
    px[i][0] = px[i][1] = px[i][2] = px[i][3] = 0;
    if (col != nullptr && cnt[i]) {
      const long long base = (static_cast<long long>(g) * kPcGroup + i) * kPcTile + threadIdx.x * kPcPerThread;
      load_rgb4(col, C, base, hw, rgb_vec, px[i]);
    }
    int v = cnt[i];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    inc[i] = v;
    if (lane == 31) s_warp[i][wid] = v;
  }
  __syncthreads();

  const double t0 = cm[0], t1 = cm[1], t2 = cm[2];
  const double fx = cm[12], fy = cm[13], cx = cm[14], cy = cm[15];
  // pass 1's results (tile offsets, frame offsets, the frame's colour maximum) are needed from the first point on:
  // the colour scale goes into the staged rows.  Loads, counts and prefix sums above ran before this wait.
  pdl_wait();
  // gcd.py:693: rgb.max() <= 1.0 over the frame's valid pixels -> the colours were a [0,1] image: x255
  // (written while this kernel was resident: explicit L2 loads, cspe_common.cuh PDL rule)
  const unsigned cmul = (col != nullptr && __ldcg(&L.frame[f].rgb_max) <= 1u) ? 255u : 1u;
  const long long frame_first = __ldcg(offsets + f);
  // the group's tile offsets in one round trip (thread 0 needs them one by one at every bulk store)
  if (threadIdx.x < kPcGroup && g * kPcGroup + threadIdx.x < tiles)
    s_first[threadIdx.x] = frame_first + __ldcg(L.tile_offset + static_cast<long long>(f) * tiles + g * kPcGroup + threadIdx.x);
#pragma unroll   // static indices keep d / px / cnt / inc in registers
  for (int i = 0; i < kPcGroup; ++i) {
    const int tile = g * kPcGroup + i;
    if (tile >= tiles) break;
    double2* buf = reinterpret_cast<double2*>(pc_smem + (i % kBufs) * kPcBulkBufBytes);
    if (i >= kBufs) {  // the bulk store that used this buffer last has to be done READING it
      if (threadIdx.x == 0) bulk_wait_read<kBufs - 1>();
      __syncthreads();
    }
    int before = inc[i] - cnt[i], tile_total = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) {
      const int s = s_warp[i][w];
      if (w < wid) before += s;
      tile_total += s;
    }
    // pixel index of the thread's first pixel (a frame has < 2^31 pixels): one 32-bit division per tile, the
    // other three pixels follow by stepping (u, v)
    const unsigned p0 = static_cast<unsigned>(tile) * kPcTile + threadIdx.x * kPcPerThread;
    int v = static_cast<int>(p0 / static_cast<unsigned>(W));
    int u = static_cast<int>(p0 - static_cast<unsigned>(v) * static_cast<unsigned>(W));
    int r = before;
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) {
      if (pc_valid(d[i][k])) {
        const double zc = static_cast<double>(d[i][k]);
        const double xc = ((static_cast<double>(u) - cx) * zc) / fx;
        const double yc = ((static_cast<double>(v) - cy) * zc) / fy;
        const uint32_t c = col ? px[i][k] : 0x00ffffffu;  // no image: white, gcd.py:698-700
        double2 a, b2, c2;
        a.x = ((cm[3] * xc + cm[4] * yc) + cm[5] * zc) + t0;
        a.y = ((cm[6] * xc + cm[7] * yc) + cm[8] * zc) + t1;
        b2.x = ((cm[9] * xc + cm[10] * yc) + cm[11] * zc) + t2;
        b2.y = static_cast<double>((c & 255u) * cmul);
        c2.x = static_cast<double>(((c >> 8) & 255u) * cmul);
        c2.y = static_cast<double>(((c >> 16) & 255u) * cmul);
        // 48-byte rows: the eight lanes of a quarter-warp cover all 32 banks exactly once per 16-byte store
        buf[r * 3 + 0] = a;
        buf[r * 3 + 1] = b2;
        buf[r * 3 + 2] = c2;
        ++r;
      }
      if (++u == W) {
        u = 0;
        ++v;
      }
    }
    fence_async_smem();  // the rows were written through the generic proxy, the bulk copy reads through the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long first = s_first[i];   // written before the __syncthreads above
      long long keep = capacity - first;  // points beyond `capacity` are dropped but were counted
      if (keep > tile_total) keep = tile_total;
      if (keep > 0) bulk_store_s2g(out + first * 6, buf, static_cast<uint32_t>(keep) * 48u);
      bulk_commit();   // an empty group keeps the wait_group arithmetic uniform
    }
  }
  if (threadIdx.x == 0) bulk_wait_read<0>();  // shared memory must outlive the copies that read it
}

constexpr int kPcPersistentCtasPerSm = 3;   // 85 registers: the camera block alone is 32 (4 CTAs at 64 registers spill)

// Pass 2, persistent form (the default for batches): 3 CTAs per SM, each walking a contiguous range of (frame, tile)
// pairs.  The loads of tile t+1 are issued BEFORE tile t is projected, so their round trip hides under the
// math (the per-CTA form above spends 41 % of its warp time waiting for its one burst of loads:
// profiles/r02_pc_write_ncu_summary.txt), and the bulk store of tile t drains while tile t+1 is counted.
__global__ void __launch_bounds__(kPcThreads, kPcPersistentCtasPerSm)
    pc_write_persistent_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int C, int W, long long hw,
                               int vec, int rgb_vec, const double* __restrict__ cam_all, int B, int tiles, void* ws_raw,
                               const long long* __restrict__ offsets, double* __restrict__ out, long long capacity) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char pc_smem[];
  __shared__ int s_warp[2][kPcThreads / 32];   // per-warp counts, double-buffered by tile parity
  const PcLayout L = pc_layout(ws_raw, B, tiles);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long total = static_cast<long long>(B) * tiles;
  const long long t_begin = total * blockIdx.x / gridDim.x, t_end = total * (blockIdx.x + 1) / gridDim.x;
  if (t_begin >= t_end) return;
  double2* buf = reinterpret_cast<double2*>(pc_smem);

  auto load_tile = [&](long long t, float* d, uint32_t* px) {
    const int f = static_cast<int>(t / tiles), tile = static_cast<int>(t - static_cast<long long>(f) * tiles);
    const long long base = static_cast<long long>(tile) * kPcTile + threadIdx.x * kPcPerThread;
    load4(depth + static_cast<long long>(f) * hw, base, hw, vec, d);
    px[0] = px[1] = px[2] = px[3] = 0;
    if (rgb != nullptr) load_rgb4(rgb + static_cast<long long>(f) * hw * C, C, base, hw, rgb_vec, px);
  };

  float d[kPcPerThread], dn[kPcPerThread];
  uint32_t px[kPcPerThread], pxn[kPcPerThread];
  load_tile(t_begin, dn, pxn);   // inputs of the chain: in flight before the PDL wait

  int cur_f = -1;
  double cm[16];
  double t0 = 0, t1 = 0, t2 = 0, fx = 1, fy = 1, cx = 0, cy = 0;
  unsigned cmul = 1u;
  long long frame_first = 0;
  pdl_wait();   // pass 1's tile offsets, frame offsets and colour maxima (read below through L2)
  int it = 0;
  for (long long t = t_begin; t < t_end; ++t, ++it) {
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) {
      d[k] = dn[k];
      px[k] = pxn[k];
    }
    if (t + 1 < t_end) load_tile(t + 1, dn, pxn);   // next tile's round trip hides under this tile's math
    const int f = static_cast<int>(t / tiles), tile = static_cast<int>(t - static_cast<long long>(f) * tiles);
    if (f != cur_f) {   // a CTA's range crosses at most a few frame boundaries
      cur_f = f;
#pragma unroll
      for (int j = 0; j < 16; ++j) cm[j] = __ldcg(cam_all + static_cast<long long>(f) * CSPE_CAM_STRIDE + j);
      t0 = cm[0], t1 = cm[1], t2 = cm[2];
      fx = cm[12], fy = cm[13], cx = cm[14], cy = cm[15];
      cmul = (rgb != nullptr && __ldcg(&L.frame[f].rgb_max) <= 1u) ? 255u : 1u;   // gcd.py:693, per frame
      frame_first = __ldcg(offsets + f);
    }
    long long first = 0;
    if (threadIdx.x == 0) first = frame_first + __ldcg(L.tile_offset + static_cast<long long>(f) * tiles + tile);
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) cnt += pc_valid(d[k]);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[it & 1][wid] = inc;
    if (threadIdx.x == 0 && it > 0) bulk_wait_read<0>();   // the previous tile's bulk store is done reading `buf`
    __syncthreads();
    int before = inc - cnt, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) {
      const int sw = s_warp[it & 1][w];
      if (w < wid) before += sw;
      tile_total += sw;
    }
    const unsigned p0 = static_cast<unsigned>(tile) * kPcTile + threadIdx.x * kPcPerThread;
    int v = static_cast<int>(p0 / static_cast<unsigned>(W));
    int u = static_cast<int>(p0 - static_cast<unsigned>(v) * static_cast<unsigned>(W));
    int r = before;
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) {
      if (pc_valid(d[k])) {
        const double zc = static_cast<double>(d[k]);
        const double xc = ((static_cast<double>(u) - cx) * zc) / fx;
        const double yc = ((static_cast<double>(v) - cy) * zc) / fy;
        const uint32_t c = rgb ? px[k] : 0x00ffffffu;  // no image: white, gcd.py:698-700
        double2 a, b2, c2;
        a.x = ((cm[3] * xc + cm[4] * yc) + cm[5] * zc) + t0;
        a.y = ((cm[6] * xc + cm[7] * yc) + cm[8] * zc) + t1;
        b2.x = ((cm[9] * xc + cm[10] * yc) + cm[11] * zc) + t2;
        b2.y = static_cast<double>((c & 255u) * cmul);
        c2.x = static_cast<double>(((c >> 8) & 255u) * cmul);
        c2.y = static_cast<double>(((c >> 16) & 255u) * cmul);
        buf[r * 3 + 0] = a;
        buf[r * 3 + 1] = b2;
        buf[r * 3 + 2] = c2;
        ++r;
      }
      if (++u == W) {
        u = 0;
        ++v;
      }
    }
    fence_async_smem();  // rows written through the generic proxy, read by the bulk copy through the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      long long keep = capacity - first;  // points beyond `capacity` are dropped but were counted
      if (keep > tile_total) keep = tile_total;
      if (keep > 0) bulk_store_s2g(out + first * 6, buf, static_cast<uint32_t>(keep) * 48u);
      bulk_commit();
    }
  }
  if (threadIdx.x == 0) bulk_wait_read<0>();  // shared memory must outlive the copies that read it
}

// Pass 2, fallback for an `out` that is not 16-byte aligned: the tile's points are staged as x[] y[] z[] + packed colour
// and copied out by the threads themselves.
template <int kPcGroup>
__global__ void __launch_bounds__(kPcThreads)
    pc_write_loop_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int C, int W, long long hw, int vec,
                    int rgb_vec, const double* __restrict__ cam_all, int B, int tiles, int groups, void* ws_raw,
                    const long long* __restrict__ offsets, double* __restrict__ out, long long capacity) {
  pdl_launch_dependents();
  // structure of arrays: 28 bytes per point instead of 48 doubles the resident CTAs per SM, and
  // consecutive ranks hit consecutive banks
  extern __shared__ __align__(16) double stage[];
  double* sx = stage;
  double* sy = stage + kPcTile;
  double* sz = stage + 2 * kPcTile;
  uint32_t* sc = reinterpret_cast<uint32_t*>(stage + 3 * kPcTile);
  __shared__ int s_warp[kPcGroup][kPcThreads / 32];
  const PcLayout L = pc_layout(ws_raw, B, tiles);
  const int f = blockIdx.x / groups, g = blockIdx.x - f * groups;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* dep = depth + static_cast<long long>(f) * hw;
  const uint8_t* col = rgb ? rgb + static_cast<long long>(f) * hw * C : nullptr;
  const double* cam = cam_all + static_cast<long long>(f) * CSPE_CAM_STRIDE;
  // depth / rgb are inputs of the chain: everything up to the first store runs before waiting for pass 1
  float d[kPcGroup][kPcPerThread];
  uint32_t px[kPcGroup][kPcPerThread];
  int cnt[kPcGroup], inc[kPcGroup];
#pragma unroll
  for (int i = 0; i < kPcGroup; ++i) {
    const long long base = (static_cast<long long>(g) * kPcGroup + i) * kPcTile + threadIdx.x * kPcPerThread;
    if (g * kPcGroup + i < tiles) load4(dep, base, hw, vec, d[i]);
    else d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.0f;
  }
#pragma unroll
  for (int i = 0; i < kPcGroup; ++i) {
    cnt[i] = 0;
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) cnt[i] += pc_valid(d[i][k]);
    px[i][0] = px[i][1] = px[i][2] = px[i][3] = 0;
    if (col != nullptr && cnt[i]) {
      const long long base = (static_cast<long long>(g) * kPcGroup + i) * kPcTile + threadIdx.x * kPcPerThread;
      load_rgb4(col, C, base, hw, rgb_vec, px[i]);
    }
    int v = cnt[i];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    inc[i] = v;
    if (lane == 31) s_warp[i][wid] = v;
  }
  __syncthreads();

  // the camera block is read before the PDL wait as well: L1 bypass
  double cm[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) cm[j] = __ldcg(cam + j);
  const double t0 = cm[0], t1 = cm[1], t2 = cm[2];
  const double fx = cm[12], fy = cm[13], cx = cm[14], cy = cm[15];
  bool waited = false;
  bool scale = false;
  long long frame_first = 0;
#pragma unroll   // static indices keep d / px / cnt / inc in registers
  for (int i = 0; i < kPcGroup; ++i) {
    const int tile = g * kPcGroup + i;
    if (tile >= tiles) break;
    int before = inc[i] - cnt[i], tile_total = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) {
      const int s = s_warp[i][w];
      if (w < wid) before += s;
      tile_total += s;
    }
    // pixel index of the thread's first pixel (a frame has < 2^31 pixels): one 32-bit division per tile, the
    // other three pixels follow by stepping (u, v)
    const unsigned p0 = static_cast<unsigned>(tile) * kPcTile + threadIdx.x * kPcPerThread;
    int v = static_cast<int>(p0 / static_cast<unsigned>(W));
    int u = static_cast<int>(p0 - static_cast<unsigned>(v) * static_cast<unsigned>(W));
    int r = before;
#pragma unroll
    for (int k = 0; k < kPcPerThread; ++k) {
      if (pc_valid(d[i][k])) {
        const double zc = static_cast<double>(d[i][k]);
        const double xc = ((static_cast<double>(u) - cx) * zc) / fx;
        const double yc = ((static_cast<double>(v) - cy) * zc) / fy;
        sx[r] = ((cm[3] * xc + cm[4] * yc) + cm[5] * zc) + t0;
        sy[r] = ((cm[6] * xc + cm[7] * yc) + cm[8] * zc) + t1;
        sz[r] = ((cm[9] * xc + cm[10] * yc) + cm[11] * zc) + t2;
        sc[r] = col ? px[i][k] : 0x00ffffffu;  // no image: white, gcd.py:698-700
        ++r;
      }
      if (++u == W) {
        u = 0;
        ++v;
      }
    }
    __syncthreads();  // stage complete
    if (!waited) {
      pdl_wait();  // tile offsets, frame offsets and the colour maximum come from pass 1
      waited = true;
      // gcd.py:693: rgb.max() <= 1.0 over the frame's valid pixels -> the colours were a [0,1] image: x255
      // (written while this kernel was resident: explicit L2 loads, cspe_common.cuh PDL rule)
      scale = col != nullptr && __ldcg(&L.frame[f].rgb_max) <= 1u;
      frame_first = __ldcg(offsets + f);
    }
    // stream the tile's points out: ranks are consecutive, so it is one contiguous run
    const long long first = frame_first + __ldcg(L.tile_offset + static_cast<long long>(f) * tiles + tile);
    long long keep = capacity - first;  // points beyond `capacity` are dropped but were counted
    if (keep > tile_total) keep = tile_total;
    if (keep > 0) {
      const int n2 = static_cast<int>(keep) * 3;  // 16-byte pairs: (x,y) (z,r) (g,b)
      double* dst = out + first * 6;
      const unsigned cm = scale ? 255u : 1u;   // [0,1] image: 0 / 1 -> 0 / 255, exact in integers
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
            vv.y = static_cast<double>((c & 255u) * cm);
          } else {
            vv.x = static_cast<double>(((c >> 8) & 255u) * cm);
            vv.y = static_cast<double>(((c >> 16) & 255u) * cm);
          }
          asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(d2 + j), "d"(vv.x), "d"(vv.y) : "memory");
        }
      } else {
        for (int j = threadIdx.x; j < n2 * 2; j += kPcThreads) {
          const int pt = j / 6, m = j - pt * 6;
          dst[j] = m == 0 ? sx[pt] : m == 1 ? sy[pt] : m == 2 ? sz[pt]
                                                              : static_cast<double>(((sc[pt] >> (8 * (m - 3))) & 255u) * cm);
        }
      }
    }
    __syncthreads();  // the stage is rewritten by the next tile
  }
}

}  // namespace
}  // namespace cspe

using namespace cspe;

static long long pc_tiles(long long hw) { return (hw + kPcTile - 1) / kPcTile; }

extern "C" size_t cspe_pointcloud_batch_workspace_bytes(int B, int H, int W) {
  if (B <= 0) B = 1;
  const long long tiles = (H <= 0 || W <= 0) ? 0 : pc_tiles(static_cast<long long>(H) * W);
  return pc_layout(nullptr, B, tiles).bytes + 16;
}

extern "C" size_t cspe_pointcloud_workspace_bytes(int H, int W) { return cspe_pointcloud_batch_workspace_bytes(1, H, W); }

static int pointcloud_batch_impl(const float* depth, const uint8_t* rgb, int rgb_channels, int B, int H, int W,
                                 const double* cam, double* out, int64_t capacity, int64_t* offsets, void* workspace,
                                 void* stream, int64_t* total_out) {
  CSPE_REQUIRE(B >= 0 && H >= 0 && W >= 0 && capacity >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_depth_to_pointcloud: negative size");
  if (B == 0) return CSPE_OK;
  CSPE_REQUIRE(offsets && workspace && cam, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: null pointer");
  CSPE_REQUIRE(rgb == nullptr || rgb_channels >= 3, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_depth_to_pointcloud: rgb needs >= 3 channels (got %d)", rgb_channels);
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(offsets) & 7) == 0,
               CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: workspace must be 16-byte, out / offsets 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long hw = static_cast<long long>(H) * W;
  if (hw == 0) {
    CSPE_CUDA_OK(cudaMemsetAsync(offsets, 0, sizeof(int64_t) * (static_cast<size_t>(B) + 1), st));
    if (total_out) CSPE_CUDA_OK(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    return CSPE_OK;
  }
  CSPE_REQUIRE(depth != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: depth is null");
  CSPE_REQUIRE(capacity == 0 || out != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: out is null");
  CSPE_REQUIRE(hw < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_depth_to_pointcloud: frame too large");
  const long long tiles = pc_tiles(hw);
  const bool small = tiles * B < 8192;   // not enough 4-tile CTAs to fill 148 SMs several times over
  const int group0 = small ? kPcGroupSmall : ((getenv("CSPE_PC_VARIANT") && atoi(getenv("CSPE_PC_VARIANT")) == 2) ? 2 : kPcGroupBatch);
  const long long groups = (tiles + group0 - 1) / group0;
  CSPE_REQUIRE(groups * B < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_depth_to_pointcloud: batch too large");
  const PcLayout L = pc_layout(workspace, B, tiles);
  // frames of a batch start at multiples of hw floats: vector loads need every frame base 16-byte aligned
  const int vec = (reinterpret_cast<uintptr_t>(depth) & 15) == 0 && (B == 1 || hw % 4 == 0);
  const int rgb_vec = rgb != nullptr && rgb_channels == 4 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0 &&
                      (B == 1 || hw % 4 == 0);
  // bulk (TMA) stores need a 16-byte aligned destination; every tile starts at a multiple of 48 bytes from `out`
  const bool bulk = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  auto count_k = small ? pc_count_kernel<kPcGroupSmall> : (group0 == 2 ? pc_count_kernel<2> : pc_count_kernel<kPcGroupBatch>);
  // CSPE_PC_VARIANT (A/B runs): 0 = 4 tiles per CTA, 2 shared-memory buffers (2 CTAs per SM: 1.50 ms for 64 x 1080p RGBA),
  // 1 (default) = 4 tiles, 1 buffer, 64 registers (4 CTAs per SM: 1.33 ms), 2 = 2 tiles, 1 buffer (2.27 ms: the fixed
  // cost of a CTA — launch, load round trip, tickets — is what a CTA has to amortise)
  static const int variant = []() {
    const char* e = getenv("CSPE_PC_VARIANT");
    return e ? atoi(e) : 3;   // 3 = the persistent, software-pipelined form
  }();
  const int group = small ? kPcGroupSmall : (variant == 2 ? 2 : kPcGroupBatch);
  const int bufs = (small || variant >= 1) ? 1 : 2;
  void (*write_k)(const float*, const uint8_t*, int, int, long long, int, int, const double*, int, int, int, void*,
                  const long long*, double*, long long);
  // a single frame is latency-bound: the loop form projects its tile BEFORE it waits for pass 1 and keeps 28 KB of
  // shared memory per CTA (the bulk forms need the frame's colour scale first: 0.086 ms against 0.04 ms for one 1080p frame)
  if (!bulk || small) write_k = small ? pc_write_loop_kernel<kPcGroupSmall> : (variant == 2 ? pc_write_loop_kernel<2> : pc_write_loop_kernel<kPcGroupBatch>);
  else if (variant == 2) write_k = pc_write_kernel<2, 1>;
  else if (variant == 1) write_k = pc_write_kernel<kPcGroupBatch, 1>;
  else write_k = pc_write_kernel<kPcGroupBatch, 2>;
  const size_t smem = static_cast<size_t>((bulk && !small) ? bufs * kPcBulkBufBytes : kPcStageBytes);
  // dynamic shared-memory limits, once per process (the call costs tens of microseconds of host time)
  static const cudaError_t attrs = []() {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto k, int bytes) {
      const cudaError_t r = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e == cudaSuccess) e = r;
    };
    set(pc_write_kernel<kPcGroupSmall, 1>, kPcBulkBufBytes);
    set(pc_write_kernel<2, 1>, kPcBulkBufBytes);
    set(pc_write_kernel<kPcGroupBatch, 1>, kPcBulkBufBytes);
    set(pc_write_kernel<kPcGroupBatch, 2>, 2 * kPcBulkBufBytes);
    set(pc_write_loop_kernel<kPcGroupSmall>, kPcStageBytes);
    set(pc_write_loop_kernel<2>, kPcStageBytes);
    set(pc_write_loop_kernel<kPcGroupBatch>, kPcStageBytes);
    set(pc_write_persistent_kernel, kPcBulkBufBytes);
    return e;
  }();
  CSPE_CUDA_OK(attrs);
  CSPE_CUDA_OK(cudaMemsetAsync(workspace, 0, L.header_bytes, st));
  const unsigned grid = static_cast<unsigned>(groups * B);
  // plain launch first (serialised behind whatever produced depth / rgb), then the PDL-chained writer
  count_k<<<grid, kPcThreads, 0, st>>>(depth, rgb, rgb_channels, hw, vec, rgb_vec, B, static_cast<int>(tiles),
                                      static_cast<int>(groups), workspace, reinterpret_cast<long long*>(offsets),
                                      reinterpret_cast<long long*>(total_out));
  CSPE_LAUNCH_OK("pc_count_kernel");
  if (capacity > 0 && bulk && !small && variant == 3) {
    const long long all_tiles = tiles * B;
    const long long ctas = static_cast<long long>(sm_count()) * kPcPersistentCtasPerSm;
    CSPE_CUDA_OK(launch_pdl(pc_write_persistent_kernel, dim3(static_cast<unsigned>(all_tiles < ctas ? all_tiles : ctas)),
                            dim3(kPcThreads), static_cast<size_t>(kPcBulkBufBytes), st, depth, rgb, rgb_channels, W, hw, vec,
                            rgb_vec, cam, B, static_cast<int>(tiles), workspace,
                            static_cast<const long long*>(reinterpret_cast<long long*>(offsets)), out,
                            static_cast<long long>(capacity)));
  } else if (capacity > 0) {
    CSPE_CUDA_OK(launch_pdl(write_k, dim3(grid), dim3(kPcThreads), smem, st, depth, rgb, rgb_channels, W, hw, vec, rgb_vec,
                            cam, B, static_cast<int>(tiles), static_cast<int>(groups), workspace,
                            static_cast<const long long*>(reinterpret_cast<long long*>(offsets)), out,
                            static_cast<long long>(capacity)));
  }
  return CSPE_OK;
}

extern "C" int cspe_depth_to_pointcloud_batch(const float* depth, const uint8_t* rgb, int rgb_channels, int B, int H, int W,
                                              const double* cam, double* out, int64_t capacity, int64_t* offsets,
                                              void* workspace, void* stream) {
  return pointcloud_batch_impl(depth, rgb, rgb_channels, B, H, W, cam, out, capacity, offsets, workspace, stream, nullptr);
}

// one frame: offsets[1] doubles as the point count (n_points = &pair[1] of a caller-side int64[2] is not required —
// the single-frame entry point keeps its one-value result by scanning into a two-element scratch in the workspace)
extern "C" int cspe_depth_to_pointcloud(const float* depth, const uint8_t* rgb, int rgb_channels, int H, int W,
                                        const double* cam, double* out, int64_t capacity, int64_t* n_points,
                                        void* workspace, void* stream) {
  CSPE_REQUIRE(n_points != nullptr && workspace != nullptr, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_depth_to_pointcloud: null pointer");
  CSPE_REQUIRE(H >= 0 && W >= 0, CSPE_ERR_INVALID_ARGUMENT, "cspe_depth_to_pointcloud: negative size");
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_depth_to_pointcloud: workspace must be 16-byte aligned");
  // offsets[0..1] live at the end of the workspace; the total is copied to n_points on the stream
  const long long tiles = (H <= 0 || W <= 0) ? 0 : pc_tiles(static_cast<long long>(H) * W);
  int64_t* pair = reinterpret_cast<int64_t*>(static_cast<unsigned char*>(workspace) + pc_layout(nullptr, 1, tiles).bytes);
  // offsets[0..1] live in the workspace; the count kernel writes the total to n_points as well
  return pointcloud_batch_impl(depth, rgb, rgb_channels, 1, H, W, cam, out, capacity, pair, workspace, stream, n_points);
}
