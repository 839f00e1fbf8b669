// K1 — single-pass instance-ID mask scan (fills the hole the reference
// leaves at gcd.py:1908-1910, where the instance mask is a -1 placeholder).
//
// For every (frame, slot): pixel count and inclusive x/y extents of all pixels whose
// instance id maps to that slot.  HBM-bound: the u32 mask is read exactly once (4*H*W bytes
// per frame); everything else stays on chip.
//
// Design (B200 / sm_100a):
//   * persistent grid, two CTAs per SM (296 on a B200), each warp-specialised: 1 producer warp +
//     8 consumer warps, capped at 80 registers so that small kernels can run beside it;
//   * the batch is one 2-D tensor [B*H rows][W] behind a TMA tensor map; a tile is 4 boxes of
//     32 px x 64 rows (128-byte rows, SWIZZLE_128B) = 128 px x 64 rows = 32 KB, streamed into a
//     3-stage shared-memory ring by cp.async.bulk.tensor (SASS UTMALDG) signalled on mbarriers,
//     L2 evict_first — one 3-D box {32 px, 64 rows, 4 strips} per tile when W % 32 == 0 (the mask
//     seen as [W/32 strips][B*H rows][32 px]), else one 2-D box per strip.  (Measured: tensor-map boxes stream at 6.4-6.9 TB/s whatever the
//     geometry; one 1-D bulk copy per row segment capped at 4.3 TB/s because every bulk op costs
//     ~85 cycles of TMA.)  Two CTAs per SM = two independent rings per SM: a slow warp only holds
//     back its own ring;
//   * consumer warp w owns box (w & 3) and row half (w >> 2) of the tile: a 32-pixel column
//     strip whose LANES RUN DOWN THE ROWS (lane l = row l).  Fast path (the strip is one id):
//     8 LDS.128 + 16 LOP3.  Slow path: all lanes build a 32-bit run-start mask in lockstep, then
//     walk their runs (one shared-memory load per run).  The 128-byte swizzle makes every
//     LDS.128 bank-conflict free.  A lanes-along-x mapping left 10.8 of 32 lanes active, because
//     every warp meets an object edge;
//   * each thread keeps the two most recent ids with their partial {count, xmin, xmax, ymin,
//     ymax} in registers and touches shared memory only when a third id shows up; evicted
//     entries are reduced across the warp (match.all + redux), then merged into a per-CTA
//     shared-memory table with red.shared add/min/max through a shared-memory copy of the frame's
//     id->slot LUT; the table merges into global memory with red.global once per (CTA, frame);
//   * work is a contiguous range of passes per CTA, but the pass order inside a frame is
//     stride-permuted so every CTA samples busy and empty image regions alike (a plain
//     contiguous split left SMs idle 43 % of the time on instance-dense frames);
//   * programmatic dependent launch: the kernel releases its dependents right after set-up (K2 /
//     K3 then run beside it) and waits for the kernel before it either at entry or — overlapped
//     mode — only before its first merge into `out`, so it can stream while the previous batch's
//     K4 is still reading that table;
//   * masks whose rows are not 16-byte aligned (W % 4 != 0 or a misaligned base) cannot use
//     TMA: the producer warp then fills the same swizzled layout with plain loads.
//
// Integer only, order independent => bit-exact against the numpy oracle.
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "cspe_common.cuh"

namespace cspe {
namespace {

constexpr int kStripPx = 32;                      // one TMA box row = 128 B = 8 x 16 B
constexpr int kBoxRows = 64;
constexpr int kBoxBytes = kStripPx * 4 * kBoxRows;     // 8192
constexpr int kTileRows = kBoxRows;               // 64
constexpr int kStages = 3;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kDefaultBoxes = 4;  // measured best: 2 CTAs per SM (profiles/)
constexpr int kRunWalkMax = 4;    // rows with fewer run STARTS than this walk their runs (<= 4 runs)
constexpr int kMaxDistinct = 3;   // ids accounted per equality mask before the run walk takes the rest

// Geometry of one CTA: kBoxes TMA boxes side by side per tile, two consumer warps per box (row
// halves) + one producer warp.  CTAs per SM = 8 / kBoxes, so an SM always runs 16 consumer warps
// over 3 x 64 KB of ring.  kBoxes = 4 (2 CTAs per SM) measured best; 8 and 2 stay selectable with
// CSPE_SCAN_BOXES for tuning on other parts.
template <int kBoxes>
struct Geo {
  static constexpr int kConsumerWarps = 2 * kBoxes;
  static constexpr int kConsumers = kConsumerWarps * 32;
  static constexpr int kThreads = kConsumers + 32;
  static constexpr int kCtasPerSm = 8 / kBoxes;
  static constexpr int kTileCols = kBoxes * kStripPx;
  static constexpr int kStageBytes = kBoxes * kBoxBytes;
  static constexpr int kBarBytes = 2 * kStages * 8;
  static constexpr int kSmemFixed = kStages * kStageBytes + kBarBytes + 80;
  // what is left of the SM's shared memory for this CTA's slot table and LUT copy (1 KB per CTA
  // is reserved by the driver, and the dynamic base is aligned up to 1 KB)
  static constexpr int kSmemFree = kSmemLimit / kCtasPerSm - kSmemFixed - 2048 * (kCtasPerSm > 1);
};

struct ScanParams {
  const uint32_t* mask;
  const int32_t* lut;
  int32_t* out;
  long long lut_stride;
  long long total_passes;
  int B, H, W, N, lut_len;
  int nseg;      // tile-wide column segments per row
  int nrb;       // 64-row blocks per frame
  int ppf;       // passes per frame = nseg * nrb
  int stride;    // pass permutation inside a frame: q -> (q * stride) % ppf, gcd(stride, ppf) = 1
  int tma;       // 2: one 3-D TMA box per tile (W % 32 == 0), 1: one 2-D TMA box per strip (W % 4 == 0, 16 B
                 // aligned base), 0: producer-warp copy
  int smem_lut;  // 1: the frame's LUT is staged in shared memory
  int lazy_wait; // 1: wait for the previous kernel only before the first merge into `out` (overlapped mode)
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

struct Entry {
  uint32_t id;
  int cnt, xmn, xmx, ymn, ymx;
};

__device__ __forceinline__ void entry_reset(Entry& e, uint32_t id) {
  e.id = id;
  e.cnt = 0;
  e.xmn = INT_MAX;
  e.xmx = -1;
  e.ymn = INT_MAX;
  e.ymx = -1;
}

__device__ __forceinline__ void red_shared_add(int32_t* p, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_shared_min(int32_t* p, int v) {
  asm volatile("red.shared.min.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_shared_max(int32_t* p, int v) {
  asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_global_add(int32_t* p, int v) {
  asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_global_min(int32_t* p, int v) {
  asm volatile("red.global.min.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_global_max(int32_t* p, int v) {
  asm volatile("red.global.max.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Merge one evicted register entry into the CTA table (shared) or straight into `out`.
//
// Lanes run down the rows of one strip, so a whole warp usually evicts the SAME id at the same moment:
// match.all finds that case, redux reduces across the lanes and one lane merges.  Where an object edge crosses
// the 32 rows the lanes evict two or three ids and every lane merges on its own: 32 lanes x 5 reds on a few
// addresses, which the shared-memory pipe serialises (16x16-pixel id blocks: 21.7 M conflict wavefronts, the pipe
// 61 % busy).  The reds are fire-and-forget, so the warp itself does not wait for them — which is why both
// attempts to aggregate them first measured SLOWER on every mask content (profiles/r02_scan_variants.log):
// match.any + one redux per group (divergent groups: 0.52 vs 0.70 of peak on the id blocks), and up to three
// converged shfl / ballot / 5 x redux rounds (0.60).
// kFlush = 1: before its four min / max reds a lane reads the entry and drops the ones that cannot change it
// (extents only ever grow, so a value that does not improve the current entry never will): interior strips of a
// large object then cost one red, the count.  kFlush = 0: always five reds.
template <bool kSmemTable, int kFlush>
__device__ __noinline__ void flush_entry(uint32_t id, int cnt, int xmn, int xmx, int ymn, int ymx,
                                         const int32_t* lut, int lut_len, int N, int32_t* tab, bool lut_in_smem) {
  const unsigned active = __activemask();
  int same;
  __match_all_sync(active, id, &same);
  if (same) {
    cnt = __reduce_add_sync(active, cnt);
    xmn = __reduce_min_sync(active, xmn);
    xmx = __reduce_max_sync(active, xmx);
    ymn = __reduce_min_sync(active, ymn);
    ymx = __reduce_max_sync(active, ymx);
    if ((threadIdx.x & 31) != __ffs(active) - 1) return;
  }
  if (id >= static_cast<uint32_t>(lut_len)) return;
  // shared-memory copy of the frame's LUT when it fits, else global memory (L1 bypassed: may run before the PDL wait)
  const int slot = lut_in_smem ? lut[id] : __ldcg(lut + id);
  if (static_cast<uint32_t>(slot) >= static_cast<uint32_t>(N)) return;
  int32_t* e = tab + slot * CSPE_SCAN_FIELDS;
  if (kSmemTable) {
    red_shared_add(e + CSPE_SCAN_COUNT, cnt);
    if (kFlush == 1 && !same) {
      const volatile int32_t* ev = e;
      const int c_xmn = ev[CSPE_SCAN_XMIN], c_ymn = ev[CSPE_SCAN_YMIN], c_xmx = ev[CSPE_SCAN_XMAX],
                c_ymx = ev[CSPE_SCAN_YMAX];
      if (xmn < c_xmn) red_shared_min(e + CSPE_SCAN_XMIN, xmn);
      if (ymn < c_ymn) red_shared_min(e + CSPE_SCAN_YMIN, ymn);
      if (xmx > c_xmx) red_shared_max(e + CSPE_SCAN_XMAX, xmx);
      if (ymx > c_ymx) red_shared_max(e + CSPE_SCAN_YMAX, ymx);
    } else {
      red_shared_min(e + CSPE_SCAN_XMIN, xmn);
      red_shared_min(e + CSPE_SCAN_YMIN, ymn);
      red_shared_max(e + CSPE_SCAN_XMAX, xmx);
      red_shared_max(e + CSPE_SCAN_YMAX, ymx);
    }
  } else {
    red_global_add(e + CSPE_SCAN_COUNT, cnt);
    red_global_min(e + CSPE_SCAN_XMIN, xmn);
    red_global_min(e + CSPE_SCAN_YMIN, ymn);
    red_global_max(e + CSPE_SCAN_XMAX, xmx);
    red_global_max(e + CSPE_SCAN_YMAX, ymx);
  }
}

__device__ __forceinline__ int table_identity(int field) {
  return field == CSPE_SCAN_COUNT ? 0 : (field <= CSPE_SCAN_YMIN ? INT_MAX : -1);
}

// make `V` the MRU entry e0 (swap with e1, or evict e1)
#define CSPE_SWITCH(V)                                                                  \
  do {                                                                                  \
    const uint32_t v__ = (V);                                                           \
    if (v__ != e0.id) {                                                                 \
      if (v__ == e1.id) {                                                               \
        Entry t__ = e0;                                                                 \
        e0 = e1;                                                                        \
        e1 = t__;                                                                       \
      } else {                                                                          \
        if (e1.cnt)                                                                     \
          flush_entry<kSmemTable, kFlush>(e1.id, e1.cnt, e1.xmn, e1.xmx, e1.ymn, e1.ymx, lut,   \
                                  p.lut_len, p.N, tab, p.smem_lut != 0);                \
        e1 = e0;                                                                        \
        entry_reset(e0, v__);                                                           \
      }                                                                                 \
    }                                                                                   \
  } while (0)

#define CSPE_ACCUM(CNT, X0, X1)                                                         \
  do {                                                                                  \
    e0.cnt += (CNT);                                                                    \
    e0.xmn = min(e0.xmn, (X0));                                                         \
    e0.xmx = max(e0.xmx, (X1));                                                         \
    e0.ymn = min(e0.ymn, y);                                                            \
    e0.ymx = max(e0.ymx, y);                                                            \
  } while (0)

template <bool kSmemTable, int kBoxes, int kFlush, int kSlow>
// 80 registers (one 4-byte spill): leaves ~19 K registers per SM beside the two resident scan CTAs for
// the small kernels that overlap with it
__global__ void __maxnreg__(80)
    mask_scan_kernel(const ScanParams p, const __grid_constant__ CUtensorMap tmap) {
  using G = Geo<kBoxes>;
  constexpr int kConsumerWarps = G::kConsumerWarps, kConsumers = G::kConsumers, kThreads = G::kThreads;
  constexpr int kTileCols = G::kTileCols, kStageBytes = G::kStageBytes;
  constexpr int kBoxesPerTile = kBoxes;
  extern __shared__ __align__(1024) unsigned char smem[];  // SWIZZLE_128B boxes need 1024-byte alignment
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  int32_t* table = reinterpret_cast<int32_t*>(empty_bar + kStages);
  int32_t* lut_s = table + (kSmemTable ? p.N * CSPE_SCAN_FIELDS : 0);

  const int tid = threadIdx.x;
  const long long p_begin = p.total_passes * blockIdx.x / gridDim.x;
  const long long p_end = p.total_passes * (blockIdx.x + 1) / gridDim.x;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  if (kSmemTable) {
    for (int i = tid; i < p.N * CSPE_SCAN_FIELDS; i += kThreads) table[i] = table_identity(i % CSPE_SCAN_FIELDS);
  }
  __syncthreads();
  // Programmatic dependent launch: every CTA of this persistent grid is resident once it gets
  // here, so a kernel queued behind the scan with programmatic stream serialisation (K2, which
  // does not read the scan's output) may start NOW and run in the registers/SM slots the scan
  // leaves free — without ever displacing a scan CTA (its static work split needs all of them
  // running together).  A no-op when nothing was launched that way.
  pdl_launch_dependents();
  // ... and this kernel is itself launched that way: block placement and the set-up above overlap
  // the tail of the previous kernel; its writes — and the mask, if a kernel produced it — are
  // visible after this.  In overlapped mode the caller promises that only `out` can still be in
  // use by the previous kernel (K4 of the previous batch reading and resetting it), so the wait
  // moves to just before this CTA's first merge into `out` and the streaming starts right away.
  const bool lazy = kSmemTable && p.lazy_wait;
  if (!lazy) pdl_wait();

  // virtual pass index v -> (frame, q); the tile is pq = (q * stride) % ppf -> (row block, segment)
  int frame = static_cast<int>(p_begin / p.ppf);
  int q = static_cast<int>(p_begin % p.ppf);
  int pq = static_cast<int>((static_cast<long long>(q) * p.stride) % p.ppf);

  if (tid >= kConsumers) {
    // ===================== producer warp =====================
    const int lane = tid - kConsumers;
    if (p.tma && lane != 0) return;
    const uint64_t policy = l2_policy_evict_first();
    int it = 0;
    for (long long pp = p_begin; pp < p_end; ++pp, ++it) {
      const int stage = it % kStages;
      const uint32_t parity = (it / kStages) & 1;
      const int rb = pq / p.nseg, seg = pq - rb * p.nseg;
      const int row0 = rb * kTileRows;
      const int col0 = seg * kTileCols;
      unsigned char* dst = smem + stage * kStageBytes;
      mbar_wait(&empty_bar[stage], parity ^ 1);
      if (p.tma == 2) {
        // the whole tile in ONE instruction: the mask seen as [W/32 strips][B*H rows][32 px], box =
        // {32 px, 64 rows, kBoxes strips} lands exactly like kBoxes 2-D boxes side by side (measured
        // +2.7 % on the load side: tools/tma_stream_probe.cu); strips past the last one are zero fill
        mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(kStageBytes));
        tma_load_3d(dst, &tmap, 0, frame * p.H + row0, col0 / kStripPx, &full_bar[stage], policy);
      } else if (p.tma) {
        // boxes that start beyond W are skipped; a box is always written in full (zero fill
        // past the tensor edge), so the byte count is whole boxes
        const int nbox = min(kBoxesPerTile, (p.W - col0 + kStripPx - 1) / kStripPx);
        mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(nbox) * kBoxBytes);
        const int y = frame * p.H + row0;
        for (int b = 0; b < nbox; ++b)
          tma_load_2d(dst + b * kBoxBytes, &tmap, col0 + b * kStripPx, y, &full_bar[stage], policy);
      } else {
        // unaligned fallback: the producer warp fills the same swizzled layout with 4-byte loads
        const int rows = min(kTileRows, p.H - row0);
        const int cols = min(kTileCols, p.W - col0);
        const uint32_t* src = p.mask + (static_cast<long long>(frame) * p.H + row0) * p.W + col0;
        for (int r = 0; r < rows; ++r)
          for (int c = lane; c < cols; c += 32) {
            const int off = (c >> 5) * kBoxBytes + r * 128 + ((((c & 31) >> 2) ^ (r & 7)) << 4) + ((c & 3) << 2);
            // (L1 bypass: in overlapped mode the mask is read before the PDL wait, cspe_common.cuh)
            *reinterpret_cast<uint32_t*>(dst + off) = __ldcg(src + static_cast<long long>(r) * p.W + c);
          }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);  // release: orders the warp's stores (after __syncwarp)
      }
      pq += p.stride;
      if (pq >= p.ppf) pq -= p.ppf;
      if (++q == p.ppf) {
        q = 0;
        pq = 0;
        ++frame;
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  const int lane = tid & 31;
  const int wid = tid >> 5;
  const int box = wid & (kBoxesPerTile - 1);
  const int trow = (wid / kBoxes) * 32 + lane;      // row inside the tile, 0..63
  const int sw = lane & 7;                          // = trow & 7: the 128-byte swizzle phase of this row
  const int sm_off = box * kBoxBytes + trow * 128;
  Entry e0, e1;
  entry_reset(e0, 0u);
  entry_reset(e1, 0u);

  int cur_frame = frame;
  bool waited = false;
  const int32_t* lut = nullptr;
  int32_t* tab = nullptr;

  auto open_frame = [&]() {
    const int32_t* glut = p.lut + static_cast<long long>(cur_frame) * p.lut_stride;
    if (p.smem_lut) {
      for (int i = tid; i < p.lut_len; i += kConsumers) lut_s[i] = __ldcg(glut + i);   // L1 bypass (PDL rule)
      named_bar_sync(1, kConsumers);
      lut = lut_s;
    } else {
      lut = glut;
    }
    tab = kSmemTable ? table : p.out + static_cast<long long>(cur_frame) * p.N * CSPE_SCAN_FIELDS;
  };

  auto close_frame = [&]() {
    if (e0.cnt)
      flush_entry<kSmemTable, kFlush>(e0.id, e0.cnt, e0.xmn, e0.xmx, e0.ymn, e0.ymx, lut, p.lut_len, p.N, tab, p.smem_lut != 0);
    if (e1.cnt)
      flush_entry<kSmemTable, kFlush>(e1.id, e1.cnt, e1.xmn, e1.xmx, e1.ymn, e1.ymx, lut, p.lut_len, p.N, tab, p.smem_lut != 0);
    entry_reset(e0, 0u);
    entry_reset(e1, 0u);
    if (kSmemTable || p.smem_lut) named_bar_sync(1, kConsumers);  // all flushes landed / LUT no longer read
    if (lazy && !waited) {
      pdl_wait();
      waited = true;
    }
    if (kSmemTable) {
      int32_t* gout = p.out + static_cast<long long>(cur_frame) * p.N * CSPE_SCAN_FIELDS;
      for (int i = tid; i < p.N * CSPE_SCAN_FIELDS; i += kConsumers) {
        const int f = i % CSPE_SCAN_FIELDS;
        const int v = table[i];
        const int ident = table_identity(f);
        if (v != ident) {
          if (f == CSPE_SCAN_COUNT) red_global_add(gout + i, v);
          else if (f <= CSPE_SCAN_YMIN) red_global_min(gout + i, v);
          else red_global_max(gout + i, v);
          table[i] = ident;
        }
      }
      named_bar_sync(1, kConsumers);
    }
  };

  if (p_begin < p_end) open_frame();

  int it = 0;
  for (long long pp = p_begin; pp < p_end; ++pp, ++it) {
    if (frame != cur_frame) {
      close_frame();
      cur_frame = frame;
      open_frame();
    }
    const int stage = it % kStages;
    const uint32_t parity = (it / kStages) & 1;
    const int rb = pq / p.nseg, seg = pq - rb * p.nseg;
    const int row0 = rb * kTileRows;
    const int rows = min(kTileRows, p.H - row0);
    const int col0 = seg * kTileCols;

    mbar_wait(&full_bar[stage], parity);

    const int x0 = col0 + box * kStripPx;
    const int len = min(kStripPx, p.W - x0);
    if (trow < rows && len > 0) {
      const int y = row0 + trow;
      const unsigned char* base = smem + stage * kStageBytes + sm_off;
      // logical 16-byte chunk k of this row lives at physical chunk k ^ sw; pixels past `len`
      // (partial last strip) are zero fill or stale and are masked out below
      const uint4 q0 = *reinterpret_cast<const uint4*>(base + ((0 ^ sw) << 4));
      const uint4 q1 = *reinterpret_cast<const uint4*>(base + ((1 ^ sw) << 4));
      const uint4 q2 = *reinterpret_cast<const uint4*>(base + ((2 ^ sw) << 4));
      const uint4 q3 = *reinterpret_cast<const uint4*>(base + ((3 ^ sw) << 4));
      const uint4 q4 = *reinterpret_cast<const uint4*>(base + ((4 ^ sw) << 4));
      const uint4 q5 = *reinterpret_cast<const uint4*>(base + ((5 ^ sw) << 4));
      const uint4 q6 = *reinterpret_cast<const uint4*>(base + ((6 ^ sw) << 4));
      const uint4 q7 = *reinterpret_cast<const uint4*>(base + ((7 ^ sw) << 4));
      const uint32_t a = q0.x;
      uint32_t d = (q0.y ^ a) | (q0.z ^ a) | (q0.w ^ a);
      d |= (q1.x ^ a) | (q1.y ^ a) | (q1.z ^ a) | (q1.w ^ a);
      d |= (q2.x ^ a) | (q2.y ^ a) | (q2.z ^ a) | (q2.w ^ a);
      d |= (q3.x ^ a) | (q3.y ^ a) | (q3.z ^ a) | (q3.w ^ a);
      d |= (q4.x ^ a) | (q4.y ^ a) | (q4.z ^ a) | (q4.w ^ a);
      d |= (q5.x ^ a) | (q5.y ^ a) | (q5.z ^ a) | (q5.w ^ a);
      d |= (q6.x ^ a) | (q6.y ^ a) | (q6.z ^ a) | (q6.w ^ a);
      d |= (q7.x ^ a) | (q7.y ^ a) | (q7.z ^ a) | (q7.w ^ a);
      if (d == 0 && len == kStripPx) {
        // fast path: the whole strip is one id
        CSPE_SWITCH(a);
        CSPE_ACCUM(kStripPx, x0, x0 + kStripPx - 1);
      } else {
        // run decomposition: bit j of `bm` marks a run start (pixel j differs from pixel j-1);
        // all lanes of the warp that are here build the mask in lockstep
        const uint32_t v[kStripPx] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z,
                                      q2.w, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w, q5.x, q5.y,
                                      q5.z, q5.w, q6.x, q6.y, q6.z, q6.w, q7.x, q7.y, q7.z, q7.w};
        uint32_t bm = 0;
#pragma unroll
        for (int j = 1; j < kStripPx; ++j) bm |= (v[j] != v[j - 1]) ? (1u << j) : 0u;
        const uint32_t live = len < kStripPx ? (1u << len) - 1u : 0xffffffffu;
        bm &= live;
        if (kSlow == 1 && __popc(bm) >= kRunWalkMax) {
          // Many runs (a see-through texture: wire mesh, foliage, or an edge that zig-zags): the row
          // usually still holds only two or three DISTINCT ids, so account for it per id instead of
          // per run — equality bit mask of the id over the 32 pixels, count = popc, extent = ffs / clz.
          // After kMaxDistinct ids whatever is left goes through a run walk over the remaining pixels.
          uint32_t rem = live;   // pixels not yet accounted for
          uint32_t id = a;
#pragma unroll 1
          for (int k = 0; k < kMaxDistinct; ++k) {
            // the row is re-read from shared memory (8 conflict-free LDS.128) instead of keeping 32
            // pixel registers alive across the loop: the kernel is capped at 80 registers
            uint32_t m = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 r = *reinterpret_cast<const uint4*>(base + ((c ^ sw) << 4));
              m |= (r.x == id ? 1u : 0u) << (4 * c) | (r.y == id ? 2u : 0u) << (4 * c) |
                   (r.z == id ? 4u : 0u) << (4 * c) | (r.w == id ? 8u : 0u) << (4 * c);
            }
            m &= rem;
            CSPE_SWITCH(id);
            CSPE_ACCUM(__popc(m), x0 + __ffs(m) - 1, x0 + 31 - __clz(m));
            rem &= ~m;
            if (rem == 0) break;
            const int s = __ffs(rem) - 1;
            id = *reinterpret_cast<const uint32_t*>(base + (((s >> 2) ^ sw) << 4) + ((s & 3) << 2));
          }
          const uint32_t stops = bm | ~rem;   // a run ends before the next run start or accounted pixel
#pragma unroll 1
          while (rem) {
            const int s0 = __ffs(rem) - 1;
            const uint32_t t = stops & (0xfffffffeu << s0);
            const int e = t ? __ffs(t) - 2 : len - 1;   // last pixel of the run starting at s0
            const uint32_t rid =
                *reinterpret_cast<const uint32_t*>(base + (((s0 >> 2) ^ sw) << 4) + ((s0 & 3) << 2));
            CSPE_SWITCH(rid);
            CSPE_ACCUM(e - s0 + 1, x0 + s0, x0 + e);
            rem &= 0xfffffffeu << e;   // clears bits 0..e (everything below s0 is already clear)
          }
        } else {
          // few runs (an object edge crossing the strip): walk them, one shared-memory load per run
          int s0 = 0;
#pragma unroll 1
          while (s0 < len) {
            const uint32_t t = bm & (0xfffffffeu << s0);
            const int e = t ? __ffs(t) - 2 : len - 1;   // last pixel of the run starting at s0
            const uint32_t id =
                *reinterpret_cast<const uint32_t*>(base + (((s0 >> 2) ^ sw) << 4) + ((s0 & 3) << 2));
            CSPE_SWITCH(id);
            CSPE_ACCUM(e - s0 + 1, x0 + s0, x0 + e);
            s0 = e + 1;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);

    pq += p.stride;
    if (pq >= p.ppf) pq -= p.ppf;
    if (++q == p.ppf) {
      q = 0;
      pq = 0;
      ++frame;
    }
  }
  if (p_begin < p_end) close_frame();
}

__global__ void scan_init_kernel(int32_t* out, long long n_entries, int W, int H) {
  pdl_launch_dependents();  // the scan's CTAs may set up their rings while this finishes
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_entries * CSPE_SCAN_FIELDS) return;
  const int f = static_cast<int>(i % CSPE_SCAN_FIELDS);
  out[i] = f == CSPE_SCAN_COUNT ? 0 : f == CSPE_SCAN_XMIN ? W : f == CSPE_SCAN_YMIN ? H : -1;
}

int check_common(const uint32_t* mask, int B, int H, int W, const int32_t* id2slot, int lut_len, int64_t lut_stride,
                 int N, int32_t* out) {
  CSPE_REQUIRE(B >= 0 && H >= 0 && W >= 0 && N >= 0 && lut_len >= 0 && lut_stride >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_mask_scan: negative size (B=%d H=%d W=%d N=%d lut_len=%d)", B, H, W, N, lut_len);
  CSPE_REQUIRE(static_cast<long long>(H) * W < (1ll << 31), CSPE_ERR_UNSUPPORTED,
               "cspe_mask_scan: frame of %dx%d pixels overflows int32 counts", W, H);
  if (B == 0 || N == 0) return CSPE_OK;
  CSPE_REQUIRE(out != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_mask_scan: out is null");
  if (H == 0 || W == 0) return CSPE_OK;
  CSPE_REQUIRE(mask != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_mask_scan: mask is null");
  CSPE_REQUIRE(lut_len == 0 || id2slot != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_mask_scan: id2slot is null");
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(mask) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
               CSPE_ERR_INVALID_ARGUMENT, "cspe_mask_scan: mask/out must be 4-byte aligned");
  return 1;  // work to do
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = []() -> TensorMapEncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<TensorMapEncodeFn>(p);
  }();
  return fn;
}

// CSPE_SCAN_TMA_DIMS=2 keeps the one-2-D-box-per-strip producer (for A/B runs); default 3
int scan_tma_dims() {
  static int v = []() {
    const char* e = getenv("CSPE_SCAN_TMA_DIMS");
    return (e && atoi(e) == 2) ? 2 : 3;
  }();
  return v;
}

// A/B switches, read per launch: CSPE_SCAN_FLUSH = 0 always issues the five reds (default 1: lanes that merge on
// their own skip min / max reds that cannot change the entry), CSPE_SCAN_SLOW = 0 the run walk only (default 1: rows
// with many runs are accounted per distinct id).  Measured on B200 (profiles/r02_scan_variants.log): the defaults
// cost ~1.5 % on config 2 and buy 2.5x on see-through textures and +5 % on 16x16-pixel id blocks.
constexpr int kDefaultFlush = 1, kDefaultSlow = 1;
int scan_variant(const char* name, int dflt) {
  const char* e = getenv(name);
  if (!e || !*e) return dflt;
  return atoi(e) != 0 ? 1 : 0;
}

template <int kBoxes>
int launch_scan_geo(const uint32_t* mask, int B, int H, int W, const int32_t* id2slot, int lut_len, int64_t lut_stride,
                    int N, int32_t* out, cudaStream_t st, int lazy_wait) {
  using G = Geo<kBoxes>;
  const int sms = sm_count();
  CSPE_REQUIRE(sms > 0, CSPE_ERR_NO_DEVICE, "cspe_mask_scan: no CUDA device");

  ScanParams p{};
  p.mask = mask;
  p.lut = id2slot;
  p.out = out;
  p.lut_stride = lut_stride;
  p.B = B;
  p.H = H;
  p.W = W;
  p.N = N;
  p.lut_len = lut_len;
  p.lazy_wait = lazy_wait;
  p.nseg = (W + G::kTileCols - 1) / G::kTileCols;
  p.nrb = (H + kTileRows - 1) / kTileRows;
  const long long ppf = static_cast<long long>(p.nseg) * p.nrb;
  CSPE_REQUIRE(ppf < (1ll << 30), CSPE_ERR_UNSUPPORTED, "cspe_mask_scan: frame of %dx%d has too many tiles", W, H);
  CSPE_REQUIRE(static_cast<long long>(B) * H < (1ll << 31), CSPE_ERR_UNSUPPORTED,
               "cspe_mask_scan: B*H = %lld rows exceeds the int32 tensor coordinate", static_cast<long long>(B) * H);
  p.ppf = static_cast<int>(ppf);
  p.total_passes = ppf * B;
  const long long max_ctas = static_cast<long long>(sms) * G::kCtasPerSm;
  const int grid = static_cast<int>(p.total_passes < max_ctas ? p.total_passes : max_ctas);

  // TMA needs 16-byte aligned rows; otherwise the producer warp copies by hand
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  p.tma = 0;
  TensorMapEncodeFn encode = tensor_map_encoder();
  if (encode != nullptr && (W % 4 == 0) && (reinterpret_cast<uintptr_t>(mask) & 15) == 0) {
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(B) * H};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(W) * 4};
    const cuuint32_t box[2] = {kStripPx, kBoxRows};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(mask), gdim, gstride, box,
                              estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    p.tma = (r == CUDA_SUCCESS) ? 1 : 0;
    if (p.tma && W % kStripPx == 0 && scan_tma_dims() == 3) {
      // same bytes, viewed as [W/32 strips][B*H rows][32 px] (strip stride 128 B): one box per tile
      CUtensorMap tmap3;
      const cuuint64_t gdim3[3] = {kStripPx, static_cast<cuuint64_t>(B) * H, static_cast<cuuint64_t>(W / kStripPx)};
      const cuuint64_t gstride3[2] = {static_cast<cuuint64_t>(W) * 4, kStripPx * 4};
      const cuuint32_t box3[3] = {kStripPx, kBoxRows, static_cast<cuuint32_t>(kBoxes)};
      const cuuint32_t estride3[3] = {1, 1, 1};
      if (encode(&tmap3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint32_t*>(mask), gdim3, gstride3, box3, estride3,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
        tmap = tmap3;
        p.tma = 2;
      }
    }
  }

  // pass permutation: when several CTAs share a frame, stride the tile order so each CTA's
  // contiguous range samples the whole frame (busy and empty regions alike)
  const long long per_cta = p.total_passes / grid;
  int stride = 1;
  if (per_cta < ppf) {
    stride = static_cast<int>((ppf + per_cta - 1) / per_cta);
    auto gcd = [](long long x, long long y) {
      while (y) {
        const long long t = x % y;
        x = y;
        y = t;
      }
      return x;
    };
    while (gcd(stride, ppf) != 1) ++stride;
    if (stride >= ppf) stride = 1;
  }
  p.stride = stride;

  const size_t table_bytes = static_cast<size_t>(N) * CSPE_SCAN_FIELDS * 4;
  const size_t lut_bytes = static_cast<size_t>(lut_len) * 4;
  const bool smem_table = table_bytes <= static_cast<size_t>(G::kSmemFree);
  p.smem_lut = lut_len > 0 && (smem_table ? table_bytes : 0) + lut_bytes <= static_cast<size_t>(G::kSmemFree);
  const size_t smem_bytes = G::kSmemFixed + (smem_table ? table_bytes : 0) + (p.smem_lut ? lut_bytes : 0);

  const int flush_v = scan_variant("CSPE_SCAN_FLUSH", kDefaultFlush), slow_v = scan_variant("CSPE_SCAN_SLOW", kDefaultSlow);
  auto kern = smem_table ? mask_scan_kernel<true, kBoxes, kDefaultFlush, kDefaultSlow>
                         : mask_scan_kernel<false, kBoxes, kDefaultFlush, kDefaultSlow>;
  if constexpr (kBoxes == kDefaultBoxes) {   // A/B variants exist for the default geometry only
    if (flush_v == 0 && slow_v == 0)
      kern = smem_table ? mask_scan_kernel<true, kBoxes, 0, 0> : mask_scan_kernel<false, kBoxes, 0, 0>;
    else if (flush_v == 0 && slow_v == 1)
      kern = smem_table ? mask_scan_kernel<true, kBoxes, 0, 1> : mask_scan_kernel<false, kBoxes, 0, 1>;
    else if (flush_v == 1 && slow_v == 0)
      kern = smem_table ? mask_scan_kernel<true, kBoxes, 1, 0> : mask_scan_kernel<false, kBoxes, 1, 0>;
    else
      kern = smem_table ? mask_scan_kernel<true, kBoxes, 1, 1> : mask_scan_kernel<false, kBoxes, 1, 1>;
  }
  CSPE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  CSPE_CUDA_OK(launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(G::kThreads), smem_bytes, st, p, tmap));
  return CSPE_OK;
}

// boxes per tile: 8 -> 1 CTA/SM, 4 -> 2 CTAs/SM, 2 -> 4 CTAs/SM (CSPE_SCAN_BOXES overrides, for tuning)
int scan_boxes() {
  static int v = []() {
    const char* e = getenv("CSPE_SCAN_BOXES");
    const int x = e ? atoi(e) : 0;
    return (x == 8 || x == 4 || x == 2) ? x : kDefaultBoxes;
  }();
  return v;
}

int launch_scan(const uint32_t* mask, int B, int H, int W, const int32_t* id2slot, int lut_len, int64_t lut_stride, int N,
                int32_t* out, cudaStream_t st, int lazy_wait = 0) {
  switch (scan_boxes()) {
    case 8:
      return launch_scan_geo<8>(mask, B, H, W, id2slot, lut_len, lut_stride, N, out, st, lazy_wait);
    case 2:
      return launch_scan_geo<2>(mask, B, H, W, id2slot, lut_len, lut_stride, N, out, st, lazy_wait);
    default:
      return launch_scan_geo<4>(mask, B, H, W, id2slot, lut_len, lut_stride, N, out, st, lazy_wait);
  }
}

int launch_init(int32_t* out, int B, int N, int W, int H, cudaStream_t st) {
  const long long n = static_cast<long long>(B) * N;
  const long long total = n * CSPE_SCAN_FIELDS;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  scan_init_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(out, n, W, H);
  CSPE_LAUNCH_OK("scan_init_kernel");
  return CSPE_OK;
}

}  // namespace
}  // namespace cspe

using namespace cspe;

// 1 when the scan of a [B][H][W] batch launches its full persistent grid (every SM holds its two CTAs and
// is out of shared memory): only then is it impossible for the scan of batch i+2 to become resident before
// the scan of batch i+1 has exited, which is what orders the double-buffered K2 / K3 outputs of an overlapped
// pipeline behind the K4 that still reads them.
extern "C" int cspe_mask_scan_fills_device(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  const int sms = sm_count();
  if (sms <= 0) return 0;
  switch (scan_boxes()) {
    case 8: {
      using G = Geo<8>;
      const long long passes = static_cast<long long>((W + G::kTileCols - 1) / G::kTileCols) * ((H + kTileRows - 1) / kTileRows) * B;
      return passes >= static_cast<long long>(sms) * G::kCtasPerSm;
    }
    case 2: {
      using G = Geo<2>;
      const long long passes = static_cast<long long>((W + G::kTileCols - 1) / G::kTileCols) * ((H + kTileRows - 1) / kTileRows) * B;
      return passes >= static_cast<long long>(sms) * G::kCtasPerSm;
    }
    default: {
      using G = Geo<4>;
      const long long passes = static_cast<long long>((W + G::kTileCols - 1) / G::kTileCols) * ((H + kTileRows - 1) / kTileRows) * B;
      return passes >= static_cast<long long>(sms) * G::kCtasPerSm;
    }
  }
}

extern "C" int cspe_mask_scan_accumulate(const uint32_t* mask, int B, int H, int W, const int32_t* id2slot,
                                         int lut_len, int64_t lut_stride, int N, int32_t* out, void* stream) {
  const int c = check_common(mask, B, H, W, id2slot, lut_len, lut_stride, N, out);
  if (c <= 0) return c;
  return launch_scan(mask, B, H, W, id2slot, lut_len, lut_stride, N, out,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int cspe_mask_scan_accumulate_overlapped(const uint32_t* mask, int B, int H, int W, const int32_t* id2slot,
                                                    int lut_len, int64_t lut_stride, int N, int32_t* out,
                                                    void* stream) {
  const int c = check_common(mask, B, H, W, id2slot, lut_len, lut_stride, N, out);
  if (c <= 0) return c;
  return launch_scan(mask, B, H, W, id2slot, lut_len, lut_stride, N, out, static_cast<cudaStream_t>(stream), 1);
}

extern "C" int cspe_mask_scan(const uint32_t* mask, int B, int H, int W, const int32_t* id2slot, int lut_len,
                              int64_t lut_stride, int N, int32_t* out, void* stream) {
  const int c = check_common(mask, B, H, W, id2slot, lut_len, lut_stride, N, out);
  if (c < 0) return c;
  if (B <= 0 || N <= 0) return CSPE_OK;
  const int rc = launch_init(out, B, N, W, H, static_cast<cudaStream_t>(stream));
  if (rc != CSPE_OK || c == 0) return rc;
  return launch_scan(mask, B, H, W, id2slot, lut_len, lut_stride, N, out,
                            static_cast<cudaStream_t>(stream));
}

