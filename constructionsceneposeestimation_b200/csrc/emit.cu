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
#include <stddef.h>

#include "cspe_common.cuh"

static_assert(sizeof(cspe_record) == 408, "cspe_record layout changed: update the host dtype");

namespace cspe {
namespace {

constexpr int kEmitChunk = 256;   // slots decided per pass (= header staging capacity)
constexpr int kHdrWords = 22;                                   // 88-byte integer/float header of cspe_record
constexpr int kRecWords = sizeof(cspe_record) / 4;              // 102
static_assert(offsetof(cspe_record, uv) == kHdrWords * 4, "header size");
static_assert(offsetof(cspe_record, z) == offsetof(cspe_record, uv) + 128, "uv block");
static_assert(offsetof(cspe_record, pose) == offsetof(cspe_record, z) + 64, "z block");

// Phase 1: one thread per slot decides keep / computes the 88-byte header into shared memory
// and its stable rank (ballot + prefix).  Phase 2: each warp writes its kept records
// cooperatively — 102 coalesced 4-byte words per record (header from shared memory, the 40
// doubles straight from K2's arrays) instead of one thread issuing ~100 serial stores.
// kThreads = 256 for frames of up to 256 slots, 1024 beyond: the first 256 threads decide, ALL
// threads stream the records (that copy is what takes time when a frame holds hundreds of objects)
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
    emit_kernel(const int32_t* scan, int32_t* scan_reset, const double* __restrict__ uv, const double* __restrict__ z,
                const double* __restrict__ pose, const double* __restrict__ loose, const uint8_t* __restrict__ flags,
                const int32_t* __restrict__ slot_class, int N, int H, int W, int min_pixels, int frame_base,
                const int32_t* __restrict__ frame_base_dev, cspe_record* __restrict__ records, int32_t* __restrict__ n_out,
                unsigned long long* __restrict__ class_hist) {
  __shared__ int warp_sums[kEmitChunk / 32];
  __shared__ int base_s;
  __shared__ int hist_s[CSPE_NUM_CLASSES];
  __shared__ int32_t hdr_s[kEmitChunk][kHdrWords + 1];  // +1: odd pitch, conflict-free column writes
  __shared__ int kept_s[kEmitChunk];                      // rank within the chunk -> thread (slot - n0)
  const int f = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  // let whatever is queued behind K4 get placed while K4 runs (the next batch's mask scan in
  // overlapped mode streams its mask meanwhile and waits for us before it merges into `scan`)
  pdl_launch_dependents();
  if (tid < CSPE_NUM_CLASSES) hist_s[tid] = 0;
  if (tid == 0) base_s = 0;
  __syncthreads();
  // launched with programmatic stream serialisation: everything above overlapped the tail of the
  // kernel before us; its results (and, transitively, the mask scan's) are complete after this
  pdl_wait();

  // graph-replayable frame numbering: the id of the batch's first frame may come from device memory
  // Everything below was written by the kernels this one depends on WHILE it was already resident (programmatic
  // dependent launch), so none of it is read-only for this kernel's lifetime: no ld.global.nc (__ldg / const
  // __restrict__), every such load is an explicit L2 load (__ldcg) — cspe_common.cuh, PDL rule.
  if (frame_base_dev) frame_base += __ldcg(frame_base_dev);
  const int32_t* uv32 = reinterpret_cast<const int32_t*>(uv);
  const int32_t* z32 = reinterpret_cast<const int32_t*>(z);
  const int32_t* pose32 = reinterpret_cast<const int32_t*>(pose);

  for (int n0 = 0; n0 < N; n0 += kEmitChunk) {
    const int n = tid < kEmitChunk ? n0 + tid : N;  // threads beyond the chunk only help with the copy
    const long long o = static_cast<long long>(f) * N + n;
    bool keep = false;
    int cls = -1, cnt = 0, x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    uint8_t fl = 0;
    if (n < N) {
      cls = __ldcg(slot_class + o);
      const int32_t* sc = scan + o * CSPE_SCAN_FIELDS;
      cnt = __ldcg(sc + CSPE_SCAN_COUNT);
      x0 = __ldcg(sc + CSPE_SCAN_XMIN);
      y0 = __ldcg(sc + CSPE_SCAN_YMIN);
      x1 = __ldcg(sc + CSPE_SCAN_XMAX);
      y1 = __ldcg(sc + CSPE_SCAN_YMAX);
      if (scan_reset) {  // leave the entry as cspe_mask_scan's init would: the next batch can accumulate
        int32_t* sr = scan_reset + o * CSPE_SCAN_FIELDS;
        sr[CSPE_SCAN_COUNT] = 0;
        sr[CSPE_SCAN_XMIN] = W;
        sr[CSPE_SCAN_YMIN] = H;
        sr[CSPE_SCAN_XMAX] = -1;
        sr[CSPE_SCAN_YMAX] = -1;
      }
      fl = __ldcg(flags + o);
      keep = (cls >= 0) && (cnt >= min_pixels) && (fl & CSPE_OBJ_ANY_FRONT);
    }
    // stable rank = exclusive prefix sum of keep flags
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0 && wid < kEmitChunk / 32) warp_sums[wid] = __popc(bal);

    if (keep) {
      cspe_record h;  // only the header fields are filled; lives in registers
      h.frame = frame_base + f;
      h.inst_idx = n;
      h.class_id = cls;
      h.count = cnt;
      h.x_min = x0;
      h.y_min = y0;
      h.x_max = x1;
      h.y_max = y1;
      h.flags = fl;
      h.pad0 = 0;
      const int tw = x1 - x0 + 1, th = y1 - y0 + 1;
      const long long tight_area = static_cast<long long>(tw) * th;
      // loose box: projected 3D box clipped to the image and integerised
      const double umin = __ldcg(loose + o * 4 + 0), vmin = __ldcg(loose + o * 4 + 1), umax = __ldcg(loose + o * 4 + 2),
                   vmax = __ldcg(loose + o * 4 + 3);
      const double dW = static_cast<double>(W), dH = static_cast<double>(H);
      const int lx0 = static_cast<int>(fmin(fmax(floor(umin), 0.0), dW));
      const int ly0 = static_cast<int>(fmin(fmax(floor(vmin), 0.0), dH));
      const int lx1 = static_cast<int>(fmax(fmin(ceil(umax) - 1.0, dW - 1.0), -1.0));
      const int ly1 = static_cast<int>(fmax(fmin(ceil(vmax) - 1.0, dH - 1.0), -1.0));
      const int lw = max(0, lx1 - lx0 + 1), lh = max(0, ly1 - ly0 + 1);
      const long long loose_area = static_cast<long long>(lw) * lh;
      if (loose_area > 0) {
        h.loose[0] = lx0;
        h.loose[1] = ly0;
        h.loose[2] = lx1;
        h.loose[3] = ly1;
      } else {
        h.loose[0] = 0;
        h.loose[1] = 0;
        h.loose[2] = -1;
        h.loose[3] = -1;
      }
      const float fcnt = static_cast<float>(cnt);
      const float vis = loose_area > 0 ? fminf(1.0f, fcnt / static_cast<float>(loose_area)) : 0.0f;
      h.visible_frac = vis;
      h.occlusion = 1.0f - vis;
      h.fill = cnt > 0 ? fcnt / static_cast<float>(tight_area) : 0.0f;  // min_pixels == 0 keeps unseen objects
      const double ua = (umax - umin) * (vmax - vmin);
      const double cwid = fmax(fmin(umax, dW) - fmax(umin, 0.0), 0.0);
      const double chei = fmax(fmin(vmax, dH) - fmax(vmin, 0.0), 0.0);
      const double ca = cwid * chei;
      h.truncation = ua > 0.0 ? 1.0f - static_cast<float>(ca) / static_cast<float>(ua) : 1.0f;
      const float fW = static_cast<float>(W), fH = static_cast<float>(H);
      h.yolo[0] = cnt > 0 ? (static_cast<float>(x0 + x1 + 1) * 0.5f) / fW : 0.0f;
      h.yolo[1] = cnt > 0 ? (static_cast<float>(y0 + y1 + 1) * 0.5f) / fH : 0.0f;
      h.yolo[2] = cnt > 0 ? static_cast<float>(tw) / fW : 0.0f;
      h.yolo[3] = cnt > 0 ? static_cast<float>(th) / fH : 0.0f;
      int32_t* hs = hdr_s[tid];
      hs[0] = h.frame;
      hs[1] = h.inst_idx;
      hs[2] = h.class_id;
      hs[3] = h.count;
      hs[4] = h.x_min;
      hs[5] = h.y_min;
      hs[6] = h.x_max;
      hs[7] = h.y_max;
      hs[8] = h.flags;
      hs[9] = h.loose[0];
      hs[10] = h.loose[1];
      hs[11] = h.loose[2];
      hs[12] = h.loose[3];
      hs[13] = h.pad0;
      hs[14] = __float_as_int(h.occlusion);
      hs[15] = __float_as_int(h.fill);
      hs[16] = __float_as_int(h.truncation);
      hs[17] = __float_as_int(h.visible_frac);
      hs[18] = __float_as_int(h.yolo[0]);
      hs[19] = __float_as_int(h.yolo[1]);
      hs[20] = __float_as_int(h.yolo[2]);
      hs[21] = __float_as_int(h.yolo[3]);
      if (cls < CSPE_NUM_CLASSES) atomicAdd(&hist_s[cls], 1);
    }
    __syncthreads();  // warp_sums and hdr_s visible
    int before = 0;
    int chunk_total = 0;
#pragma unroll
    for (int w = 0; w < kEmitChunk / 32; ++w) {
      const int s = warp_sums[w];
      if (w < wid) before += s;
      chunk_total += s;
    }
    if (keep) kept_s[before + in_warp] = tid;
    __syncthreads();
    // cooperative write: the kept records of this chunk are one contiguous run of 102-word
    // records starting at rank `base_s`; all 256 threads stream it word by word (coalesced stores,
    // independent loads), looking the source slot up through kept_s
    {
      const int chunk_base = base_s;
      int32_t* dst = reinterpret_cast<int32_t*>(records + static_cast<long long>(f) * N + chunk_base);
      const int words = chunk_total * kRecWords;
#pragma unroll 8
      for (int w = tid; w < words; w += kThreads) {
        const int r = w / kRecWords, k = w - r * kRecWords;
        const int t = kept_s[r];
        const long long so = static_cast<long long>(f) * N + n0 + t;
        int32_t v;
        if (k < kHdrWords) v = hdr_s[t][k];
        else if (k < kHdrWords + 32) v = __ldcg(uv32 + so * 32 + (k - kHdrWords));
        else if (k < kHdrWords + 48) v = __ldcg(z32 + so * 16 + (k - kHdrWords - 32));
        else v = __ldcg(pose32 + so * 32 + (k - kHdrWords - 48));
        dst[w] = v;
      }
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

static int emit_impl(const int32_t* scan, int32_t* scan_reset, const double* uv, const double* z, const double* pose,
                     const double* loose, const uint8_t* flags, const int32_t* slot_class, int B, int N, int H, int W,
                     int min_pixels, int frame_base, const int32_t* frame_base_dev, cspe_record* records,
                     int32_t* n_out, int64_t* class_hist, void* stream) {
  CSPE_REQUIRE(B >= 0 && N >= 0 && H >= 0 && W >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_emit: negative size (B=%d N=%d H=%d W=%d)", B, N, H, W);
  if (B == 0) return CSPE_OK;
  CSPE_REQUIRE(n_out && class_hist, CSPE_ERR_INVALID_ARGUMENT, "cspe_emit: n_out/class_hist is null");
  CSPE_REQUIRE(N == 0 || (scan && uv && z && pose && loose && flags && slot_class && records),
               CSPE_ERR_INVALID_ARGUMENT, "cspe_emit: null pointer");
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(records) & 7) == 0 && (reinterpret_cast<uintptr_t>(class_hist) & 7) == 0,
               CSPE_ERR_INVALID_ARGUMENT, "cspe_emit: records/class_hist must be 8-byte aligned");
  if (N <= kEmitChunk) {
    CSPE_CUDA_OK(launch_pdl(emit_kernel<256>, dim3(static_cast<unsigned>(B)), dim3(256), 0,
                            static_cast<cudaStream_t>(stream), scan, scan_reset, uv, z, pose, loose, flags, slot_class,
                            N, H, W, min_pixels, frame_base, frame_base_dev, records, n_out,
                            reinterpret_cast<unsigned long long*>(class_hist)));
  } else {
    CSPE_CUDA_OK(launch_pdl(emit_kernel<1024>, dim3(static_cast<unsigned>(B)), dim3(1024), 0,
                            static_cast<cudaStream_t>(stream), scan, scan_reset, uv, z, pose, loose, flags, slot_class,
                            N, H, W, min_pixels, frame_base, frame_base_dev, records, n_out,
                            reinterpret_cast<unsigned long long*>(class_hist)));
  }
  return CSPE_OK;
}

extern "C" int cspe_emit(const int32_t* scan, const double* uv, const double* z, const double* pose,
                         const double* loose, const uint8_t* flags, const int32_t* slot_class, int B, int N, int H,
                         int W, int min_pixels, int frame_base, cspe_record* records, int32_t* n_out,
                         int64_t* class_hist, void* stream) {
  return emit_impl(scan, nullptr, uv, z, pose, loose, flags, slot_class, B, N, H, W, min_pixels, frame_base, nullptr,
                   records, n_out, class_hist, stream);
}

extern "C" int cspe_emit_reset_scan(int32_t* scan, const double* uv, const double* z, const double* pose,
                                    const double* loose, const uint8_t* flags, const int32_t* slot_class, int B, int N,
                                    int H, int W, int min_pixels, int frame_base, cspe_record* records,
                                    int32_t* n_out, int64_t* class_hist, void* stream) {
  return emit_impl(scan, scan, uv, z, pose, loose, flags, slot_class, B, N, H, W, min_pixels, frame_base, nullptr,
                   records, n_out, class_hist, stream);
}

extern "C" int cspe_emit_reset_scan_indirect(int32_t* scan, const double* uv, const double* z, const double* pose,
                                             const double* loose, const uint8_t* flags, const int32_t* slot_class,
                                             int B, int N, int H, int W, int min_pixels, int frame_base,
                                             const int32_t* frame_base_dev, cspe_record* records, int32_t* n_out,
                                             int64_t* class_hist, void* stream) {
  CSPE_REQUIRE(frame_base_dev != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_emit_reset_scan_indirect: frame_base_dev is null");
  return emit_impl(scan, scan, uv, z, pose, loose, flags, slot_class, B, N, H, W, min_pixels, frame_base,
                   frame_base_dev, records, n_out, class_hist, stream);
}
