// f3 — "%.6f" text serialisation on the device.
//
// The reference writes its two largest per-frame files with numpy's text writer:
//   np.savetxt(depth_csv_path, depth_data, delimiter=' ', fmt='%.6f')                gcd.py:1688
//   np.savetxt(pcd_path, xyzrgb, fmt='%.6f', delimiter=' ', header='x y z r g b',
//              comments='')                                                           gcd.py:1752-1753
// i.e. 2 M (depth) and up to 12 M (point cloud) numbers per 1080p frame formatted one Python
// '%' call per row — seconds per frame, far more than every numeric stage together.  Here the
// matrix is formatted where it already lives: one thread per value decodes the IEEE bits into
// an integer part and six correctly rounded decimals with integer arithmetic only (exact for
// every finite input, ties to even like printf / Python), the text of a 1024-value tile is
// assembled in shared memory and streamed out as aligned 16-byte stores.  The output is the
// byte stream np.savetxt writes (header line, ' ' between columns, '\n' after each row,
// "nan" / "inf" / "-inf" as Python prints them).
//
// Three launches chained by programmatic dependent launch:
//   1. per tile: decode, sum of the text lengths -> tile_count;
//   2. one-CTA exclusive scan -> tile byte offsets, total size;
//   3. per tile: decode again, in-tile offsets, build the text in shared memory, stream it out.
// Decoding twice is cheaper than spilling 12 M (ip, frac) pairs through HBM.
#include <math.h>
#include <string.h>

#include "cspe_common.cuh"

namespace cspe {
namespace {

constexpr int kTxThreads = 256;
constexpr int kTxPerThread = 4;
constexpr int kTxTile = kTxThreads * kTxPerThread;  // 1024 values
constexpr int kTxMaxChars = 48;                     // '-' + 39 digits + '.' + 6 + delimiter
constexpr int kTxStageBytes = kTxTile * kTxMaxChars + 16;
constexpr int kTxMaxHeader = 63;

typedef unsigned __int128 u128;

struct TxWorkspace {  // layout of the caller-provided scratch
  unsigned int flags;  // bit 0: a finite value too large for the formatter (|x| >= 2^128)
  unsigned int pad;
  long long total;
  // followed by: int64 tile_offset[tiles]; int32 tile_count[tiles]
};

struct TxHeader {
  char text[kTxMaxHeader + 1];  // header line including its '\n'
  int len;
};

enum : int { kFinite = 0, kWide = 1, kNan = 2, kInf = 3, kTooLarge = 4 };

struct Fixed6 {
  unsigned long long ip;  // integer part (kFinite)
  unsigned long long fm;  // fraction = fm / 2^s (before rounding)
  unsigned int frac;      // six decimals, 0..999999 (after round_fixed6)
  int s;
  int kind;
  bool neg;
};

__constant__ unsigned long long kPow10[20] = {1ull,
                                              10ull,
                                              100ull,
                                              1000ull,
                                              10000ull,
                                              100000ull,
                                              1000000ull,
                                              10000000ull,
                                              100000000ull,
                                              1000000000ull,
                                              10000000000ull,
                                              100000000000ull,
                                              1000000000000ull,
                                              10000000000000ull,
                                              100000000000000ull,
                                              1000000000000000ull,
                                              10000000000000000ull,
                                              100000000000000000ull,
                                              1000000000000000000ull,
                                              10000000000000000000ull};

// x = m * 2^E with integer m < 2^53: the integer part is a shift, the fraction is the s = -E low
// bits of m.  No rounding yet.
__device__ __forceinline__ Fixed6 split_fixed6(double x) {
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(x));
  Fixed6 r;
  r.neg = (bits >> 63) != 0;
  r.ip = 0;
  r.fm = 0;
  r.frac = 0;
  r.s = 0;
  r.kind = kFinite;
  const int be = static_cast<int>((bits >> 52) & 0x7ffu);
  unsigned long long m = bits & ((1ull << 52) - 1);
  if (be == 0x7ff) {
    r.kind = m ? kNan : kInf;
    if (m) r.neg = false;  // Python prints "nan" whatever the sign bit
    return r;
  }
  int E = -1074;
  if (be) {
    m |= 1ull << 52;
    E = be - 1075;
  }
  if (E >= 0) {  // |x| >= 2^52: an integer
    if (E <= 11)
      r.ip = m << E;
    else
      r.kind = E <= 75 ? kWide : kTooLarge;
    return r;
  }
  r.s = -E;  // 1..1074 fraction bits
  r.fm = m;
  if (r.s < 64) {
    r.ip = m >> r.s;
    r.fm = m & ((1ull << r.s) - 1);
  }
  return r;
}

// round-half-even(fm * 10^6 / 2^s) for 64 <= s < 75 (128-bit product; rare, kept out of line)
__device__ __noinline__ unsigned long long round_tiny(unsigned long long fm, int s) {
  const u128 P = static_cast<u128>(fm) * 1000000u;
  const u128 qq = P >> s;
  const u128 rem = P - (qq << s), half = static_cast<u128>(1) << (s - 1);
  const unsigned long long q = static_cast<unsigned long long>(qq);
  return q + ((rem > half || (rem == half && (q & 1))) ? 1 : 0);
}

// Six decimals of fm / 2^s, exactly rounded (ties to even); returns true when the fraction rounds up to
// 1.000000, i.e. carries into the integer part.  Integer arithmetic only.
__device__ __forceinline__ bool round_fixed6(unsigned long long fm, int s, unsigned int* frac) {
  *frac = 0;
  if (fm == 0) return false;
  const int tz = __ffsll(static_cast<long long>(fm)) - 1;  // strip trailing zeros: float32 inputs lose 29 bits here
  fm >>= tz;
  s -= tz;
  if (s >= 75) return false;  // fm * 10^6 < 2^73: rounds to .000000
  unsigned long long q;
  bool up;
  if (s <= 50) {
    // one 64-bit product: fm * 10^6 < 2^64 for s <= 44, else fm * 15625 / 2^(s-6) (10^6 = 2^6 * 15625)
    const int sh = s <= 44 ? s : s - 6;
    const unsigned long long P = fm * (s <= 44 ? 1000000ull : 15625ull);
    q = P >> sh;
    const unsigned long long rem = P & ((1ull << sh) - 1), half = 1ull << (sh - 1);
    up = rem > half || (rem == half && (q & 1));
  } else if (s <= 63) {
    const unsigned long long lo = fm * 1000000ull, hi = __umul64hi(fm, 1000000ull);
    q = (hi << (64 - s)) | (lo >> s);
    const unsigned long long rem = lo & ((1ull << s) - 1), half = 1ull << (s - 1);
    up = rem > half || (rem == half && (q & 1));
  } else {
    q = round_tiny(fm, s);  // |x| < 2^-11
    up = false;
  }
  q += up;
  if (q == 1000000ull) return true;
  *frac = static_cast<unsigned int>(q);
  return false;
}

__device__ __forceinline__ Fixed6 decode_fixed6(double x) {
  Fixed6 r = split_fixed6(x);
  if (r.kind == kFinite && round_fixed6(r.fm, r.s, &r.frac)) ++r.ip;  // ip < 2^53: cannot overflow
  return r;
}

// |x| as a 128-bit integer for 2^64 <= |x| < 2^128 (kWide)
__device__ __noinline__ u128 wide_integer(double x) {
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(x));
  const int be = static_cast<int>((bits >> 52) & 0x7ffu);
  const unsigned long long m = (bits & ((1ull << 52) - 1)) | (1ull << 52);
  return static_cast<u128>(m) << (be - 1075);
}

__device__ __forceinline__ int digits_u32(unsigned int v) {
  int n = 1;
  if (v >= 100000000u) {
    v /= 100000000u;
    n += 8;
  }
  if (v >= 10000u) {
    v /= 10000u;
    n += 4;
  }
  if (v >= 100u) {
    v /= 100u;
    n += 2;
  }
  return n + (v >= 10u);
}

__device__ __forceinline__ int digits_u64(unsigned long long v) {
  if ((v >> 32) == 0) return digits_u32(static_cast<unsigned int>(v));  // coordinates, depths, colours
  int n = 0;
  while (v >> 32) {
    v /= 100000000ull;
    n += 8;
  }
  return n + digits_u32(static_cast<unsigned int>(v));
}

__device__ __noinline__ int digits_u128(u128 v) {
  int n = 0;
  const u128 chunk = static_cast<u128>(10000000000000000000ull);  // 10^19
  while (v >= chunk) {
    v /= chunk;
    n += 19;
  }
  return n + digits_u64(static_cast<unsigned long long>(v));
}

__device__ __noinline__ int special_length(int kind, bool neg, double x) {
  if (kind == kNan) return 3;
  if (kind == kInf) return 3 + static_cast<int>(neg);
  if (kind == kWide) return static_cast<int>(neg) + digits_u128(wide_integer(x)) + 7;
  return 1;  // kTooLarge: the whole call is flagged; one '?' keeps the layout consistent
}

// length of the text of one value without its delimiter
__device__ __forceinline__ int fixed6_length(const Fixed6& r, double x) {
  if (r.kind == kFinite) return static_cast<int>(r.neg) + digits_u64(r.ip) + 7;
  return special_length(r.kind, r.neg, x);
}

// pass 1: the length needs the rounding only when a carry could add a digit (integer part 9, 99, 999 ...)
__device__ __forceinline__ int fixed6_length_only(double x, bool* too_large) {
  const Fixed6 r = split_fixed6(x);
  if (r.kind == kFinite) {
    int n = digits_u64(r.ip);
    if (r.ip + 1 == kPow10[n]) {
      unsigned int frac;
      n += round_fixed6(r.fm, r.s, &frac);
    }
    return static_cast<int>(r.neg) + n + 7;
  }
  *too_large |= r.kind == kTooLarge;
  return fixed6_length(r, x);
}

__device__ __noinline__ void special_put(char* p, int kind, bool neg, unsigned int frac, double x, int len) {
  if (kind == kWide) {
    char* e = p + len;
    for (int i = 0; i < 6; ++i) {
      *--e = static_cast<char>('0' + frac % 10u);
      frac /= 10u;
    }
    *--e = '.';
    u128 v = wide_integer(x);
    do {
      *--e = static_cast<char>('0' + static_cast<unsigned int>(v % 10u));
      v /= 10u;
    } while (v);
    if (neg) *--e = '-';
  } else if (kind == kNan) {
    p[0] = 'n', p[1] = 'a', p[2] = 'n';
  } else if (kind == kInf) {
    if (neg) *p++ = '-';
    p[0] = 'i', p[1] = 'n', p[2] = 'f';
  } else {
    p[0] = '?';
  }
}

// write the text of one value (len characters, as fixed6_length says) at p
__device__ __forceinline__ void fixed6_put(char* p, const Fixed6& r, double x, int len) {
  if (r.kind == kFinite) {
    char* e = p + len;  // write backwards
    unsigned int f = r.frac;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      *--e = static_cast<char>('0' + f % 10u);
      f /= 10u;
    }
    *--e = '.';
    unsigned long long v = r.ip;
    while (v >> 32) {  // rare: peel eight digits at a time down to 32 bits
      unsigned int lo = static_cast<unsigned int>(v % 100000000ull);
      v /= 100000000ull;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        *--e = static_cast<char>('0' + lo % 10u);
        lo /= 10u;
      }
    }
    unsigned int w = static_cast<unsigned int>(v);
    do {
      *--e = static_cast<char>('0' + w % 10u);
      w /= 10u;
    } while (w);
    if (r.neg) *--e = '-';
    return;
  }
  special_put(p, r.kind, r.neg, r.frac, x, len);
}

template <bool kF64>
__device__ __forceinline__ double load_value(const void* __restrict__ values, long long i) {
  // read before griddepcontrol.wait in tx_write_kernel: L1 must be bypassed (cspe_common.cuh, PDL rule)
  if (kF64) return __ldcg(static_cast<const double*>(values) + i);
  return static_cast<double>(__ldcg(static_cast<const float*>(values) + i));  // exact widening
}

__device__ __forceinline__ long long live_values(long long max_rows, const long long* n_rows, int cols) {
  long long rows = max_rows;
  if (n_rows) {
    const long long n = __ldcg(n_rows);   // L1 bypass: read before the PDL wait (cspe_common.cuh)
    rows = n < 0 ? 0 : (n < max_rows ? n : max_rows);
  }
  return rows * cols;
}

template <bool kF64>
__global__ void __launch_bounds__(kTxThreads)
    tx_length_kernel(const void* __restrict__ values, long long max_rows, const long long* __restrict__ n_rows, int cols,
                     int32_t* __restrict__ tile_count, TxWorkspace* ws) {
  pdl_launch_dependents();
  __shared__ int s_len[kTxThreads / 32];
  const long long total = live_values(max_rows, n_rows, cols);
  const long long base = static_cast<long long>(blockIdx.x) * kTxTile + threadIdx.x * kTxPerThread;
  int len = 0;
  bool too_large = false;
#pragma unroll
  for (int k = 0; k < kTxPerThread; ++k) {
    if (base + k < total) len += fixed6_length_only(load_value<kF64>(values, base + k), &too_large) + 1;  // + ' ' or '\n'
  }
  if (too_large) atomicOr(&ws->flags, 1u);
  len = __reduce_add_sync(0xffffffffu, len);
  if ((threadIdx.x & 31) == 0) s_len[threadIdx.x >> 5] = len;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kTxThreads / 32; ++w) t += s_len[w];
    tile_count[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) tx_scan_kernel(const int32_t* __restrict__ tile_count, long long* tile_offset,
                                                      int tiles, long long base, TxWorkspace* ws, long long* n_bytes) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ long long s_warp[32];
  __shared__ long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    s_base = base;
    if (tiles == 0) tile_offset[0] = base;  // empty matrix: the single writer CTA only emits the header
  }
  __syncthreads();
  for (int t0 = 0; t0 < tiles; t0 += 1024) {
    const int t = t0 + tid;
    const long long v = t < tiles ? __ldcg(tile_count + t) : 0;   // written by the kernel this one waited for: L2 load
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
    if (t < tiles) tile_offset[t] = s_base + (wid ? s_warp[wid - 1] : 0) + inc - v;
    __syncthreads();
    if (tid == 0) s_base += s_warp[31];
    __syncthreads();
  }
  if (tid == 0) {
    ws->total = s_base;
    *n_bytes = (__ldcg(&ws->flags) & 1u) ? -1 : s_base;
  }
}

template <bool kF64>
__global__ void __launch_bounds__(kTxThreads)
    tx_write_kernel(const void* __restrict__ values, long long max_rows, const long long* __restrict__ n_rows, int cols,
                    TxHeader header, const long long* __restrict__ tile_offset, char* __restrict__ text,
                    long long capacity, long long split_values, long long* __restrict__ split_offsets) {
  pdl_launch_dependents();
  extern __shared__ __align__(16) char stage[];
  __shared__ int s_warp[kTxThreads / 32];
  const long long total = live_values(max_rows, n_rows, cols);
  const long long base = static_cast<long long>(blockIdx.x) * kTxTile + threadIdx.x * kTxPerThread;
  // the values are an input of the chain: decode before waiting for the scan
  double x[kTxPerThread];
  Fixed6 r[kTxPerThread];
  int len[kTxPerThread];
  int mine = 0;
#pragma unroll
  for (int k = 0; k < kTxPerThread; ++k) {
    len[k] = 0;
    if (base + k < total) {
      x[k] = load_value<kF64>(values, base + k);
      r[k] = decode_fixed6(x[k]);
      len[k] = fixed6_length(r[k], x[k]);
      mine += len[k] + 1;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int off = inc - mine, tile_bytes = 0;
#pragma unroll
  for (int w = 0; w < kTxThreads / 32; ++w) {
    if (w < wid) off += s_warp[w];
    tile_bytes += s_warp[w];
  }

  pdl_wait();  // tile offsets come from the scan kernel
  const long long first = __ldcg(tile_offset + blockIdx.x);   // written while this kernel was resident: L2 load
  // shared-memory text starts at the same offset modulo 16 as its place in `text`, so that the
  // copy below moves whole aligned 16-byte words
  const int skew = static_cast<int>(reinterpret_cast<uintptr_t>(text + first) & 15);
  long long col = base % cols;
  long long split_pos = split_offsets ? base % split_values : 1;  // position inside the current split
#pragma unroll
  for (int k = 0; k < kTxPerThread; ++k) {
    if (base + k >= total) break;
    char* p = stage + skew + off;
    fixed6_put(p, r[k], x[k], len[k]);
    p[len[k]] = (col == cols - 1) ? '\n' : ' ';
    if (split_offsets) {
      if (split_pos == 0) split_offsets[(base + k) / split_values] = first + off;
      if (++split_pos == split_values) split_pos = 0;
    }
    off += len[k] + 1;
    if (++col == cols) col = 0;
  }
  if (blockIdx.x == 0) {  // header line (np.savetxt: comments + header + '\n')
    for (int j = threadIdx.x; j < header.len && j < capacity; j += kTxThreads) text[j] = header.text[j];
  }
  __syncthreads();

  long long keep = capacity - first;  // bytes beyond `capacity` are dropped but were counted
  if (keep > tile_bytes) keep = tile_bytes;
  if (keep <= 0) return;
  const int begin = skew, end = skew + static_cast<int>(keep);  // byte range of the stage to copy
  char* g0 = text + first - skew;                               // 16-byte aligned
  const int w_begin = (begin + 15) >> 4, w_end = end >> 4;      // whole words
  if (w_begin <= w_end) {
    for (int j = begin + threadIdx.x; j < (w_begin << 4); j += kTxThreads) g0[j] = stage[j];
    const int4* s4 = reinterpret_cast<const int4*>(stage);
    int4* g4 = reinterpret_cast<int4*>(g0);
    for (int j = w_begin + threadIdx.x; j < w_end; j += kTxThreads) {
      const int4 v = s4[j];
      asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(g4 + j), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                   : "memory");
    }
    for (int j = (w_end << 4) + threadIdx.x; j < end; j += kTxThreads) g0[j] = stage[j];
  } else {  // the whole tile lies inside one 16-byte word
    for (int j = begin + threadIdx.x; j < end; j += kTxThreads) g0[j] = stage[j];
  }
}

long long tx_tiles(long long values) { return (values + kTxTile - 1) / kTxTile; }

}  // namespace
}  // namespace cspe

using namespace cspe;

extern "C" size_t cspe_text_workspace_bytes(int64_t max_rows, int cols) {
  long long tiles = (max_rows <= 0 || cols <= 0) ? 0 : tx_tiles(static_cast<long long>(max_rows) * cols);
  if (tiles < 1) tiles = 1;
  return sizeof(TxWorkspace) + static_cast<size_t>(tiles) * (8 + 4) + 16;
}

extern "C" int cspe_format_fixed6(const void* values, int dtype, int64_t max_rows, const int64_t* n_rows, int cols,
                                  const char* header, char* text, int64_t capacity, int64_t* n_bytes,
                                  int64_t split_rows, int64_t* split_offsets, void* workspace, void* stream) {
  CSPE_REQUIRE(dtype == CSPE_DTYPE_F32 || dtype == CSPE_DTYPE_F64, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_fixed6: dtype %d is neither CSPE_DTYPE_F32 nor CSPE_DTYPE_F64", dtype);
  CSPE_REQUIRE(max_rows >= 0 && cols >= 0 && capacity >= 0 && split_rows >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_fixed6: negative size");
  CSPE_REQUIRE(n_bytes && workspace, CSPE_ERR_INVALID_ARGUMENT, "cspe_format_fixed6: null pointer");
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_fixed6: workspace must be 8-byte aligned");
  CSPE_REQUIRE(split_rows == 0 || split_offsets != nullptr, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_format_fixed6: split_rows given without split_offsets");
  TxHeader hdr;
  memset(&hdr, 0, sizeof(hdr));
  if (header != nullptr) {
    const size_t n = strlen(header);
    CSPE_REQUIRE(n < static_cast<size_t>(kTxMaxHeader), CSPE_ERR_UNSUPPORTED,
                 "cspe_format_fixed6: header longer than %d bytes", kTxMaxHeader - 1);
    CSPE_REQUIRE(memchr(header, '\n', n) == nullptr, CSPE_ERR_INVALID_ARGUMENT,
                 "cspe_format_fixed6: header must be one line");
    memcpy(hdr.text, header, n);
    hdr.text[n] = '\n';
    hdr.len = static_cast<int>(n) + 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TxWorkspace* ws = static_cast<TxWorkspace*>(workspace);
  const long long total = static_cast<long long>(max_rows) * cols;
  CSPE_REQUIRE(cols == 0 || total / cols == max_rows, CSPE_ERR_UNSUPPORTED, "cspe_format_fixed6: matrix too large");
  const long long tiles = tx_tiles(total);
  CSPE_REQUIRE(tiles < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_format_fixed6: matrix too large");
  CSPE_REQUIRE(total == 0 || values != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_format_fixed6: values is null");
  CSPE_REQUIRE(capacity == 0 || text != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_format_fixed6: text is null");
  CSPE_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(TxWorkspace), st));
  // an empty matrix still runs the scan and one writer CTA: n_bytes and the header line are
  // produced on the stream like everything else (the workspace always holds one tile slot)
  long long* tile_offset = reinterpret_cast<long long*>(ws + 1);
  int32_t* tile_count = reinterpret_cast<int32_t*>(tile_offset + (tiles > 0 ? tiles : 1));
  const long long split_values = split_rows * cols;
  long long* split = split_values > 0 ? reinterpret_cast<long long*>(split_offsets) : nullptr;
  const long long* rows_dev = reinterpret_cast<const long long*>(n_rows);
  const unsigned grid = static_cast<unsigned>(tiles > 0 ? tiles : 1);
  const int scan_tiles = static_cast<int>(tiles);
  static const cudaError_t attr32 = cudaFuncSetAttribute(tx_write_kernel<false>,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kTxStageBytes);
  static const cudaError_t attr64 = cudaFuncSetAttribute(tx_write_kernel<true>,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kTxStageBytes);
  (void)attr32;
  (void)attr64;
  if (tiles > 0) {
    // plain launch first (serialised behind whatever produced the values), then a PDL chain
    if (dtype == CSPE_DTYPE_F64)
      tx_length_kernel<true><<<grid, kTxThreads, 0, st>>>(values, max_rows, rows_dev, cols, tile_count, ws);
    else
      tx_length_kernel<false><<<grid, kTxThreads, 0, st>>>(values, max_rows, rows_dev, cols, tile_count, ws);
    CSPE_LAUNCH_OK("tx_length_kernel");
  }
  CSPE_CUDA_OK(launch_pdl(tx_scan_kernel, dim3(1), dim3(1024), 0, st, static_cast<const int32_t*>(tile_count), tile_offset,
                          scan_tiles, static_cast<long long>(hdr.len), ws, reinterpret_cast<long long*>(n_bytes)));
  if (capacity > 0) {
    if (dtype == CSPE_DTYPE_F64)
      CSPE_CUDA_OK(launch_pdl(tx_write_kernel<true>, dim3(grid), dim3(kTxThreads), kTxStageBytes, st, values,
                              static_cast<long long>(max_rows), rows_dev, cols, hdr,
                              static_cast<const long long*>(tile_offset), text, static_cast<long long>(capacity),
                              split_values, split));
    else
      CSPE_CUDA_OK(launch_pdl(tx_write_kernel<false>, dim3(grid), dim3(kTxThreads), kTxStageBytes, st, values,
                              static_cast<long long>(max_rows), rows_dev, cols, hdr,
                              static_cast<const long long*>(tile_offset), text, static_cast<long long>(capacity),
                              split_values, split));
  }
  return CSPE_OK;
}
