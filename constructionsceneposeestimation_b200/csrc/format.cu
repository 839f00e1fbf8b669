// f3 — host-side label serialisation: YOLO text straight from the D2H
// record buffer.  Once the kernels run at HBM speed the Python formatter is the end-to-end
// limiter of a 100 k-frame sweep (~150 us per 100-object frame); this is plain C, no CUDA.
//
// Output is byte-identical to the Python reference formatter
//   f"{class_id} {cx:.6f} {cy:.6f} {w:.6f} {h:.6f}\n"
// for float32 values v in [0, 2^20): v * 1e6 is exact in double (24-bit significand times
// 2^6 * 15625), so round-half-even of that product equals the correctly rounded decimal that
// printf / Python produce.
#include <math.h>
#include <string.h>

#include "cspe_common.cuh"

namespace cspe {
namespace {

inline char* put_uint(char* p, unsigned long long v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = static_cast<char>('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

// "%.6f" of a finite float32 with |v| < 2^20; returns nullptr if v is out of that domain
inline char* put_fixed6(char* p, float v) {
  if (!(fabsf(v) < 1048576.0f)) return nullptr;
  double t = static_cast<double>(v) * 1e6;  // exact
  if (signbit(v)) {
    *p++ = '-';
    t = -t;
  }
  const unsigned long long q = static_cast<unsigned long long>(nearbyint(t));  // ties to even (default mode)
  p = put_uint(p, q / 1000000ull);
  *p++ = '.';
  unsigned long long frac = q % 1000000ull;
  for (int i = 5; i >= 0; --i) {
    p[i] = static_cast<char>('0' + frac % 10);
    frac /= 10;
  }
  return p + 6;
}

}  // namespace
}  // namespace cspe

using namespace cspe;

extern "C" int64_t cspe_format_yolo_host(const cspe_record* records_host, const int32_t* n_out_host, int B, int N,
                                         int frames, char* out_host, int64_t capacity, int64_t* offsets_host) {
  if (B < 0 || N < 0 || frames < 0 || frames > B || capacity < 0 || !offsets_host ||
      (frames > 0 && (!records_host || !n_out_host)) || (capacity > 0 && !out_host)) {
    set_error("cspe_format_yolo_host: invalid argument");
    return CSPE_ERR_INVALID_ARGUMENT;
  }
  char* p = out_host;
  char* const end = out_host + capacity;
  for (int f = 0; f < frames; ++f) {
    offsets_host[f] = p - out_host;
    const int n = n_out_host[f];
    if (n < 0 || n > N) {
      set_error("cspe_format_yolo_host: n_out[%d] = %d outside [0, %d]", f, n, N);
      return CSPE_ERR_INVALID_ARGUMENT;
    }
    const cspe_record* r = records_host + static_cast<int64_t>(f) * N;
    for (int i = 0; i < n; ++i, ++r) {
      if (end - p < 96) {  // longest line: 11 + 4 * (1 + 8 + 6) + 5 < 96
        set_error("cspe_format_yolo_host: output buffer of %lld bytes is too small", static_cast<long long>(capacity));
        return CSPE_ERR_INVALID_ARGUMENT;
      }
      if (r->class_id < 0) *p++ = '-';
      p = put_uint(p, static_cast<unsigned long long>(r->class_id < 0 ? -static_cast<long long>(r->class_id) : r->class_id));
      for (int k = 0; k < 4; ++k) {
        *p++ = ' ';
        char* q = put_fixed6(p, r->yolo[k]);
        if (!q) {
          set_error("cspe_format_yolo_host: frame %d record %d has a non-finite / out-of-range box", f, i);
          return CSPE_ERR_INVALID_ARGUMENT;
        }
        p = q;
      }
      *p++ = '\n';
    }
  }
  offsets_host[frames] = p - out_host;
  return p - out_host;
}
