// Python's repr() of round(x, 6) — shared by the host COCO formatter (label_json.cpp) and the device one
// (yolo_text.cu).  COCO "occlusion" / "truncation" are float32 ratios; the label files carry them rounded to six
// decimals (formats.coco_annotations: round(float(v), 6)), which is what makes them printable on the device: for a
// float32 v with |v| < 2^20 the product v * 1e6 is exact in double, q = round-half-even(|v| * 1e6) is the correctly
// rounded 6-decimal value, and repr(q / 1e6) is those decimals with trailing zeros dropped — fixed notation from
// 1e-4 up, d[.ddd]e-0X below (Python switches to the exponent form when the decimal point would need more than
// three leading zeros), "0.0" / "-0.0" for zero.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define CSPE_HD __host__ __device__ __forceinline__
#else
#define CSPE_HD inline
#endif

namespace cspe {

// decimal digits of v, most significant first; returns the count (v = 0 -> "0")
CSPE_HD int put_decimal(unsigned long long v, char* out) {
  char tmp[20];
  int n = 0;
  do {
    tmp[n++] = static_cast<char>('0' + static_cast<int>(v % 10ull));
    v /= 10ull;
  } while (v);
  for (int i = 0; i < n; ++i) out[i] = tmp[n - 1 - i];
  return n;
}

// repr(round(x, 6)) for x = (neg ? -1 : 1) * q * 1e-6; at most 28 characters; returns the length
CSPE_HD int repr_units6(bool neg, unsigned long long q, char* out) {
  int n = 0;
  if (neg) out[n++] = '-';
  if (q == 0) {
    out[n++] = '0';
    out[n++] = '.';
    out[n++] = '0';
    return n;
  }
  if (q < 100ull) {  // below 1e-4: exponent form, e.g. 5e-06, 1.2e-05
    const int tens = static_cast<int>(q / 10ull), ones = static_cast<int>(q % 10ull);
    if (tens == 0) {
      out[n++] = static_cast<char>('0' + ones);
      out[n++] = 'e';
      out[n++] = '-';
      out[n++] = '0';
      out[n++] = '6';
    } else {
      out[n++] = static_cast<char>('0' + tens);
      if (ones) {
        out[n++] = '.';
        out[n++] = static_cast<char>('0' + ones);
      }
      out[n++] = 'e';
      out[n++] = '-';
      out[n++] = '0';
      out[n++] = '5';
    }
    return n;
  }
  n += put_decimal(q / 1000000ull, out + n);
  out[n++] = '.';
  unsigned int frac = static_cast<unsigned int>(q % 1000000ull);
  char d[6];
  for (int i = 5; i >= 0; --i) {
    d[i] = static_cast<char>('0' + frac % 10u);
    frac /= 10u;
  }
  int last = 5;
  while (last > 0 && d[last] == '0') --last;   // keep at least one fractional digit
  for (int i = 0; i <= last; ++i) out[n++] = d[i];
  return n;
}

}  // namespace cspe
