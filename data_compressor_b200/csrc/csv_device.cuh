// csv_device.cuh -- per-value bodies of the csv codec: `encode csv` (DCLib/src/csv.c:46-65, WriteCSV) = printf("%.*f\n") of
// a float32, described next, and `decode csv`'s field conversion (csv_to_float, at the end).
//
// WriteCSV prints every value with sprintf("%.*f\n", num_decimal_places, value) (csv.c:59).  printf converts the exact
// binary value and rounds it to the requested digits half-to-even; the sign bit prints as '-' (also for zero and nan),
// the non-finite values print as "inf" / "nan".  A float32 is m * 2^e with m < 2^24, so the exact conversion is
// integer work:
//   e <  0 : T = round_half_even(m * 10^d / 2^-e) with m * 10^d < 2^44 (d <= 6), one 64-bit shift and a remainder test;
//            the digits of T with the point d places from the right are the text;
//   e >= 0 : the value is the integer m << e (up to 2^128); its digits, then '.' and d zeros.
// The same routine runs twice per value: once to measure (WRITE = false), once to write at the scanned position.
#pragma once
#include <stdint.h>

namespace dc {

static const uint32_t kCsvMaxDecimals = 6;     // the reference's option range (enc_dec.c:69)
static const uint32_t kCsvMaxBody = 48;        // '-' + 39 digits + '.' + 6 digits + '\n'

__host__ __device__ __forceinline__ uint64_t csv_pow10(uint32_t d)
{
  uint64_t p = 1;
  for (uint32_t i = 0; i < d; i++)
    p *= 10;
  return p;
}

__host__ __device__ __forceinline__ uint32_t csv_count_digits(uint64_t t)
{
  uint32_t n = 1;
  uint64_t p = 10;
  while (n < 20 && t >= p)
  {
    p *= 10;
    n++;
  }
  return n;
}

// Length in bytes of the text of `v` with `d` decimals including the newline; with WRITE the text goes to dst.
template <bool WRITE>
__host__ __device__ __forceinline__ uint32_t csv_print(const uint32_t bits, const uint32_t d, uint8_t *dst)
{
  uint32_t len = 0;
  if (bits >> 31)
  {
    if (WRITE)
      dst[0] = '-';
    len = 1;
  }
  const uint32_t ex = (bits >> 23) & 0xFF, fr = bits & 0x7FFFFF;
  if (ex == 0xFF)
  {
    if (WRITE)
    {
      dst[len] = fr ? 'n' : 'i';
      dst[len + 1] = fr ? 'a' : 'n';
      dst[len + 2] = fr ? 'n' : 'f';
      dst[len + 3] = '\n';
    }
    return len + 4;
  }
  const uint32_t m = ex ? (fr | 0x800000u) : fr;
  const int e = (ex ? (int)ex : 1) - 150;       // value = m * 2^e
  const uint32_t point = d ? 1 : 0;
  if (e <= 40)
  {
    uint64_t t;              // the digits to print: the value scaled by 10^d (e < 0) or the integer itself
    if (e < 0)
    {
      const uint32_t s = (uint32_t)(-e);
      const uint64_t p = (uint64_t)m * csv_pow10(d);
      if (s >= 64)
        t = 0;               // p < 2^44: below one half of the last digit
      else
      {
        t = p >> s;
        const uint64_t r = p & ((1ull << s) - 1), half = 1ull << (s - 1);
        if (r > half || (r == half && (t & 1)))
          t++;
      }
      uint32_t nd = csv_count_digits(t);
      if (nd < d + 1)
        nd = d + 1;          // "0.05": at least one digit before the point
      const uint32_t total = nd + point;
      if (WRITE)
      {
        uint8_t *q = dst + len + total - 1;
        for (uint32_t i = 0; i < nd; i++)
        {
          if (d && i == d)
            *q-- = '.';
          const uint64_t t10 = t / 10;
          *q-- = (uint8_t)('0' + (uint32_t)(t - t10 * 10));
          t = t10;
        }
        dst[len + total] = '\n';
      }
      return len + total + 1;
    }
    t = (uint64_t)m << e;
    const uint32_t nd = csv_count_digits(t);
    if (WRITE)
    {
      uint8_t *q = dst + len + nd - 1;
      for (uint32_t i = 0; i < nd; i++)
      {
        const uint64_t t10 = t / 10;
        *q-- = (uint8_t)('0' + (uint32_t)(t - t10 * 10));
        t = t10;
      }
      q = dst + len + nd;
      if (d)
        *q++ = '.';
      for (uint32_t i = 0; i < d; i++)
        *q++ = '0';
      *q = '\n';
    }
    return len + nd + point + d + 1;
  }
  // integers of 2^64 and more: four 32-bit limbs, nine digits per pass of a division by 10^9
  uint32_t limb[4] = {0, 0, 0, 0};
  {
    const uint32_t w = (uint32_t)e >> 5, sh = (uint32_t)e & 31;   // e <= 104: w <= 3
    const uint64_t v = (uint64_t)m << sh;
    for (uint32_t k = 0; k < 4; k++)
    {
      if (k == w)
        limb[k] = (uint32_t)v;
      if (k == w + 1)
        limb[k] = (uint32_t)(v >> 32);
    }
  }
  uint32_t chunk[5];
  uint32_t nchunk = 0;
  for (;;)
  {
    uint64_t rem = 0;
    for (int k = 3; k >= 0; k--)
    {
      const uint64_t cur = (rem << 32) | limb[k];
      limb[k] = (uint32_t)(cur / 1000000000u);
      rem = cur % 1000000000u;
    }
    chunk[nchunk++] = (uint32_t)rem;
    if ((limb[0] | limb[1] | limb[2] | limb[3]) == 0 || nchunk == 5)
      break;
  }
  const uint32_t nd = csv_count_digits(chunk[nchunk - 1]) + 9 * (nchunk - 1);
  if (WRITE)
  {
    uint8_t *q = dst + len + nd - 1;
    for (uint32_t c = 0; c < nchunk; c++)
    {
      uint32_t t = chunk[c];
      const uint32_t k = c + 1 == nchunk ? nd - 9 * c : 9;
      for (uint32_t i = 0; i < k; i++)
      {
        const uint32_t t10 = t / 10;
        *q-- = (uint8_t)('0' + (t - t10 * 10));
        t = t10;
      }
    }
    q = dst + len + nd;
    if (d)
      *q++ = '.';
    for (uint32_t i = 0; i < d; i++)
      *q++ = '0';
    *q = '\n';
  }
  return len + nd + point + d + 1;
}

// ---- `decode csv`: one field -> float32 --------------------------------------------------------------------------
__device__ __forceinline__ bool csv_is_space(uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
__device__ __forceinline__ bool csv_is_digit(uint8_t c) { return c >= '0' && c <= '9'; }

// strtof on text[fs, fe) for the part of its grammar that one correctly rounded operation converts (see the head of
// csv_kernels.cu); `hard` is set for the rest (the returned value is then meaningless and the caller asks strtof).
__device__ inline float csv_to_float(const uint8_t *text, const uint64_t fs, const uint64_t fe, bool &hard_out)
{
  const float pow10f[11] = { 1.0f, 10.0f, 100.0f, 1000.0f, 10000.0f, 100000.0f, 1000000.0f, 10000000.0f, 100000000.0f, 1000000000.0f, 10000000000.0f };
  uint64_t i = fs;
  while (i < fe && csv_is_space(text[i]))
    i++;
  bool neg = false;
  if (i < fe && (text[i] == '+' || text[i] == '-'))
    neg = text[i++] == '-';
  const uint64_t num_start = i;
  uint64_t M = 0;
  uint32_t sig = 0, k = 0, ndig = 0;
  bool hard = false, frac = false;
  for (; i < fe; i++)
  {
    const uint8_t c = text[i];
    if (csv_is_digit(c))
    {
      ndig++;
      if (sig > 0 || c != '0')
      {
        if (sig < 19)
          M = M * 10 + (uint64_t)(c - '0');
        else
          hard = true;
        sig++;
      }
      if (frac)
        k++;
    }
    else if (c == '.' && !frac)
      frac = true;
    else
      break;
  }
  float f = 0.0f;
  if (ndig == 0)
  {
    // no conversion: strtof returns +0 ... unless the text spells inf / nan
    if (i < fe && (text[i] == 'i' || text[i] == 'I' || text[i] == 'n' || text[i] == 'N'))
      hard = true;
    neg = false;
  }
  else
  {
    const uint8_t c = i < fe ? text[i] : 0;
    if ((c == 'x' || c == 'X') && i == num_start + 1 && text[num_start] == '0')
      hard = true;                                            // hex float
    // decimal exponent: e[+-]digits; an 'e' without digits is not part of the number (strtof stops before it)
    int e10 = 0;
    if ((c == 'e' || c == 'E') && !hard)
    {
      uint64_t q = i + 1;
      bool eneg = false;
      if (q < fe && (text[q] == '+' || text[q] == '-'))
        eneg = text[q++] == '-';
      if (q < fe && csv_is_digit(text[q]))
      {
        uint32_t ev = 0;
        for (; q < fe && csv_is_digit(text[q]); q++)
          ev = ev < 100000u ? ev * 10 + (uint32_t)(text[q] - '0') : ev;
        e10 = eneg ? -(int)ev : (int)ev;
      }
    }
    while (M != 0 && M % 10 == 0 && sig > 1)                  // trailing zeros of the digit string move into the exponent
    {
      M /= 10;
      e10++;
      sig--;
    }
    e10 -= (int)k;                                            // value = M * 10^e10
    if (hard)
      ;
    else if (M == 0)
      f = 0.0f;
    else if (e10 == 0)
      f = __ull2float_rn(M);                                  // exact integer, one rounding
    else if (e10 < 0)
    {
      if (M < (1ull << 24) && e10 >= -10)
        f = __fdiv_rn((float)M, pow10f[-e10]);                // both operands exact: the division rounds once
      else
        hard = true;
    }
    else if (sig + (uint32_t)e10 <= 19)
    {
      uint64_t v = M;
      for (int t = 0; t < e10; t++)
        v *= 10;
      f = __ull2float_rn(v);
    }
    else if (M < (1ull << 24) && e10 <= 10)
      f = __fmul_rn((float)M, pow10f[e10]);
    else
      hard = true;
  }
  hard_out = hard;
  return neg ? -f : f;
}

}  // namespace dc
