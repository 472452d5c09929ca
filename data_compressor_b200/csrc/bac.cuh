// bac.cuh -- device-side core of the binary arithmetic coder (reference: DCLib/src/bac.c).
//
// The reference is the Witten/Neal/Cleary coder with RANGE_BITS = 16 over the three symbols {0, 1, EOF}
// (bac.c:13,29-31).  This restatement keeps the same integer results but re-expresses the state so that one
// symbol costs a fixed, branch-free instruction sequence per thread (the kernel is integer-issue bound):
//
//   state     reference keeps (start_range, end_range) (bac.c:83); we keep (LO, rng): LO = start_range << 16
//             (top-aligned in a 32-bit register) and rng = end - start + 1, because every renormalisation step
//             doubles rng whatever branch it takes.
//   model     the reference keeps freq[4]/cum[4]/two index maps (bac.c:33-37).  EOF never moves and keeps
//             frequency 1 (UpdateModel is never called with index 3; halving maps 1 -> 1, bac.c:63) and index 0 is
//             a sentinel, so the whole adaptive model is (f_top, c1 = f_oth + 1, topmask, total): the counts of
//             the data symbols at index 1 / 2 and which bit sits at index 1.  cum[] = {total, c1, 1, 0}.
//   update    index 1: hi unchanged, lo += A                    A = floor(rng * c1 / total)
//             index 2: lo += B, rng = A - B                      B = floor(rng / total)
//             EOF    : rng = B                                    (bac.c:109-111 with the cum[] above)
//   division  total <= 16383 and rng * c1 <= 2^29, so floor(x / total) == umulhi(x, m) >> s with a
//             precomputed round-up reciprocal (m, s) per divisor (kernels.h: DivEntry); exact, no correction step.
//   renorm    bac.c:112-137 is a loop of three cases.  Closed form: first n1 = number of equal leading bits of
//             (lo, hi) "E1/E2" steps, each emitting that bit; then n3 = length of the run "lo = 01..., hi = 10..."
//             below the top bit, each an "E3" step (pending++).  After an E3 step lo < 0x8000 <= hi, so no E1/E2
//             step can follow within the same symbol.  With HI = (hi << 16) | 0xFFFF both counts are one CLZ each
//             and the low ones of HI cap them without a special case for lo == hi.
//   emission  OutputNextBits (bac.c:93-105) writes a bit followed by `pending` inverted bits.  For n1 >= 1 E1/E2
//             steps the output is the top n1 bits of lo with the pending run spliced in after the first one, i.e.
//             the (n1 + pending)-bit number  bits + 2^(n1+pending-1) - 2^(n1-1)  -- one append, no loop.
#pragma once

#include <stdint.h>

#include "bitio.cuh"
#include "kernels.h"

namespace dc {

static const uint32_t kBacMaxFreq = 0x3FFFu;   // MAX_FREQUENCY, bac.c:27

// PTX shifts clamp the amount to 32 (C++ shifts by >= 32 are undefined)
#ifdef DC_HOST_SIM
static inline uint32_t shl_c(uint32_t a, uint32_t n) { return n >= 32 ? 0u : a << n; }
static inline uint32_t shr_c(uint32_t a, uint32_t n) { return n >= 32 ? 0u : a >> n; }
#else
__device__ __forceinline__ uint32_t shl_c(uint32_t a, uint32_t n)
{
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(n));
  return r;
}
__device__ __forceinline__ uint32_t shr_c(uint32_t a, uint32_t n)
{
  uint32_t r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(n));
  return r;
}
#endif

// Count of leading zeros of a NON-ZERO word: bfind.shiftamt is a single FLO.U32.SH (the generic __clz costs FLO + IADD)
#ifdef DC_HOST_SIM
static inline uint32_t clz_nz(uint32_t a) { return (uint32_t)__builtin_clz(a); }
#else
__device__ __forceinline__ uint32_t clz_nz(uint32_t a)
{
  uint32_t r;
  asm("bfind.shiftamt.u32 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}
#endif

// floor(x / d) for x < 2^30 with the table entry of d
__device__ __forceinline__ uint32_t div_magic(uint32_t x, const DivEntry e) { return __umulhi(x, e.m) >> e.s; }

struct BacModel
{
  uint32_t f_top;     // freq[1]
  uint32_t topmask;   // index_to_symbol[1] as 0 / 0x80000000
  uint32_t total;     // cum[0] = freq[1] + freq[2] + 1, so cum[1] = freq[2] + 1 = total - f_top

  __device__ __forceinline__ void init()   // InitModel, bac.c:39-52
  {
    f_top = 1;
    topmask = 0;
    total = 3;
  }

  __device__ __forceinline__ uint32_t c1() const { return total - f_top; }   // cum[1]

  // bac.c:57-67: halve, rounding up; EOF stays at 1.  Only when total == kBacMaxFreq.
  __device__ __forceinline__ void halve()
  {
    const uint32_t f_oth = (total - f_top) >> 1;   // ((freq[2]) + 1) / 2 with freq[2] = total - f_top - 1
    f_top = (f_top + 1) >> 1;
    total = f_top + f_oth + 1;
  }

  // UpdateModel, bac.c:68-80, for a data symbol; bm = the coded bit as 0 / 0x80000000, is_top == (bm == topmask)
  __device__ __forceinline__ void bump(uint32_t bm, bool is_top)
  {
    // bac.c:68: a symbol at index 2 moves left past its equal-frequency neighbour (freq[2] == freq[1], i.e.
    // total == 2 * freq[1] + 1), then freq[1]++; otherwise freq[2]++ (implied by total++)
    const bool inc_top = is_top || (total == 2 * f_top + 1);
    topmask = inc_top ? bm : topmask;     // is_top: bm == topmask already; tie: the coded symbol takes index 1
    f_top += inc_top ? 1u : 0u;
    total++;
  }
};

// Interval after coding a symbol: LO0 (top-aligned lo), r1 (range).  Returns the E1/E2 count n1, the E3 count n3 and
// P1 = 2^n1, and leaves the renormalised interval in (LO, rng).  bac.c:112-137.
//
// Pipe balance: ncu shows this kernel bound by the integer ALU pipe (shifts, logic, compares) with the FMA pipe (IMAD)
// two-thirds idle, so variable shifts are done as multiplications by 2^n wherever the product is needed anyway.
// MUL_SHIFT picks how the final `<< n` of (rng, LO) is done: as multiplications by 2^n (FMA pipe) or as shifts (ALU pipe).
// The encoder is FMA-heavy elsewhere and takes shifts; the decoder is ALU-heavy elsewhere and takes multiplications
// (profiles/r1d_dega_ncu_summary.txt: both pipes end up near 70 %).
template <bool MUL_SHIFT>
__device__ __forceinline__ void bac_renorm(uint32_t LO0, uint32_t r1, uint32_t &LO, uint32_t &rng, uint32_t &n1, uint32_t &n3, uint32_t &P1)
{
  const uint32_t HI = LO0 + (r1 << 16) - 1;   // (hi << 16) | 0xFFFF
  n1 = clz_nz(LO0 ^ HI);                      // 0..16 (the low 16 bits of the xor are ones)
  P1 = shl_c(1u, n1);                         // opaque to the optimiser: keeps the multiplications below
  const uint32_t z = ~LO0 | HI;               // clear exactly where lo has 1 and hi has 0
  n3 = clz_nz((z * P1) << 1);                 // run below the first differing bit, capped by the guard ones
  if (MUL_SHIFT)
  {
    const uint32_t Pn = shl_c(P1, n3);        // n1 + n3 <= 31
    rng = r1 * Pn;
    LO = (LO0 * Pn) & 0x7FFF0000u;
  }
  else
  {
    const uint32_t n = n1 + n3;               // <= 31
    rng = r1 << n;
    LO = (LO0 << n) & 0x7FFF0000u;
  }
}

struct BacEncoder
{
  uint32_t LO;      // start_range << 16
  uint32_t rng;     // end_range - start_range + 1, in [1, 65536]
  uint32_t pend;    // next_bits, bac.c:84

  __device__ __forceinline__ void init()   // StartEncoding, bac.c:86-91
  {
    LO = 0;
    rng = 0x10000u;
    pend = 0;
  }

  // renormalise (LO0, r1) and emit (bac.c:112-137 + OutputNextBits, bac.c:93-105).
  // The writer's 64-bit accumulator may hold up to 64 bits here; the CALLER drains it (w.drain()) every few symbols,
  // at a warp-uniform point, so that the data-dependent "word is full" branch is not taken per symbol.
  __device__ __forceinline__ void renorm_emit(uint32_t LO0, uint32_t r1, BitWriter &w)
  {
    uint32_t n1, n3, P1;
    bac_renorm<false>(LO0, r1, LO, rng, n1, n3, P1);
    const bool some = n1 != 0;
    const uint32_t k = some ? n1 + pend : 0u;   // bits leaving now: n1 plus the pending run spliced after the first
    const uint32_t p0 = pend;
    pend = (some ? 0u : pend) + n3;
    // One rare region per symbol: k >= 32 (more than 16 consecutive E3 steps) or the accumulator would overflow
    // (nacc + k > 64 with k < 32 implies k + nacc/2 > 33).
    if (k + (w.nacc >> 1) >= 32)
    {
      w.drain();
      if (k > 32)
      {
        const uint32_t first = LO0 >> 31;
        w.put(first, 1);
        w.run(first ^ 1u, p0);
        w.put(shr_c(LO0 << 1, 33 - n1), n1 - 1);
      }
      else
        w.put(shr_c(LO0, 32 - n1) + shl_c(1u, k - 1) - shl_c(1u, n1 - 1), k);
      return;
    }
    // top n1 bits of lo, with the pending run spliced in: bits + 2^(k-1) - 2^(n1-1); k <= 31 here
    const uint32_t Pk = 1u << k;
    w.acc = (w.acc << k) | (uint64_t)(__umulhi(LO0, P1) + ((Pk - P1) >> 1));
    w.nacc += k;
  }

  // EncodeSymbol for a data bit (bac.c:107-111)
  __device__ __forceinline__ void encode_bit(bool is_top, uint32_t c1, const DivEntry e, BitWriter &w)
  {
    const uint32_t A = div_magic(rng * c1, e);
    const uint32_t B = div_magic(rng, e);
    const uint32_t dlo = is_top ? A : B;
    const uint32_t base = is_top ? rng : A;
    renorm_emit(LO + (dlo << 16), base - dlo, w);
  }

  // EOF symbol (bac.c:163) followed by FinishEncoding (bac.c:141-145)
  __device__ __forceinline__ void finish(const DivEntry e, BitWriter &w)
  {
    renorm_emit(LO, div_magic(rng, e), w);
    w.drain();
    const uint32_t bit = LO < 0x40000000u ? 0u : 1u;
    w.put(bit, 1);
    w.run(bit ^ 1u, (uint64_t)pend + 1);
    pend = 0;
  }
};

}  // namespace dc
