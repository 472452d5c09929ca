// bac.cuh -- device-side core of the binary arithmetic coder (reference: DCLib/src/bac.c).
//
// The reference is the Witten/Neal/Cleary coder with RANGE_BITS = 16 over the three symbols {0, 1, EOF}
// (bac.c:13,29-31).  This restatement keeps the same integer results but re-expresses the state so that one
// symbol costs a fixed, branch-light instruction sequence per thread:
//
//   state     reference keeps (start_range, end_range) (bac.c:83); we keep (lo, rng) with rng = hi - lo + 1,
//             because every renormalisation step doubles rng whatever branch it takes.
//   model     the reference keeps freq[4]/cum[4]/two index maps (bac.c:33-37).  EOF never moves and keeps
//             frequency 1 (UpdateModel is never called with index 3; halving maps 1 -> 1, bac.c:63), and index 0
//             is a sentinel, so the whole adaptive model is (f_top, f_oth, top): the counts of the data symbol at
//             index 1 / index 2 and which bit sits at index 1.  cum[] = {f_top + f_oth + 1, f_oth + 1, 1, 0}.
//   update    index 1: hi unchanged, lo += A                    A = floor(rng * (f_oth + 1) / total)
//             index 2: lo += B, rng = A - B                      B = floor(rng / total)
//             EOF    : rng = B                                    (bac.c:109-111 with the cum[] above)
//   division  total <= 16383 and rng * cum < 2^30, so floor(x / total) == umulhi(x << 2, m) >> l with a
//             precomputed round-up reciprocal (m, l) per divisor (kernels.h: DivEntry); exact, no correction step.
//   renorm    bac.c:112-137 is a loop of three cases.  Closed form: first n1 = number of equal leading bits of
//             (lo, hi) "E1/E2" steps, each emitting that bit; then n3 = length of the run "lo = 01..., hi = 10..."
//             below the top bit, each an "E3" step (pending++).  After an E3 step lo < 0x8000 <= hi, so no E1/E2
//             step can follow within the same symbol.  Both counts come from one CLZ each.
#pragma once

#include <stdint.h>

#include "bitio.cuh"
#include "kernels.h"

namespace dc {

static const uint32_t kBacMaxFreq = 0x3FFFu;   // MAX_FREQUENCY, bac.c:27

struct BacModel
{
  uint32_t f_top;   // freq[1]
  uint32_t f_oth;   // freq[2]
  uint32_t top;     // index_to_symbol[1]
  uint32_t total;   // cum[0]

  __device__ __forceinline__ void init()   // InitModel, bac.c:39-52
  {
    f_top = 1;
    f_oth = 1;
    top = 0;
    total = 3;
  }

  // UpdateModel, bac.c:54-81, for a data symbol (bit b; is_top == (b == top) evaluated before the update)
  __device__ __forceinline__ void update(uint32_t b, bool is_top)
  {
    if (total == kBacMaxFreq)   // bac.c:57-67: halve, rounding up; EOF stays at 1
    {
      f_top = (f_top + 1) >> 1;
      f_oth = (f_oth + 1) >> 1;
      total = f_top + f_oth + 1;
    }
    if (is_top)
      f_top++;
    else if (f_oth == f_top)    // bac.c:68-77: the coded symbol moves left past its equal-frequency neighbour
    {
      top = b;
      f_top++;
    }
    else
      f_oth++;
    total++;
  }
};

// floor(x / d) for x < 2^30 with the table entry of d
__device__ __forceinline__ uint32_t div_magic(uint32_t x, const DivEntry e) { return __umulhi(x << 2, e.m) >> e.l; }

// Number of E1/E2 and E3 steps for the interval (lo, rng), and the renormalised interval (bac.c:112-137).
__device__ __forceinline__ void bac_renorm(uint32_t &lo, uint32_t &rng, uint32_t &n1, uint32_t &n3)
{
  const uint32_t hi = lo + rng - 1;
  n1 = __clz(lo ^ hi) - 16;              // equal leading bits of two 16-bit values (16 when lo == hi)
  const uint32_t lo1 = (lo << n1) & 0xFFFFu;
  const uint32_t rng1 = rng << n1;
  const uint32_t hi1 = lo1 + rng1 - 1;
  const uint32_t y = (~lo1 | hi1) & 0x7FFFu;   // bit clear where lo has 1 and hi has 0
  n3 = __clz((y << 17) | 0x10000u);
  lo = (lo1 << n3) & 0x7FFFu;
  rng = rng1 << n3;
}

struct BacEncoder
{
  uint32_t lo;      // start_range
  uint32_t rng;     // end_range - start_range + 1, in [1, 65536]
  uint64_t pend;    // next_bits, bac.c:84

  __device__ __forceinline__ void init()   // StartEncoding, bac.c:86-91
  {
    lo = 0;
    rng = 0x10000u;
    pend = 0;
  }

  // renormalise and emit (bac.c:112-137 + OutputNextBits, bac.c:93-105)
  __device__ __forceinline__ void renorm_emit(BitWriter &w)
  {
    const uint32_t lo0 = lo;
    uint32_t n1, n3;
    bac_renorm(lo, rng, n1, n3);
    if (n1)
    {
      if (pend == 0)
        w.put(lo0 >> (16 - n1), n1);
      else
      {
        const uint32_t first = lo0 >> 15;
        w.put(first, 1);
        w.run(first ^ 1u, pend);
        w.put((lo0 >> (16 - n1)) & ((1u << (n1 - 1)) - 1u), n1 - 1);
        pend = 0;
      }
    }
    pend += n3;
  }

  // EncodeSymbol for a data bit (bac.c:107-111)
  __device__ __forceinline__ void encode_bit(bool is_top, uint32_t c1, const DivEntry e, BitWriter &w)
  {
    const uint32_t A = div_magic(rng * c1, e);
    const uint32_t B = div_magic(rng, e);
    if (is_top)
    {
      lo += A;
      rng -= A;
    }
    else
    {
      lo += B;
      rng = A - B;
    }
    renorm_emit(w);
  }

  // EOF symbol (bac.c:163) followed by FinishEncoding (bac.c:141-145)
  __device__ __forceinline__ void finish(const DivEntry e, BitWriter &w)
  {
    rng = div_magic(rng, e);
    renorm_emit(w);
    pend++;
    const uint32_t bit = lo < 0x4000u ? 0u : 1u;
    w.put(bit, 1);
    w.run(bit ^ 1u, pend);
    pend = 0;
  }
};

}  // namespace dc
