// bac.cuh -- device-side core of the binary arithmetic coder (reference: DCLib/src/bac.c).
//
// The reference is the Witten/Neal/Cleary coder with RANGE_BITS = 16 over the three symbols {0, 1, EOF}
// (bac.c:13,29-31).  This restatement keeps the same integer results but re-expresses the state so that one
// symbol costs a short, branch-free instruction sequence per thread (the kernels are integer-issue bound):
//
//   state     reference keeps (start_range, end_range, next_bits) (bac.c:83-84).  We keep (LO, rng):
//             rng = end - start + 1 (every renormalisation step doubles it whatever branch it takes) and LO = the
//             TRUE low end of the interval, top-aligned in a 32-bit register (bits 31..16, low half zero), i.e. the
//             16-bit window of the exact, arbitrarily long binary fraction that the coder is narrowing down.
//   pending   the reference never lets a carry reach bits it has already written: in the "E3" case
//             (lo = 01.., hi = 10..) it subtracts 0x4000, counts a pending bit and later writes the run inverted
//             (bac.c:93-105,124-129).  In the exact-fraction view that is nothing but carry propagation: while
//             bits are pending, the true window is the reference's window with the top bit flipped
//             (lo_true = lo ^ 0x8000, same for hi), the bits that left the window are 0 1 1 1 ..., and the
//             reference's "lo >= 0x8000 -> write 1 then zeros" is the carry rippling through that run.  So here the
//             window simply overflows into the bits shifted out before (a0 / the CarryWriter accumulator, and --
//             about once in 2^32 flushes -- into words already stored, which are private to the thread).  Output
//             bits are identical to the reference's; there is no pending counter and no splice on the hot path.
//   counts    the number of doublings per symbol is the reference's: n1 = number of equal leading bits of
//             (lo, hi) "E1/E2" steps, then n3 = length of the run "lo = .1.., hi = .0.." below the first differing bit
//             "E3" steps (after an E3 step lo < 0x8000 <= hi, so no E1/E2 step can follow within the same symbol).
//             Both windows carry the same flipped top bit while bits are pending, so lo ^ hi and the bits below the
//             top are the same in both views: one CLZ each, with HI = (hi << 16) | 0xFFFF capping them (lo == hi).
//   model     the reference keeps freq[4]/cum[4]/two index maps (bac.c:33-37).  EOF never moves and keeps
//             frequency 1 (UpdateModel is never called with index 3; halving maps 1 -> 1, bac.c:63) and index 0 is
//             a sentinel, so the whole adaptive model is (f_top, f_oth, top): the counts of the data symbols at
//             index 1 / 2 and which bit sits at index 1.  cum[] = {f_top + f_oth + 1, f_oth + 1, 1, 0}.
//   update    index 1: hi unchanged, lo += A                    A = floor(rng * cum[1] / cum[0])
//             index 2: lo += B, rng = A - B                      B = floor(rng / cum[0])
//             EOF    : rng = B                                    (bac.c:109-111 with the cum[] above)
//   division  total <= 16383 and rng * cum[1] <= 2^30, so floor(x / total) == umulhi(x, m) >> s with a
//             precomputed round-up reciprocal m per divisor (kernels.h: DivEntry) and s = ceil(log2 total) - 2; exact, no
//             correction step.  rng / total itself (rng <= 2^16) needs no shift: umulhi(rng, mb), mb = ceil(2^32 / total).
//   end       EOF symbol, then FinishEncoding (bac.c:141-145): one more pending bit, then the bit
//             (lo < 0x4000 ? 0 : 1) followed by the inverted run -- in the exact view: add 0x4000 to the window
//             (with carry) and shift two more bits out.
#pragma once

#include <stdint.h>

#include "bitio.cuh"
#include "kernels.h"

namespace dc {

static const uint32_t kBacMaxFreq = 0x3FFFu;   // MAX_FREQUENCY, bac.c:27

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

// Position of the leading one, 0xFFFFFFFF (-1) for zero: bfind.u32 is a single FLO.U32
#ifdef DC_HOST_SIM
static inline uint32_t bfind_u32(uint32_t a) { return a ? 31u - (uint32_t)__builtin_clz(a) : 0xFFFFFFFFu; }
#else
__device__ __forceinline__ uint32_t bfind_u32(uint32_t a)
{
  uint32_t r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}
#endif

// shift that goes with DivEntry::m of divisor d >= 3: ceil(log2 d) - 2
__device__ __forceinline__ uint32_t div_shift(uint32_t d) { return 30u - clz_nz(d - 1u); }
// floor(x / d) for x < 2^30 with the table entry of d and sh = div_shift(d)
__device__ __forceinline__ uint32_t div_magic(uint32_t x, const DivEntry e, uint32_t sh) { return __umulhi(x, e.m) >> sh; }
// floor(x / d) for x <= 2^16
__device__ __forceinline__ uint32_t div_small(uint32_t x, const DivEntry e) { return __umulhi(x, e.mb); }

struct BacModel
{
  uint32_t f_top;   // freq[1]
  uint32_t f_oth;   // freq[2]; cum[1] = f_oth + 1, cum[0] = f_top + f_oth + 1
  uint32_t top;     // index_to_symbol[1] as 0 / 0xFFFFFFFF

  __device__ __forceinline__ void init()   // InitModel, bac.c:39-52
  {
    f_top = 1;
    f_oth = 1;
    top = 0;
  }

  __device__ __forceinline__ uint32_t total() const { return f_top + f_oth + 1; }   // cum[0]

  // bac.c:57-67: halve, rounding up; EOF stays at 1.  Only when total() == kBacMaxFreq.
  __device__ __forceinline__ void halve()
  {
    f_top = (f_top + 1) >> 1;
    f_oth = (f_oth + 1) >> 1;
  }

  // UpdateModel, bac.c:68-80, for a data symbol that sat at index 1 (is_top) or 2.  Returns true when the two data
  // symbols traded places (bac.c:68-77: a symbol at index 2 moves left past its equal-frequency neighbour,
  // freq[2] == freq[1], and then counts as freq[1]).  That is rare on skewed data, so it is the only branch; the
  // common case is one predicated increment.
  __device__ __forceinline__ bool bump(bool is_top)
  {
    if (!is_top && f_oth == f_top)
    {
      top = ~top;
      f_top++;
      return true;
    }
    if (is_top)
      f_top++;
    else
      f_oth++;
    return false;
  }
};

// ---------------------------------------------------------------------------------------------------
// Hot-path form of "which symbol, its sub-interval, UpdateModel" for the unrolled symbol groups (no halving inside).
// Written as predicated PTX because the compiler's own if-conversion of the C++ above costs three times as many
// instructions; results are those of BacModel::bump.  In: A, B (see "update" at the top), rng, and
//   encoder: x = input word ^ top, mask = the symbol's bit; a set bit means the symbol sits at index 2.  When the
//            symbols trade places x is inverted, so it stays "word ^ top".
//   decoder: o = value - lo; the symbol sits at index 2 when o < A.  nontop collects those bits, inv collects `top`
//            for every position of the group (it changes for all later positions when the symbols trade places).
// Out: dlo = what is added to lo, base - dlo = the new range.
// SWAP = false is for groups in which the symbols cannot trade places (freq[1] - freq[2] >= group length at the
// start of the group): the update is then a single predicated increment of f_oth, and the caller adds the number
// of index-1 symbols to f_top after the group.  That form (bac_take_*) also avoids the two selects: a select, like a
// shift or a logic operation, takes the scheduler two cycles on this part while an add, a compare or a predicated
// move takes one (profiles/r2_ubench_pipes.txt), so it computes r1 = rng - A, then on the rarer symbol overwrites it
// with A - B and moves B over A.
template <bool ADAPTIVE, bool SWAP>
__device__ __forceinline__ void bac_pick_enc(uint32_t &x, uint32_t mask, uint32_t A, uint32_t B, uint32_t rng, BacModel &m, uint32_t &dlo,
                                             uint32_t &base)
{
#ifdef DC_HOST_SIM
  const bool is_top = (x & mask) == 0;
  dlo = is_top ? A : B;
  base = is_top ? rng : A;
  const uint32_t top0 = m.top, f_top0 = m.f_top;   // the caller derives `top` from x at the end of the group
  if (ADAPTIVE && m.bump(is_top))
    x = ~x;
  m.top = top0;
  if (!SWAP)
    m.f_top = f_top0;
#else
  if (ADAPTIVE && SWAP)
    asm("{\n\t"
        ".reg .pred pn, ps, pi;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %4, %5;\n\t"
        "setp.ne.u32 pn, t, 0;\n\t"
        "selp.u32 %0, %7, %6, pn;\n\t"
        "selp.u32 %1, %6, %8, pn;\n\t"
        "setp.eq.and.u32 ps, %3, %2, pn;\n\t"
        "@ps not.b32 %4, %4;\n\t"
        "not.pred pi, pn;\n\t"
        "or.pred pi, pi, ps;\n\t"
        "@pi add.u32 %2, %2, 1;\n\t"
        "@!pi add.u32 %3, %3, 1;\n\t"
        "}"
        : "=&r"(dlo), "=&r"(base), "+r"(m.f_top), "+r"(m.f_oth), "+r"(x)
        : "r"(mask), "r"(A), "r"(B), "r"(rng));
  else if (ADAPTIVE)
    asm("{\n\t"
        ".reg .pred pn;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %3, %4;\n\t"
        "setp.ne.u32 pn, t, 0;\n\t"
        "selp.u32 %0, %6, %5, pn;\n\t"
        "selp.u32 %1, %5, %7, pn;\n\t"
        "@pn add.u32 %2, %2, 1;\n\t"
        "}"
        : "=&r"(dlo), "=&r"(base), "+r"(m.f_oth)
        : "r"(x), "r"(mask), "r"(A), "r"(B), "r"(rng));
  else
    asm("{\n\t"
        ".reg .pred pn;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 pn, t, 0;\n\t"
        "selp.u32 %0, %5, %4, pn;\n\t"
        "selp.u32 %1, %4, %6, pn;\n\t"
        "}"
        : "=&r"(dlo), "=&r"(base)
        : "r"(x), "r"(mask), "r"(A), "r"(B), "r"(rng));
#endif
}

// In: A, B, rng; the symbol sits at index 2 when (x & mask) != 0.  Out: A = what is added to lo, r1 = the new range.
template <bool ADAPTIVE>
__device__ __forceinline__ void bac_take_enc(uint32_t x, uint32_t mask, uint32_t &A, uint32_t B, uint32_t rng, BacModel &m, uint32_t &r1)
{
#ifdef DC_HOST_SIM
  const bool nontop = (x & mask) != 0;
  r1 = nontop ? A - B : rng - A;
  if (nontop)
    A = B;
  if (ADAPTIVE && nontop)
    m.f_oth++;
#else
  if (ADAPTIVE)
    asm("{\n\t"
        ".reg .pred pn;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %3, %4;\n\t"
        "setp.ne.u32 pn, t, 0;\n\t"
        "sub.u32 %0, %6, %1;\n\t"
        "@pn sub.u32 %0, %1, %5;\n\t"
        "@pn mov.u32 %1, %5;\n\t"
        "@pn add.u32 %2, %2, 1;\n\t"
        "}"
        : "=&r"(r1), "+r"(A), "+r"(m.f_oth)
        : "r"(x), "r"(mask), "r"(B), "r"(rng));
  else
    asm("{\n\t"
        ".reg .pred pn;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 pn, t, 0;\n\t"
        "sub.u32 %0, %5, %1;\n\t"
        "@pn sub.u32 %0, %1, %4;\n\t"
        "@pn mov.u32 %1, %4;\n\t"
        "}"
        : "=&r"(r1), "+r"(A)
        : "r"(x), "r"(mask), "r"(B), "r"(rng));
#endif
}

// Decoder: the symbol sits at index 2 when o < A; `nontop` collects those positions.
template <bool ADAPTIVE>
__device__ __forceinline__ void bac_take_dec(uint32_t o, uint32_t bit, uint32_t &A, uint32_t B, uint32_t rng, BacModel &m, uint32_t &nontop, uint32_t &r1)
{
#ifdef DC_HOST_SIM
  const bool is_nontop = o < A;
  r1 = is_nontop ? A - B : rng - A;
  if (is_nontop)
  {
    A = B;
    nontop |= bit;
    if (ADAPTIVE)
      m.f_oth++;
  }
#else
  if (ADAPTIVE)
    asm("{\n\t"
        ".reg .pred pn;\n\t"
        "setp.lt.u32 pn, %4, %1;\n\t"
        "sub.u32 %0, %7, %1;\n\t"
        "@pn sub.u32 %0, %1, %6;\n\t"
        "@pn mov.u32 %1, %6;\n\t"
        "@pn or.b32 %3, %3, %5;\n\t"
        "@pn add.u32 %2, %2, 1;\n\t"
        "}"
        : "=&r"(r1), "+r"(A), "+r"(m.f_oth), "+r"(nontop)
        : "r"(o), "r"(bit), "r"(B), "r"(rng));
  else
    asm("{\n\t"
        ".reg .pred pn;\n\t"
        "setp.lt.u32 pn, %3, %1;\n\t"
        "sub.u32 %0, %6, %1;\n\t"
        "@pn sub.u32 %0, %1, %5;\n\t"
        "@pn mov.u32 %1, %5;\n\t"
        "@pn or.b32 %2, %2, %4;\n\t"
        "}"
        : "=&r"(r1), "+r"(A), "+r"(nontop)
        : "r"(o), "r"(bit), "r"(B), "r"(rng));
#endif
}

template <bool ADAPTIVE, bool SWAP>
__device__ __forceinline__ void bac_pick_dec(uint32_t o, uint32_t bit, uint32_t A, uint32_t B, uint32_t rng, BacModel &m, uint32_t &nontop,
                                             uint32_t &inv, uint32_t &dlo, uint32_t &base)
{
#ifdef DC_HOST_SIM
  const bool is_top = o >= A;
  dlo = is_top ? A : B;
  base = is_top ? rng : A;
  if (!is_top)
    nontop |= bit;
  const uint32_t top0 = m.top, f_top0 = m.f_top;   // the caller derives `top` from inv at the end of the group
  if (ADAPTIVE && m.bump(is_top))
    inv ^= bit - 1;
  m.top = top0;
  if (!SWAP)
    m.f_top = f_top0;
#else
  if (ADAPTIVE && SWAP)
    asm("{\n\t"
        ".reg .pred pn, ps, pi;\n\t"
        "setp.lt.u32 pn, %6, %8;\n\t"
        "selp.u32 %0, %9, %8, pn;\n\t"
        "selp.u32 %1, %8, %10, pn;\n\t"
        "@pn or.b32 %4, %4, %7;\n\t"
        "setp.eq.and.u32 ps, %3, %2, pn;\n\t"
        "@ps xor.b32 %5, %5, %11;\n\t"
        "not.pred pi, pn;\n\t"
        "or.pred pi, pi, ps;\n\t"
        "@pi add.u32 %2, %2, 1;\n\t"
        "@!pi add.u32 %3, %3, 1;\n\t"
        "}"
        : "=&r"(dlo), "=&r"(base), "+r"(m.f_top), "+r"(m.f_oth), "+r"(nontop), "+r"(inv)
        : "r"(o), "r"(bit), "r"(A), "r"(B), "r"(rng), "r"(bit - 1));
  else if (ADAPTIVE)
    asm("{\n\t"
        ".reg .pred pn;\n\t"
        "setp.lt.u32 pn, %4, %6;\n\t"
        "selp.u32 %0, %7, %6, pn;\n\t"
        "selp.u32 %1, %6, %8, pn;\n\t"
        "@pn or.b32 %3, %3, %5;\n\t"
        "@pn add.u32 %2, %2, 1;\n\t"
        "}"
        : "=&r"(dlo), "=&r"(base), "+r"(m.f_oth), "+r"(nontop)
        : "r"(o), "r"(bit), "r"(A), "r"(B), "r"(rng));
  else
    asm("{\n\t"
        ".reg .pred pn;\n\t"
        "setp.lt.u32 pn, %3, %5;\n\t"
        "selp.u32 %0, %6, %5, pn;\n\t"
        "selp.u32 %1, %5, %7, pn;\n\t"
        "@pn or.b32 %2, %2, %4;\n\t"
        "}"
        : "=&r"(dlo), "=&r"(base), "+r"(nontop)
        : "r"(o), "r"(bit), "r"(A), "r"(B), "r"(rng));
#endif
}

// Number of doublings (bac.c:112-137) of the interval (LO0, r1), PLUS ONE: LO0 = top-aligned low end, r1 = range.
// n1 = leading bits on which lo and hi agree; below the first differing bit (lo 0, hi 1) the E3 run is where lo has 1
// and hi has 0, i.e. where z = ~lo | hi is clear.  Masking z down to the bits below the first differing one leaves its
// leading one at position n1 + 1 + n3.  The low 16 guard ones of z and of lo ^ hi bound both counts (n <= 16) for ANY
// input, so a caller that lost bits (see the group redo in dega_device.cuh) cannot get a wild shift count.
__device__ __forceinline__ uint32_t bac_doublings_p1(uint32_t LO0, uint32_t r1)
{
  const uint32_t HI = LO0 + (r1 << 16) - 1;          // (hi << 16) | 0xFFFF; may wrap while bits are pending
  const uint32_t n1 = clz_nz(LO0 ^ HI);              // 0..16 (the low 16 bits of the xor are ones)
  const uint32_t u = (~LO0 | HI) & (0x7FFFFFFFu >> n1);
  return clz_nz(u);                                  // n + 1, 1..17
}
__device__ __forceinline__ uint32_t bac_doublings(uint32_t LO0, uint32_t r1) { return bac_doublings_p1(LO0, r1) - 1; }

// ---------------------------------------------------------------------------------------------------
// Output of the encoder: MSB-first bit stream whose tail can still take a carry.
//   acc   low nacc bits = bits not yet stored; bit nacc (at most one, see below) = a carry that still has to be
//         added to what was stored before; nothing above.
// The interval only ever shrinks, so the bits in front of any position change at most once (by +1) after they were
// produced: one carry bit per flush is all there is, and the ripple into stored words stops at the first word that
// does not wrap.
// +1 on the big-endian number formed by the stored words (rare: the tail of the stored bits was all ones).
// Out of line and by value so that the writer's state never has to live in local memory.
__device__ __noinline__ static void bac_ripple(uint32_t *words, uint32_t nwords)
{
  for (uint32_t i = nwords; i-- > 0;)
  {
    const uint32_t v = bswap32(words[i]) + 1u;
    words[i] = bswap32(v);
    if (v != 0)
      break;
  }
}

struct CarryWriter
{
  uint32_t *words;
  uint32_t cap_words;
  uint32_t nwords;
  uint32_t nacc;      // < 32 between calls
  uint32_t acc;       // the nacc bits not stored yet, and at bit nacc a carry that still has to reach what was stored before
  bool overflow;      // valid after finish()

  __device__ __forceinline__ void init(uint8_t *p, uint64_t cap_bytes)
  {
    words = reinterpret_cast<uint32_t *>(p);
    cap_words = (uint32_t)(cap_bytes >> 2);
    nwords = 0;
    nacc = 0;
    acc = 0;
    overflow = false;
  }

  __device__ __forceinline__ void ripple()
  {
    if (nwords <= cap_words)   // else the stream has outgrown its region: it is reported as overflow and redone
      bac_ripple(words, nwords);
  }

  // appends the low k bits of v (k <= 31); bit k of v, if set, is a carry into the bits appended earlier
  __device__ __forceinline__ void push(uint32_t v, uint32_t k)
  {
    const uint64_t c = ((uint64_t)acc << k) + v;   // at most 62 bits and the carry above them
    const uint32_t tot = nacc + k;
    if (tot >= 32)
    {
      const uint32_t r = tot - 32;                 // bits that stay behind: 0..30
      const uint64_t t = c >> r;                   // 32 bits to store + the carry bit
      acc = (uint32_t)c & ~(0xFFFFFFFFu << r);
      nacc = r;
      if ((uint32_t)(t >> 32))
        ripple();
      if (nwords < cap_words)
        words[nwords] = bswap32((uint32_t)t);
      nwords++;
    }
    else
    {
      acc = (uint32_t)c;                           // fits: tot bits and the carry at bit tot <= 31
      nacc = tot;
    }
  }

  __device__ __forceinline__ uint64_t bits() const { return (uint64_t)nwords * 32 + nacc; }

  // resolves a carry that is still waiting, stores the fractional word (left aligned, zero padded)
  __device__ __forceinline__ uint64_t finish()
  {
    if (acc >> nacc)
    {
      ripple();
      acc &= ~(0xFFFFFFFFu << nacc);
    }
    const uint64_t total = bits();
    overflow = nwords + (nacc ? 1u : 0u) > cap_words;
    if (nacc && nwords < cap_words)
      words[nwords] = bswap32(acc << (32 - nacc));
    return total;
  }
};

struct BacEncoder
{
  uint32_t LO;    // true low end << 16
  uint32_t rng;   // end_range - start_range + 1, in [1, 65536]
  uint32_t a0;    // bits that left LO since the last drain: low kg bits, plus a carry at bit kg
  uint32_t kg;    // <= 31

  __device__ __forceinline__ void init()   // StartEncoding, bac.c:86-91
  {
    LO = 0;
    rng = 0x10000u;
    a0 = 0;
    kg = 0;
  }

  // hands the group's bits to the writer; called at warp-uniform points (every 8 symbols) and whenever a symbol
  // would push kg past 31 (rare: a load profile yields < 1 bit per symbol and at most 4 per symbol in practice)
  __device__ __forceinline__ void drain(CarryWriter &w)
  {
    w.push(a0, kg);
    a0 = 0;
    kg = 0;
  }

  // interval update lo += dlo, range = r1 (bac.c:109-111), renormalisation and output (bac.c:112-137)
  __device__ __forceinline__ void narrow(uint32_t dlo, uint32_t r1, CarryWriter &w)
  {
    const uint64_t x = (((uint64_t)a0 << 32) | LO) + ((uint64_t)dlo << 16);   // the window overflows into a0
    const uint32_t LO0 = (uint32_t)x;
    a0 = (uint32_t)(x >> 32);
    const uint32_t n = bac_doublings(LO0, r1);
    if (kg + n > 31)
      drain(w);
    kg += n;
    rng = r1 << n;
    a0 = __funnelshift_l(LO0, a0, n);
    LO = LO0 << n;
  }

  // Same without the capacity test: kg1 counts n + 1 per symbol and the caller checks afterwards that the total stayed
  // within a0.  This was the form of the unrolled groups until narrow_w() below replaced it there; it stays as the
  // plain statement of the step that narrow_w() is tested against (tests/host_sim: sim_check_narrow_w).
  __device__ __forceinline__ void narrow_fast(uint32_t dlo, uint32_t r1, uint32_t &kg1)
  {
    const uint64_t x = (((uint64_t)a0 << 32) | LO) + ((uint64_t)dlo << 16);
    const uint32_t LO0 = (uint32_t)x;
    const uint32_t c = bac_doublings_p1(LO0, r1);
    const uint32_t n = c - 1;
    kg1 += c;
    rng = r1 << n;
    a0 = __funnelshift_l(LO0, (uint32_t)(x >> 32), n);
    LO = LO0 << n;
  }

  // The unrolled groups of the encoder keep the interval in ONE register W instead of (a0, LO): bit 0 a guard (zero in
  // W, one in the high end), bits 16..1 the 16-bit window, bits 31..17 the bits that left the window since they were last
  // taken out -- the carry of lo += dlo runs into them by itself and a doubling is one shift.  That costs a 1-cycle add
  // and a shift per symbol where (a0, LO) takes a 64-bit multiply-add and a funnel shift; the price is room for only 14
  // bits (plus the carry), so the caller takes the bits out every half group and checks the counts afterwards.
  //   H = 2 hi + 1; first differing bit b = leading one of (W ^ H) & 0x1FFFF (the guards differ: never zero);
  //   u = (~W | H) below b; its leading one q (-1: none, lo == hi) is the first bit under the E3 run, n = 15 - q.
  // q is in [-1, 15] for ANY input, so a group that lost bits cannot produce a wild shift count.
  static __device__ __forceinline__ uint32_t to_w(uint32_t LO) { return LO >> 15; }
  static __device__ __forceinline__ uint32_t from_w(uint32_t W) { return W << 15; }
  __device__ __forceinline__ void narrow_w(uint32_t &W, uint32_t dlo, uint32_t r1, uint32_t &qsum)
  {
    W += dlo << 1;
    const uint32_t H = W + (r1 << 1) - 1u;
    const uint32_t cz = clz_nz((W ^ H) & 0x1FFFFu);
    const uint32_t q = bfind_u32((~W | H) & (0x7FFFFFFFu >> cz));
    const uint32_t n = 15u - q;
    qsum += q;
    rng = r1 << n;
    W <<= n;
  }

  // EncodeSymbol for a data bit (bac.c:107-111); c1m1 = cum[1] - 1
  __device__ __forceinline__ void encode_bit(bool is_top, uint32_t c1m1, const DivEntry e, uint32_t sh, CarryWriter &w)
  {
    const uint32_t A = div_magic(rng * c1m1 + rng, e, sh);
    const uint32_t B = div_small(rng, e);
    const uint32_t dlo = is_top ? A : B;
    const uint32_t base = is_top ? rng : A;
    narrow(dlo, base - dlo, w);
  }

  // EOF symbol (bac.c:163) followed by FinishEncoding (bac.c:141-145)
  __device__ __forceinline__ void finish(const DivEntry e, CarryWriter &w)
  {
    narrow(0, div_small(rng, e), w);
    drain(w);
    const uint64_t x = (((uint64_t)a0 << 32) | LO) + 0x40000000ull;
    a0 = (uint32_t)(x >> 32);
    a0 = (a0 << 2) | ((uint32_t)x >> 30);
    kg = 2;
    drain(w);
  }
};

}  // namespace dc
