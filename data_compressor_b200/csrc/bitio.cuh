// bitio.cuh -- device-side bit streams with the DCIOLib layout.
//
// DCIOLib packs bits MSB-first into bytes (DCIOLib/src/bit_file_buffer.c:183,244) and multi-bit values
// big-endian across bytes (bit_file_buffer.c:297-308).  Every bit stream that leaves a kernel is stored in
// exactly that byte order, so a stream can be handed to the reference (or to the next kernel) as plain
// bytes.  Kernels work on 32-bit words: word k of a stream holds stream bits [32k, 32k+32) and is stored
// byte-swapped (PRMT) so that memory order == stream order.
#pragma once

#include <stdint.h>

namespace dc {

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// A set of per-stream bit strings in one allocation.
//   byte_off != nullptr : stream s starts at base + byte_off[s] (densely packed payload, any alignment)
//   byte_off == nullptr : stream s starts at base + s * stride_bytes (scratch, stride multiple of 16)
//   interleave != 0     : scratch between two kernels of one pipeline.  The 32 streams of a warp share a region of
//                         32 * stride_bytes; word k of stream s sits at word (32 * k + s % 32) of the region of warp
//                         s / 32, so that the lanes of a warp, which move through their streams at about the same
//                         pace, touch the same 128-byte lines (coalesced loads and stores, whole lines leaving L2)
//                         instead of 32 sectors 4.6 KB apart.  The allocation covers round_up(n, 32) streams.
struct BitSrc
{
  const uint8_t *base;
  const uint64_t *byte_off;
  uint64_t stride_bytes;
  const uint64_t *nbits;
  uint32_t interleave;
};

struct BitDst
{
  uint8_t *base;
  uint64_t stride_bytes;   // per-stream capacity, multiple of 16
  uint64_t *nbits;
  uint32_t interleave;
};

// ---------------------------------------------------------------------------------------------------
// Writer: private to one thread, appends MSB-first.  Capacity overflow is recorded, never written.
struct BitWriter
{
  uint32_t *words;
  uint32_t cap_words;
  uint32_t nwords;
  uint32_t wstride;  // distance between consecutive words of the stream, in words (1, or 32 when interleaved)
  uint32_t nacc;     // pending bits in acc, always < 32 between calls
  uint64_t acc;      // low nacc bits are pending
  bool overflow;     // valid after finish(): the stream outgrew cap_words (nothing is ever written past the end)

  __device__ __forceinline__ void init(uint8_t *p, uint64_t cap_bytes, uint32_t word_stride = 1)
  {
    words = reinterpret_cast<uint32_t *>(p);
    cap_words = (uint32_t)(cap_bytes >> 2);
    wstride = word_stride;
    nwords = 0;
    nacc = 0;
    acc = 0;
    overflow = false;
  }

  __device__ __forceinline__ void flush_word()
  {
    nacc -= 32;
    const uint32_t w = (uint32_t)(acc >> nacc);
    if (nwords < cap_words)
      words[(uint64_t)nwords * wstride] = bswap32(w);
    nwords++;
  }

  // brings nacc back below 32 (the arithmetic coder lets it grow up to 64 between warp-uniform drain points)
  __device__ __forceinline__ void drain()
  {
    while (nacc >= 32)
      flush_word();
  }

  // n in [0, 32], v < 2^n; requires nacc < 32
  __device__ __forceinline__ void put(uint32_t v, uint32_t n)
  {
    acc = (acc << n) | v;
    nacc += n;
    if (nacc >= 32)
      flush_word();
  }

  // a whole word at a word boundary (nacc == 0): no trip through the accumulator
  __device__ __forceinline__ void put_word(uint32_t v)
  {
    if (nwords < cap_words)
      words[(uint64_t)nwords * wstride] = bswap32(v);
    nwords++;
  }

  // n in [0, 64], v < 2^n
  __device__ __forceinline__ void put64(uint64_t v, uint32_t n)
  {
    if (n > 32)
    {
      put((uint32_t)(v >> 32), n - 32);
      put((uint32_t)v, 32);
    }
    else
      put((uint32_t)v, n);
  }

  __device__ __forceinline__ void zeros(uint32_t n)
  {
    while (n > 32)
    {
      put(0u, 32);
      n -= 32;
    }
    put(0u, n);
  }

  // n copies of `bit` (n unbounded)
  __device__ __forceinline__ void run(uint32_t bit, uint64_t n)
  {
    const uint32_t pat = bit ? 0xFFFFFFFFu : 0u;
    while (n >= 32)
    {
      put(pat, 32);
      n -= 32;
    }
    if (n)
      put(pat >> (32 - (uint32_t)n), (uint32_t)n);
  }

  __device__ __forceinline__ uint64_t bits() const { return (uint64_t)nwords * 32 + nacc; }

  // Flushes the fractional word (left aligned, zero padded); returns the exact bit count.
  __device__ __forceinline__ uint64_t finish()
  {
    const uint64_t total = bits();
    overflow = nwords + (nacc ? 1u : 0u) > cap_words;
    if (nacc && nwords < cap_words)
      words[(uint64_t)nwords * wstride] = bswap32((uint32_t)(acc << (32 - nacc)));
    return total;
  }
};

// ---------------------------------------------------------------------------------------------------
// Reader: private to one thread.  The stream may start at any byte address; bits past `nbits` read as 0
// (the stream that follows in a dense payload is masked off).
struct BitReader
{
  const uint32_t *words;   // 4-byte aligned base
  uint32_t phase;          // 0, 8, 16, 24: bit offset of the stream inside words[0]
  uint64_t nbits;
  uint64_t pos;            // bits consumed
  uint64_t win;            // MSB-aligned look-ahead
  uint32_t wn;             // bits held in win
  uint32_t next;           // next 32-bit chunk to fetch
  uint32_t carry;          // BE(words[next]) when phase != 0
  uint32_t wstride;        // distance between consecutive words of the stream, in words (1, or 32 when interleaved)

  __device__ __forceinline__ void init(const uint8_t *p, uint64_t nbits_, uint32_t word_stride = 1)
  {
    wstride = word_stride;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    words = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    phase = (uint32_t)(a & 3) * 8;
    nbits = nbits_;
    pos = 0;
    win = 0;
    wn = 0;
    next = 0;
    carry = 0;
    if (phase != 0 && nbits != 0)
      carry = bswap32(__ldg(words));
  }

  // stream bits [32k, 32k+32), zero beyond nbits; k must be fetched in increasing order
  __device__ __forceinline__ uint32_t fetch32()
  {
    const uint64_t bitpos = (uint64_t)next * 32;
    uint32_t w = 0;
    if (bitpos < nbits)
    {
      const uint64_t remaining = nbits - bitpos;
      if (phase == 0)
        w = bswap32(__ldg(words + (uint64_t)next * wstride));
      else
      {
        uint32_t nx = 0;
        if (remaining > 32 - phase)            // the chunk really extends into the next aligned word (plain layout only)
          nx = bswap32(__ldg(words + next + 1));
        w = __funnelshift_l(nx, carry, phase);
        carry = nx;
      }
      if (remaining < 32)
        w &= ~(0xFFFFFFFFu >> (uint32_t)remaining);
    }
    next++;
    return w;
  }

  __device__ __forceinline__ void refill()
  {
    if (wn <= 32)
    {
      win |= (uint64_t)fetch32() << (32 - wn);
      wn += 32;
    }
  }

  __device__ __forceinline__ uint64_t avail() const { return pos < nbits ? nbits - pos : 0; }
  __device__ __forceinline__ bool eof() const { return pos >= nbits; }

  // n in [0, 32]; bits past the end read as zero, pos keeps counting
  __device__ __forceinline__ uint32_t take(uint32_t n)
  {
    refill();
    const uint32_t v = n ? (uint32_t)(win >> (64 - n)) : 0u;
    win <<= n;
    wn -= n;
    pos += n;
    return v;
  }

  // n in [0, 64]
  __device__ __forceinline__ uint64_t take64(uint32_t n)
  {
    if (n > 32)
    {
      const uint64_t hi = take(n - 32);
      return (hi << 32) | take(32);
    }
    return take(n);
  }

  // look at the next 32 bits without consuming
  __device__ __forceinline__ uint32_t peek32()
  {
    refill();
    return (uint32_t)(win >> 32);
  }
};

// ---------------------------------------------------------------------------------------------------
// Reader of the arithmetic decoder: one stream of a dense payload (any byte address, the next stream right behind it).
// Only what the decoder uses: "the 32 bits at the current position" and "move on by up to 32 bits", positions in 32 bits.
// Two consecutive aligned words of the payload are held (byte order undone, bits past the stream's end cleared when the
// word is loaded) with the number of bits of the first that are used up: a peek is one funnel shift, moving on an add and,
// when the first word is used up, one predicated load and byte swap.  Bits past the end read as zero (bac.c:173-183); the
// caller keeps counting them in `pos` and decides when that is an error.
struct DenseReader
{
  const uint32_t *words;   // 4-byte aligned base
  uint32_t span;           // aligned words the stream touches
  uint32_t endmask;        // the stream's bits in the last of them
  uint32_t idx;            // index of the aligned word in `nxt`
  uint32_t cur, nxt;
  uint32_t off;            // bits of `cur` used up: 0..31 between calls (starts at the stream's byte phase)
  uint32_t pos;            // bits consumed, counted past the end too

  __device__ __forceinline__ uint32_t aligned(uint32_t i) const
  {
    if (i >= span)
      return 0u;
    const uint32_t w = bswap32(__ldg(words + i));
    return i + 1 == span ? w & endmask : w;
  }

  __device__ __forceinline__ void init(const uint8_t *p, uint32_t nbits)
  {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    words = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    off = (uint32_t)(a & 3) * 8;
    const uint32_t end = off + nbits;   // the stream's end, in bits from words[0]
    span = (end + 31) >> 5;
    endmask = (end & 31) ? ~(0xFFFFFFFFu >> (end & 31)) : 0xFFFFFFFFu;
    cur = aligned(0);
    nxt = aligned(1);
    idx = 1;
    pos = 0;
  }

  // the next 32 bits of the stream
  __device__ __forceinline__ uint32_t peek32() const { return __funnelshift_l(nxt, cur, off); }

  // n <= 32
  __device__ __forceinline__ void advance(uint32_t n)
  {
    off += n;
    pos += n;
    if (off >= 32)
    {
      off -= 32;
      cur = nxt;
      idx++;
      nxt = aligned(idx);
    }
  }
};

__device__ __forceinline__ const uint8_t *src_ptr(const BitSrc &s, uint64_t i)
{
  if (s.interleave)
    return s.base + (i >> 5) * 32 * s.stride_bytes + (i & 31) * 4;
  return s.base + (s.byte_off ? s.byte_off[i] : i * s.stride_bytes);
}

__device__ __forceinline__ uint8_t *dst_ptr(const BitDst &d, uint64_t i)
{
  if (d.interleave)
    return d.base + (i >> 5) * 32 * d.stride_bytes + (i & 31) * 4;
  return d.base + i * d.stride_bytes;
}

// sign extension of the low vs bits (EXTEND_IO_INT_SIGN, DCLib/inc/io_macros.h:89-90)
__device__ __forceinline__ int64_t sext(uint64_t v, uint32_t vs)
{
  return vs == 64 ? (int64_t)v : ((int64_t)(v << (64 - vs)) >> (64 - vs));
}

__device__ __forceinline__ uint64_t low_mask(uint32_t vs) { return vs == 64 ? ~0ull : ((1ull << vs) - 1); }

}  // namespace dc
