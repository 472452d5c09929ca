// lzmh_device.cuh -- LZMH (DCLib/src/lzmh.c) device code.
//
// LZMH = byte-oriented LZ77 (128-byte window, match length 3..274, a 4-entry recent-offset cache with
// move-to-front) + "ATH": literals are coded with a static 19-entry prefix code indexed by the byte's RANK in a
// count-sorted list of up to 48 symbols (lzmh.c:54-106).  The parse is greedy and sequential; what parallelises
// inside a stream is the 128-offset match search.
//
//   encoder: one WARP per stream.  The 402-byte neighbourhood of the parse position lives in shared memory; each
//            lane measures the match length of 4 of the 128 offsets, a REDUX picks (longest, then smallest
//            offset) -- the same winner as the reference's ascending scan that keeps strictly longer matches only
//            (lzmh.c:196-214).  The rank list is searched/updated with ballots; lane 0 owns the bit writer.
//   decoder: one WARP per stream too: every lane follows the bit-serial parse, the work inside a token (prefix-code
//            match, rank-list update, match copy, output to global memory) is shared out over the lanes.
//
// The reference works on a 403-byte ring (lzmh.c:68); restated on linear positions the ring always holds
// min(pos,128) bytes of history and min(remaining,275) bytes of look-ahead, with one visible quirk: an input of
// exactly 403 bytes produces no output at all (lzmh.c:169-173).
#pragma once

#include <stdint.h>

#include "bitio.cuh"
#include "kernels.h"

namespace dc {

static const int kLzWindow = 128;   // LZ_MAX_OFFSET, lzmh.c:54
static const int kLzMaxLen = 274;   // LZ_MAX_LENGTH, lzmh.c:56
static const int kLzRing = 403;     // INTERN_BUFFER_LENGTH, lzmh.c:68
static const int kLzList = 48;      // HUFF_LIST_LENGTH, lzmh.c:61
static const int kLzTree = 19;      // TREE_LENGHT, lzmh.c:85

// lzmh.c:86-106
__device__ __forceinline__ uint32_t lz_tree_len(int rank) { return rank < 4 ? 4u : rank < 9 ? 5u : rank < 13 ? 6u : rank < 15 ? 7u : 8u; }
__device__ __forceinline__ uint32_t lz_tree_code(int rank)
{
  // 1111,1110,1101,1100 | 10111..10011 | 100101..100010 | 1000011,1000010 | 10000011..10000000
  return rank < 4 ? 0x0Fu - rank : rank < 9 ? 0x17u - (rank - 4) : rank < 13 ? 0x25u - (rank - 9) : rank < 15 ? 0x43u - (rank - 13) : 0x83u - (rank - 15);
}

// Offset + length token (lzmh.c:216-278) given the recent-offset slot k (4 = not cached)
__device__ __forceinline__ void lz_token(int k, int off, int len, uint32_t &code, uint32_t &nbits)
{
  switch (k)
  {
    case 0: code = 0x06; nbits = 4; break;    // 0110
    case 1: code = 0x0E; nbits = 5; break;    // 01110
    case 2: code = 0x1E; nbits = 6; break;    // 011110
    case 3: code = 0x1F; nbits = 6; break;    // 011111
    default: code = 0x100u | (uint32_t)(off - 1); nbits = 10; break;   // 010 + 7 bit
  }
  if (len < 11)
  {
    code = (code << 4) | (uint32_t)(len - 3);
    nbits += 4;
  }
  else if (len < 19)
  {
    code = (code << 5) | 0x10u | (uint32_t)(len - 11);
    nbits += 5;
  }
  else
  {
    code = (code << 10) | 0x300u | (uint32_t)(len - 19);
    nbits += 10;
  }
}

#ifndef DC_HOST_SIM
// ---------------------------------------------------------------------------------------------------------
// Encoder, one warp per stream.  `win` is this warp's shared-memory window, `lsym`/`lcnt` its rank list.
static const int kLzWin = 2048;   // bytes of the stream kept in shared memory per warp (1024: 3.6 % slower, 4096: no better)
static const int kLzWinPad = 16;  // readable slack behind the window (word-wise match compare)

// shared memory through 32-bit addresses (the generic-pointer form recomputes the window base in the token loop)
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// four bytes starting at byte `idx` of a 4-byte aligned shared-memory array, first byte in the low bits
__device__ __forceinline__ uint32_t lz_load32(const uint8_t *base, uint32_t idx)
{
  const uint32_t *w = reinterpret_cast<const uint32_t *>(base) + (idx >> 2);
  return __funnelshift_r(w[0], w[1], (idx & 3) * 8);
}

// win[0 .. kLzWin) <- src[wbase ..), as far as the stream goes: whole words when the source is aligned (eight loads in
// flight per lane), bytes otherwise and for the last few
__device__ __forceinline__ void lz_fill_window(uint8_t *win, const uint8_t *src, uint32_t wbase, uint32_t n, uint32_t lane)
{
  const uint32_t avail = n - wbase < (uint32_t)kLzWin ? n - wbase : (uint32_t)kLzWin;
  uint32_t done = 0;
  if ((reinterpret_cast<uintptr_t>(src + wbase) & 3) == 0)
  {
    const uint32_t nw = avail >> 2;
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src + wbase);
    uint32_t *w32 = reinterpret_cast<uint32_t *>(win);
    uint32_t v[kLzWin / 128];
#pragma unroll
    for (int k = 0; k < kLzWin / 128; k++)
      v[k] = lane + 32 * k < nw ? __ldg(s32 + lane + 32 * k) : 0u;
#pragma unroll
    for (int k = 0; k < kLzWin / 128; k++)
      if (lane + 32 * k < nw)
        w32[lane + 32 * k] = v[k];
    done = nw << 2;
  }
  for (uint32_t i = done + lane; i < avail; i += 32)
    win[i] = src[wbase + i];
}

__device__ inline void lzmh_encode_warp(const LzmhArgs &a, const uint64_t s, uint8_t *win, uint8_t *lsym, uint32_t *lcnt)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint8_t *src = a.in + a.in_off[s];
  const uint32_t n = (uint32_t)(a.in_off[s + 1] - a.in_off[s]);
  BitWriter w;
  w.init(dst_ptr(a.bits_out, s), a.bits_out.stride_bytes);
  const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win);

  if (n != (uint32_t)kLzRing)   // lzmh.c:169-173: the initial fill of exactly 403 bytes wraps write == read at EOF
  {
    int recent0 = 0, recent1 = 0, recent2 = 0, recent3 = 0;   // lzmh.c:157-159
    uint32_t nlist = 0;
    uint32_t wbase = 0;
    lz_fill_window(win, src, 0, n, lane);
    __syncwarp();
    uint32_t pos = 0;
    while (pos < n)   // lzmh.c:173
    {
      if (pos + kLzMaxLen > wbase + kLzWin && wbase + kLzWin < n)   // slide the window: keep 128 bytes of history
      {
        __syncwarp();
        wbase = (pos - kLzWindow) & ~3u;   // word steps keep the alignment of the stream
        lz_fill_window(win, src, wbase, n, lane);
        __syncwarp();
      }
      const uint32_t p = pos - wbase;
      const int maxoff = pos < (uint32_t)kLzWindow ? (int)pos : kLzWindow;            // lzmh.c:175-179
      const int maxlen = n - pos > (uint32_t)kLzMaxLen ? kLzMaxLen : (int)(n - pos);   // lzmh.c:180-190
      // ---- match search (lzmh.c:192-214): 4 offsets per lane (lane + 1 + 32 j), ascending, strictly longer wins
      uint32_t key = 0;   // (length << 8) | (255 - offset): max = longest, then smallest offset
      const uint8_t c0 = win[p];
      const uint32_t sp = (p & 3) * 8, wp = win_s + (p & ~3u);
      // the twelve bytes at the parse position (the same for every lane: broadcasts)
      const uint32_t t0 = lds_u32(wp), t1 = lds_u32(wp + 4), t2 = lds_u32(wp + 8), t3 = lds_u32(wp + 12);
      const uint32_t w0 = __funnelshift_r(t0, t1, sp), w1 = __funnelshift_r(t1, t2, sp), w2 = __funnelshift_r(t2, t3, sp);
      // A candidate that matched its first four bytes is extended: the next eight bytes against registers, anything
      // longer four bytes per trip; whatever is read past maxlen is cut off.  `wq` / `sq`: word address and alignment
      // shift of the candidate's first byte.
      auto extend = [&](const uint32_t wq, const uint32_t sq, const uint32_t off) {
        int len;
        uint32_t x = __funnelshift_r(lds_u32(wq + 4), lds_u32(wq + 8), sq) ^ w1;
        if (x != 0)
          len = 4 + ((__ffs((int)x) - 1) >> 3);
        else
        {
          x = __funnelshift_r(lds_u32(wq + 8), lds_u32(wq + 12), sq) ^ w2;
          if (x != 0)
            len = 8 + ((__ffs((int)x) - 1) >> 3);
          else
          {
            len = 12;
            while (len < maxlen)
            {
              x = __funnelshift_r(lds_u32(wq + len), lds_u32(wq + len + 4), sq) ^ __funnelshift_r(lds_u32(wp + len), lds_u32(wp + len + 4), sp);
              if (x != 0)
              {
                len += (__ffs((int)x) - 1) >> 3;
                break;
              }
              len += 4;
            }
          }
        }
        len = len < maxlen ? len : maxlen;
        if (len >= 3)
          key = max(key, ((uint32_t)len << 8) | (255u - off));
      };
      // Only matches of three bytes or more count (lzmh.c:216), so the first test is on the first word of every
      // candidate, for all lanes in step; the few that match all four bytes are extended afterwards, one per lane at a time.
      if (maxoff == kLzWindow)
      {
        // 128 bytes of history (all but the first steps of a stream).  Lane l takes the four candidates that start in
        // the word at p - 128 + 4 l, i.e. the offsets 128 - 4 l - k: three aligned words, shifted like the bytes at p,
        // hold all four first words.
        const uint32_t qa = wp - kLzWindow + 4 * lane;
        const uint32_t a0 = lds_u32(qa), a1 = lds_u32(qa + 4), a2 = lds_u32(qa + 8);
        const uint32_t d0 = __funnelshift_r(a0, a1, sp), d1 = __funnelshift_r(a1, a2, sp);
        const uint32_t x0 = d0 ^ w0, x1 = __funnelshift_r(d0, d1, 8) ^ w0, x2 = __funnelshift_r(d0, d1, 16) ^ w0,
                       x3 = __funnelshift_r(d0, d1, 24) ^ w0;
        uint32_t pending = 0;
        if (maxlen >= 3)
        {
          const uint32_t kb = (3u << 8) | (127u + 4 * lane);   // 255 - offset of candidate 0; the later ones are nearer
          key = (x0 & 0x00FFFFFFu) == 0 ? kb : 0u;
          key = (x1 & 0x00FFFFFFu) == 0 ? kb + 1 : key;
          key = (x2 & 0x00FFFFFFu) == 0 ? kb + 2 : key;
          key = (x3 & 0x00FFFFFFu) == 0 ? kb + 3 : key;
          pending = (x0 == 0 ? 1u : 0u) | (x1 == 0 ? 2u : 0u) | (x2 == 0 ? 4u : 0u) | (x3 == 0 ? 8u : 0u);
        }
        while (pending)
        {
          const uint32_t k = (uint32_t)__ffs((int)pending) - 1;
          pending &= pending - 1;
          const uint32_t q = p - kLzWindow + 4 * lane + k;
          extend(win_s + (q & ~3u), (q & 3) * 8, (uint32_t)kLzWindow - 4 * lane - k);
        }
      }
      else
      {
        // the first 128 bytes of a stream: offsets lane + 1 + 32 j as far as the history goes
        const uint32_t cbase = p - lane - 1;
        const uint32_t sc = (cbase & 3) * 8, wc = win_s + (cbase & ~3u);
        const int nval = maxoff > (int)lane ? ((maxoff - (int)lane - 1) >> 5) + 1 : 0;
        for (int j = 0; j < nval; j++)
        {
          const uint32_t x = __funnelshift_r(lds_u32(wc - 32 * j), lds_u32(wc - 32 * j + 4), sc) ^ w0;
          if (x == 0)
            extend(wc - 32 * j, sc, lane + 1 + 32 * j);
          else if ((x & 0x00FFFFFFu) == 0 && maxlen >= 3)
            key = max(key, (3u << 8) | (254u - lane - 32 * j));
        }
      }
      key = __reduce_max_sync(0xFFFFFFFFu, key);
      uint32_t code, nbits;
      if (key != 0)   // lzmh.c:216-284
      {
        const int best = (int)(key >> 8);
        const int off = 255 - (int)(key & 0xFF);
        const int k = off == recent0 ? 0 : off == recent1 ? 1 : off == recent2 ? 2 : off == recent3 ? 3 : 4;
        lz_token(k, off, best, code, nbits);
        if (k >= 3)   // move to front (slot 3 and "not cached" both shift everything)
          recent3 = recent2;
        if (k >= 2)
          recent2 = recent1;
        if (k >= 1)
          recent1 = recent0;
        recent0 = off;
        pos += (uint32_t)best;
      }
      else   // ATH literal (lzmh.c:285-335)
      {
        pos++;
        // rank of the byte in the count-sorted list (symbols are unique: at most one hit)
        const bool hit0 = lane < nlist && lsym[lane] == c0;
        const bool hit1 = lane + 32 < nlist && lsym[lane + 32] == c0;
        const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, hit0);
        const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, hit1);
        int rank = b0 ? __ffs(b0) - 1 : b1 ? __ffs(b1) + 31 : -1;
        if (rank >= 0)
        {
          const uint32_t c = lcnt[rank];
          if (c < 65535u)   // lzmh.c:301-312: count + 1, bubble up past strictly smaller counts
          {
            const uint32_t c1 = c + 1;
            const bool lt0 = (int)lane < rank && lcnt[lane] < c1;
            const bool lt1 = (int)lane + 32 < rank && lcnt[lane + 32] < c1;
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, lt0);
            const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, lt1);
            const int dstpos = m0 ? __ffs(m0) - 1 : m1 ? __ffs(m1) + 31 : rank;
            // entries (dstpos, rank] take their left neighbour's symbol (their counts all equal c)
            const int i0 = (int)lane, i1 = (int)lane + 32;
            const uint8_t s0 = (i0 > dstpos && i0 <= rank) ? lsym[i0 - 1] : 0;
            const uint8_t s1 = (i1 > dstpos && i1 <= rank) ? lsym[i1 - 1] : 0;
            __syncwarp();
            if (i0 > dstpos && i0 <= rank)
              lsym[i0] = s0;
            if (i1 > dstpos && i1 <= rank)
              lsym[i1] = s1;
            if (lane == 0)
            {
              lsym[dstpos] = c0;
              lcnt[dstpos] = c1;
            }
            __syncwarp();
          }
        }
        else if (nlist < (uint32_t)kLzList)   // lzmh.c:318-322
        {
          if (lane == 0)
          {
            lsym[nlist] = c0;
            lcnt[nlist] = 1;
          }
          nlist++;
          __syncwarp();
        }
        if (rank >= 0 && rank < kLzTree)   // lzmh.c:324-333
        {
          code = lz_tree_code(rank);
          nbits = lz_tree_len(rank);
        }
        else
        {
          code = c0;   // 00 + 8 bit
          nbits = 10;
        }
      }
      if (lane == 0)
        w.put(code, nbits);   // lzmh.c:337-339
    }
  }
  if (lane == 0)
  {
    const uint64_t bits = w.finish();
    a.status[s] = w.overflow ? -6 : 0;
    a.bits_out.nbits[s] = w.overflow ? 0 : bits;
  }
}
// ---------------------------------------------------------------------------------------------------------
// Decoder, one warp per stream (lzmh.c:383-574).
//
// The parse is bit-serial and every lane follows it in step (the bit window, the recent offsets and the output
// position are warp-uniform registers); the lanes share the work inside a token: the 19 prefix codes are matched by
// 19 lanes and a ballot, the rank list (shared memory) is searched and shifted like in the encoder, a match copy
// moves 32 bytes per trip.  The last 1024 output bytes live in a shared-memory ring -- the reference's 128-byte
// history is its tail -- and leave for global memory 512 bytes at a time, 16 bytes per lane.  The compressed input is
// read 128 bytes at a time, one byte-swapped word per lane, and handed round by a shuffle every 32 bits.
//
// The reference keeps a window of `have` <= 25 bits that it tops up bit by bit before every token (lzmh.c:409-413) and
// stops when the input is used up and the window is empty (lzmh.c:571).  Here the window is the top of a 64-bit
// accumulator of the stream with zeros behind its last bit, and `rem` counts the bits no token has used yet:
// have = min(25, rem), "input used up" = rem <= 25 when the token started, "window empty" = accumulator zero.
// 83 warp instructions per token; the first version (separate window, 64-bit positions, a division per overlapping
// copy) took 150.  Tried and measured no better: the rank list in registers (two packed entries per lane, shuffles instead of
// shared memory: 12.9 ms against 12.65), byte-wise loads at the stream edges (one CTA per SM less: 13.0 ms).
static const uint32_t kLzOutRing = 1024;    // bytes of output kept in shared memory per warp (power of two)
static const uint32_t kLzOutFlush = 512;    // bytes per flush
static const uint32_t kLzModTab = 32 * 32 + 32;   // [back][lane] = lane % back for back < 32, then [back] = 32 % back

struct LzWarpBits   // warp-uniform MSB-first bit source
{
  const uint32_t *pchunk;   // word 0 of the 32-word chunk the lanes hold in w0 (w1: the chunk after it)
  const uint32_t *plast;    // the word with the stream's last bit
  uint32_t last_mask;       // its stream bits
  uint32_t w0, w1, lane;
  uint32_t wi;              // next word of the chunk to hand out
  uint32_t a_hi, a_lo;      // the next 64 bits of the stream, left-aligned, zero beyond its end
  uint32_t valid;           // how many of them are loaded: 33..64 between tokens

  // Whole aligned words, like BitReader: a word that holds a byte of the stream lies inside the allocation the stream
  // lies in (device allocations are multiples of 256 bytes), so nothing outside the caller's buffer is touched.
  __device__ __forceinline__ uint32_t load(const uint32_t *p) const
  {
    if (p > plast)
      return 0u;
    const uint32_t w = __byte_perm(__ldg(p), 0, 0x0123);
    return p == plast ? w & last_mask : w;
  }
  __device__ __forceinline__ uint32_t next_word()
  {
    const uint32_t w = __shfl_sync(0xFFFFFFFFu, w0, wi);
    if (++wi == 32)
    {
      wi = 0;
      pchunk += 32;
      w0 = w1;
      w1 = load(pchunk + 32 + lane);
    }
    return w;
  }
  __device__ __forceinline__ void consume(uint32_t used)   // used < 32
  {
    a_hi = __funnelshift_l(a_lo, a_hi, used);
    a_lo <<= used;
    valid -= used;
    if (valid <= 32)   // a_lo is empty: the next word goes `valid` bits below the top
    {
      const uint32_t w = next_word();
      a_hi |= __funnelshift_rc(w, 0u, valid);
      a_lo = __funnelshift_rc(0u, w, valid);
      valid += 32;
    }
  }
  __device__ __forceinline__ void init(const uint8_t *p, uint64_t nbits, uint32_t lane_)
  {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3);
    const uint32_t *base = reinterpret_cast<const uint32_t *>(p - mis);
    const uint64_t end = mis * 8 + nbits;
    plast = nbits ? base + ((end - 1) >> 5) : base - 1;
    last_mask = 0xFFFFFFFFu << (31 - ((uint32_t)(end - 1) & 31));
    lane = lane_;
    pchunk = base;
    w0 = load(base + lane);
    w1 = load(base + 32 + lane);
    wi = 0;
    a_hi = next_word();
    a_lo = next_word();
    valid = 64;
    consume(mis * 8);
  }
};

__device__ inline void lzmh_decode_warp(const LzmhArgs &a, const uint64_t s, const uint32_t ring, const uint32_t lsym, const uint32_t lcnt,
                                        const uint32_t modtab)
{
  const uint32_t lane = threadIdx.x & 31;
  if (a.status[s] != 0)   // marked by the index check (its byte range leaves the payload): nothing is read
  {
    if (lane == 0)
      a.out_len[s] = 0;
    return;
  }
  LzWarpBits r;
  const uint64_t nbits = a.bits_in.nbits[s];
  r.init(src_ptr(a.bits_in, s), nbits, lane);
  long long rem = (long long)nbits;   // bits no token has used yet
  uint8_t *out = a.out + a.out_off[s];
  const uint64_t cap = a.out_off[s + 1] - a.out_off[s];
  const bool out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  uint64_t flushed = 0;   // bytes already in global memory (a multiple of kLzOutFlush until the end)
  uint32_t pend = 0;      // bytes produced since; the produced total is flushed + pend
  uint32_t rpos = 0;      // the produced total mod 2^32: its low bits are the ring slot of the next byte

  for (uint32_t i = lane; i < kLzOutRing / 4; i += 32)   // never-written history reads as zero (lzmh.c:396)
    sts_u32(ring + 4 * i, 0);
  for (uint32_t i = lane; i < (uint32_t)kLzList; i += 32)
  {
    sts_u8(lsym + i, 0);
    sts_u32(lcnt + 4 * i, 0);
  }
  __syncwarp();

  // lane i < 19 matches prefix code i (lzmh.c:86-106)
  const uint32_t my_len = lz_tree_len((int)lane);
  const uint32_t my_code = lane < (uint32_t)kLzTree ? lz_tree_code((int)lane) : 0xFFFFFFFFu;
  int recent0 = 0, recent1 = 0, recent2 = 0, recent3 = 0;

  // the oldest `n` pending bytes go to global memory; a whole aligned 512-byte piece as one 16-byte store per lane
  auto flush = [&](uint32_t n) {
    __syncwarp();
    const uint32_t slot = (rpos - pend) & (kLzOutRing - 1);
    if (n == kLzOutFlush && out_aligned && flushed + kLzOutFlush <= cap)
    {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring + slot + 16 * lane) : "memory");
      *reinterpret_cast<uint4 *>(out + flushed + 16 * lane) = v;
    }
    else
      for (uint32_t i = lane; i < n; i += 32)
        if (flushed + i < cap)
          out[flushed + i] = (uint8_t)lds_u8(ring + ((slot + i) & (kLzOutRing - 1)));
    flushed += n;
    pend -= n;
    __syncwarp();   // the ring slots just read may be written again by the next token (pend can exceed 750 of the 1024)
  };

  for (;;)
  {
    if (pend >= kLzOutFlush)
      flush(kLzOutFlush);
    const uint32_t window = r.a_hi;
    const bool input_used_up = rem <= 25;   // every remaining bit is in the reference's window: its reader is at EOF
    uint32_t used;
    if (window & 0x80000000u)   // static prefix code -> rank (lzmh.c:414-450): the first code that matches
    {
      const bool hit = (window >> (32 - my_len)) == my_code && (!input_used_up || rem >= (long long)my_len);
      const uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
      if (hits == 0)   // lzmh.c:447-449: silent stop
        break;
      const int i = __ffs((int)hits) - 1;
      used = __shfl_sync(0xFFFFFFFFu, my_len, i);
      const uint32_t byte = lds_u8(lsym + i);
      const uint32_t c = lds_u32(lcnt + 4 * i);
      if (lane == 0)
        sts_u8(ring + (rpos & (kLzOutRing - 1)), byte);
      rpos++;
      pend++;
      if (c < 65535u)   // lzmh.c:431-442: count + 1, move left past strictly smaller counts (they all equal c)
      {
        const uint32_t c1 = c + 1;
        const uint32_t lt = __ballot_sync(0xFFFFFFFFu, (int)lane < i && lds_u32(lcnt + 4 * lane) < c1);
        const int dstpos = lt ? __ffs((int)lt) - 1 : i;
        const bool moves = (int)lane > dstpos && (int)lane <= i;
        const uint32_t left = moves ? lds_u8(lsym + lane - 1) : 0;
        __syncwarp();
        if (moves)
          sts_u8(lsym + lane, left);
        if (lane == 0)
        {
          sts_u8(lsym + dstpos, byte);
          sts_u32(lcnt + 4 * dstpos, c1);
        }
        __syncwarp();
      }
    }
    else if ((window & 0x40000000u) == 0)   // 00 + byte (lzmh.c:451-479)
    {
      const uint32_t byte = (window >> 22) & 0xFF;
      used = 10;
      if (lane == 0)
        sts_u8(ring + (rpos & (kLzOutRing - 1)), byte);
      rpos++;
      pend++;
      // first entry that is free or holds the byte
      const uint32_t i0 = lane, i1 = lane + 32;
      const bool in1 = i1 < (uint32_t)kLzList;
      const uint32_t cnt0 = lds_u32(lcnt + 4 * i0), cnt1 = in1 ? lds_u32(lcnt + 4 * i1) : 0;
      const uint32_t sym0 = lds_u8(lsym + i0), sym1 = in1 ? lds_u8(lsym + i1) : 0;
      const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, cnt0 == 0 || sym0 == byte);
      const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, in1 && (cnt1 == 0 || sym1 == byte));
      const int pos = b0 ? __ffs((int)b0) - 1 : b1 ? __ffs((int)b1) + 31 : kLzList;
      if (pos < kLzList)
      {
        const uint32_t c = __shfl_sync(0xFFFFFFFFu, pos < 32 ? cnt0 : cnt1, pos & 31);
        if (c < 65535u)
        {
          const uint32_t c1 = c + 1;
          const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, (int)i0 < pos && cnt0 < c1);
          const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, (int)i1 < pos && cnt1 < c1);
          const int dstpos = m0 ? __ffs((int)m0) - 1 : m1 ? __ffs((int)m1) + 31 : pos;
          // entries [dstpos, pos) move one place right (their counts all equal c, so only the symbols move)
          if ((int)i0 >= dstpos && (int)i0 < pos)
            sts_u8(lsym + i0 + 1, sym0);
          if ((int)i1 >= dstpos && (int)i1 < pos)
            sts_u8(lsym + i1 + 1, sym1);
          __syncwarp();
          if (lane == 0)
          {
            sts_u8(lsym + dstpos, byte);
            sts_u32(lcnt + 4 * dstpos, c1);
          }
          __syncwarp();
        }
      }
    }
    else   // 01: LZ token (lzmh.c:480-569)
    {
      uint32_t w = window << 2;
      used = 2;
      int off, k;
      if ((w & 0x80000000u) == 0)   // 0 + explicit 7-bit offset
      {
        off = (int)((w >> 24) & 0x7F) + 1;
        w <<= 8;
        used += 8;
        k = 3;
      }
      else   // recent offsets: 10 -> slot 0, 110 -> 1, 1110 -> 2, 1111 -> 3
      {
        const uint32_t ones = (uint32_t)__clz((int)~w);
        k = ones > 3 ? 3 : (int)ones - 1;
        const uint32_t n = ones > 3 ? 4u : ones + 1;
        w <<= n;
        used += n;
        off = k == 0 ? recent0 : k == 1 ? recent1 : k == 2 ? recent2 : recent3;
      }
      if (k >= 3)
        recent3 = recent2;
      if (k >= 2)
        recent2 = recent1;
      if (k >= 1)
        recent1 = recent0;
      recent0 = off;
      uint32_t len;
      if ((w & 0x80000000u) == 0)
      {
        len = ((w >> 28) & 0x7) + 3;
        used += 4;
      }
      else if ((w & 0x40000000u) == 0)
      {
        len = ((w >> 27) & 0x7) + 11;
        used += 5;
      }
      else
      {
        len = ((w >> 22) & 0xFF) + 19;
        used += 10;
      }
      // lzmh.c:560-568: `len` bytes from `back` bytes ago (128 for a zero offset out of the empty recent list); slots that
      // were never written hold zero
      const uint32_t back = off == 0 ? (uint32_t)kLzWindow : (uint32_t)off;
      __syncwarp();
      if (back >= 32)
      {
        uint32_t dst = rpos + lane;
        if (lane < len)
          sts_u8(ring + (dst & (kLzOutRing - 1)), lds_u8(ring + ((dst - back) & (kLzOutRing - 1))));
#pragma unroll 1
        for (uint32_t j0 = 32; j0 < len; j0 += 32)   // a trip reads bytes written before it began
        {
          __syncwarp();
          dst += 32;
          if (j0 + lane < len)
            sts_u8(ring + (dst & (kLzOutRing - 1)), lds_u8(ring + ((dst - back) & (kLzOutRing - 1))));
        }
      }
      else
      {
        // the copy may overlap itself: output byte j is byte j mod back of the `back` bytes before it, which no trip
        // overwrites
        uint32_t m = lds_u8(modtab + back * 32 + lane);
        const uint32_t src = rpos - back;
        uint32_t dst = rpos + lane;
        if (lane < len)
          sts_u8(ring + (dst & (kLzOutRing - 1)), lds_u8(ring + ((src + m) & (kLzOutRing - 1))));
        if (len > 32)
        {
          const uint32_t step = lds_u8(modtab + 32 * 32 + back);
#pragma unroll 1
          for (uint32_t j0 = 32; j0 < len; j0 += 32)
          {
            m += step;
            m = m >= back ? m - back : m;
            dst += 32;
            if (j0 + lane < len)
              sts_u8(ring + (dst & (kLzOutRing - 1)), lds_u8(ring + ((src + m) & (kLzOutRing - 1))));
          }
        }
      }
      __syncwarp();
      rpos += len;
      pend += len;
    }
    rem -= (long long)used;
    r.consume(used);
    if (input_used_up && r.a_hi == 0)   // lzmh.c:571: reader at EOF and window empty
      break;
  }
  while (pend)
    flush(pend < kLzOutFlush ? pend : kLzOutFlush);
  if (lane == 0)
  {
    const int err = flushed > cap ? -6 : 0;
    a.status[s] = err;
    a.out_len[s] = err ? 0 : flushed;
  }
}
#endif   // !DC_HOST_SIM

}  // namespace dc
