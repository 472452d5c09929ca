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
//   decoder: one THREAD per stream (bit-serial parse; copies read the thread's own output).
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
static const int kLzWin = 1024;   // bytes of the stream kept in shared memory per warp
static const int kLzWinPad = 16;  // readable slack behind the window (word-wise match compare)

// four bytes starting at byte `idx` of a 4-byte aligned shared-memory array, first byte in the low bits
__device__ __forceinline__ uint32_t lz_load32(const uint8_t *base, uint32_t idx)
{
  const uint32_t *w = reinterpret_cast<const uint32_t *>(base) + (idx >> 2);
  return __funnelshift_r(w[0], w[1], (idx & 3) * 8);
}

__device__ inline void lzmh_encode_warp(const LzmhArgs &a, const uint64_t s, uint8_t *win, uint8_t *lsym, uint32_t *lcnt)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint8_t *src = a.in + a.in_off[s];
  const uint32_t n = (uint32_t)(a.in_off[s + 1] - a.in_off[s]);
  BitWriter w;
  w.init(dst_ptr(a.bits_out, s), a.bits_out.stride_bytes);

  if (n != (uint32_t)kLzRing)   // lzmh.c:169-173: the initial fill of exactly 403 bytes wraps write == read at EOF
  {
    int recent0 = 0, recent1 = 0, recent2 = 0, recent3 = 0;   // lzmh.c:157-159
    uint32_t nlist = 0;
    uint32_t wbase = 0;
    for (uint32_t i = lane; i < kLzWin && i < n; i += 32)
      win[i] = src[i];
    __syncwarp();
    uint32_t pos = 0;
    while (pos < n)   // lzmh.c:173
    {
      if (pos + kLzMaxLen > wbase + kLzWin && wbase + kLzWin < n)   // slide the window: keep 128 bytes of history
      {
        __syncwarp();
        wbase = pos - kLzWindow;
        for (uint32_t i = lane; i < kLzWin && wbase + i < n; i += 32)
          win[i] = src[wbase + i];
        __syncwarp();
      }
      const uint32_t p = pos - wbase;
      const int maxoff = pos < (uint32_t)kLzWindow ? (int)pos : kLzWindow;            // lzmh.c:175-179
      const int maxlen = n - pos > (uint32_t)kLzMaxLen ? kLzMaxLen : (int)(n - pos);   // lzmh.c:180-190
      // ---- match search (lzmh.c:192-214): 4 offsets per lane, ascending, strictly longer wins
      uint32_t key = 0;   // (length << 8) | (255 - offset): max = longest, then smallest offset
      const uint8_t c0 = win[p];
      const uint32_t w0 = lz_load32(win, p);   // the four bytes at the parse position (same for every lane: a broadcast)
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        const int off = (int)lane + 1 + 32 * j;
        if (off <= maxoff)
        {
          // Only matches of three bytes or more count (lzmh.c:216), so the first test is on three bytes at once and
          // runs for all lanes in step; the extension loop -- four bytes per trip, unaligned reads assembled from
          // aligned words, whatever is read past maxlen cut off below -- is left to the few candidates that matched
          // four bytes already.
          uint32_t x = lz_load32(win, p - off) ^ w0;
          if ((x & 0x00FFFFFFu) == 0)
          {
            int len = 3;
            if (x == 0)
            {
              len = 4;
              while (len < maxlen)
              {
                x = lz_load32(win, p - off + len) ^ lz_load32(win, p + len);
                if (x != 0)
                {
                  len += (__ffs((int)x) - 1) >> 3;
                  break;
                }
                len += 4;
              }
            }
            len = len < maxlen ? len : maxlen;
            const uint32_t k2 = ((uint32_t)len << 8) | (uint32_t)(255 - off);
            if (len >= 3 && k2 > key)
              key = k2;
          }
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
#endif   // !DC_HOST_SIM

// ---------------------------------------------------------------------------------------------------------
// Decoder, one thread per stream (lzmh.c:383-574)
__device__ inline void lzmh_decode_body(const LzmhArgs &a, const uint64_t s)
{
  BitReader r;
  r.init(src_ptr(a.bits_in, s), a.bits_in.nbits[s]);
  uint8_t *out = a.out + a.out_off[s];
  const uint64_t cap = a.out_off[s + 1] - a.out_off[s];
  uint64_t np = 0;   // bytes produced

  uint32_t window = 0;   // code_sym, lzmh.c:393
  int have = 0;          // code_length, lzmh.c:394
  int recent[4] = { 0, 0, 0, 0 };
  uint8_t lsym[kLzList];
  uint32_t lcnt[kLzList];
  for (int i = 0; i < kLzList; i++)
  {
    lsym[i] = 0;
    lcnt[i] = 0;
  }
  int err = 0;

  // a byte leaves the decoder: output + history (the reference's 128-byte ring is the tail of the output)
#define LZ_EMIT(byte_)              \
  do                                \
  {                                 \
    if (np < cap)                   \
      out[np] = (uint8_t)(byte_);   \
    np++;                           \
  } while (0)

  do
  {
    // lzmh.c:409-413: bit-by-bit refill while at least 8 bits are free, i.e. up to 25 buffered bits
    if (have < 25)
    {
      const uint64_t av = r.avail();
      const uint32_t want = (uint32_t)(25 - have);
      const uint32_t t = av < want ? (uint32_t)av : want;
      if (t)
      {
        window |= r.take(t) << (32 - have - t);
        have += (int)t;
      }
    }
    if (window & 0x80000000u)   // static prefix code -> rank (lzmh.c:414-450)
    {
      int i;
      for (i = 0; i < kLzTree; i++)
      {
        const uint32_t L = lz_tree_len(i);
        if (have >= (int)L && (window >> (32 - L)) == lz_tree_code(i))
          break;
      }
      if (i == kLzTree)   // lzmh.c:447-449: silent stop
        break;
      const uint32_t L = lz_tree_len(i);
      have -= (int)L;
      window <<= L;
      const uint8_t byte = lsym[i];
      LZ_EMIT(byte);
      if (lcnt[i] < 65535u)   // lzmh.c:431-442
      {
        const uint32_t c = lcnt[i] + 1;
        while (i > 0 && c > lcnt[i - 1])
        {
          lsym[i] = lsym[i - 1];
          i--;
        }
        lcnt[i] = c & 0xFFFFu;
        lsym[i] = byte;
      }
    }
    else if ((window & 0x40000000u) == 0)   // 00 + byte (lzmh.c:451-479)
    {
      const uint8_t byte = (uint8_t)((window >> 22) & 0xFF);
      have -= 10;
      window <<= 10;
      LZ_EMIT(byte);
      int pos = 0;
      while (pos < kLzList && lcnt[pos] > 0 && lsym[pos] != byte)
        pos++;
      if (pos < kLzList && lcnt[pos] < 65535u)
      {
        const uint32_t c = lcnt[pos] + 1;
        while (pos > 0 && c > lcnt[pos - 1])
        {
          lsym[pos] = lsym[pos - 1];
          lcnt[pos] = lcnt[pos - 1];
          pos--;
        }
        lcnt[pos] = c;
        lsym[pos] = byte;
      }
    }
    else   // 01: LZ token (lzmh.c:480-569)
    {
      int off, len, k;
      have -= 2;
      window <<= 2;
      if ((window & 0x80000000u) == 0)   // explicit 7-bit offset
      {
        off = (int)((window >> 24) & 0x7F) + 1;
        have -= 8;
        window <<= 8;
        k = 3;
      }
      else
      {
        have -= 1;
        window <<= 1;
        if ((window & 0x80000000u) == 0)
          k = 0;
        else
        {
          have -= 1;
          window <<= 1;
          if ((window & 0x80000000u) == 0)
            k = 1;
          else
          {
            have -= 1;
            window <<= 1;
            k = (window & 0x80000000u) == 0 ? 2 : 3;
          }
        }
        have -= 1;
        window <<= 1;
        off = recent[k];
      }
      for (; k > 0; k--)
        recent[k] = recent[k - 1];
      recent[0] = off;
      if ((window & 0x80000000u) == 0)
      {
        len = (int)((window >> 28) & 0x7) + 3;
        have -= 4;
        window <<= 4;
      }
      else
      {
        have -= 1;
        window <<= 1;
        if ((window & 0x80000000u) == 0)
        {
          len = (int)((window >> 28) & 0x7) + 11;
          have -= 4;
          window <<= 4;
        }
        else
        {
          len = (int)((window >> 23) & 0xFF) + 19;
          have -= 9;
          window <<= 9;
        }
      }
      for (int j = 0; j < len; j++)   // lzmh.c:560-568
      {
        // ring slot (hpos - off) mod 128: a byte written `off` back, the byte 128 back for off == 0, and a
        // never-written (zero) slot when the stream is shorter than that
        const uint64_t back = off == 0 ? (uint64_t)kLzWindow : (uint64_t)off;
        uint8_t byte = 0;
        if (back <= np && np - back < cap)
          byte = out[np - back];
        LZ_EMIT(byte);
      }
    }
  } while (!r.eof() || window > 0);   // lzmh.c:571
#undef LZ_EMIT
  if (np > cap)
    err = -6;
  a.status[s] = err;
  a.out_len[s] = err ? 0 : np;
}

}  // namespace dc
