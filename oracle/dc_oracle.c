/* dc_oracle.c -- CPU restatement ("oracle") of the Data Compressor hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see dc_oracle.h).  Plain C, single thread, no dependency on the
 * reference tree.  Compile WITHOUT -ffast-math and with -ffp-contract=off: the reference release
 * build evaluates `v*nf`, `+-0.5f` and `(float)n / nf` as separately rounded IEEE float32 operations
 * ("Bit-exact float->int").
 *
 * Citations: paths are relative to /root/reference/DataCompressor/.
 */
#include "dc_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* Bit strings: MSB-first packing, DCIOLib/src/bit_file_buffer.c:220-248 (write), 148-187 (read) */

void dco_bits_init(dco_bits *b)
{
  b->data = NULL;
  b->cap = 0;
  b->nbits = 0;
}

void dco_bits_free(dco_bits *b)
{
  free(b->data);
  dco_bits_init(b);
}

static int bits_reserve(dco_bits *b, uint64_t extra_bits)
{
  const uint64_t need = (b->nbits + extra_bits + 7) / 8 + 8;
  if (need > b->cap)
  {
    size_t ncap = b->cap ? b->cap : 64;
    uint8_t *p;
    while (ncap < need)
      ncap *= 2;
    p = (uint8_t *)realloc(b->data, ncap);
    if (p == NULL)
      return DCO_ERR_MEMORY;
    memset(p + b->cap, 0, ncap - b->cap);
    b->data = p;
    b->cap = ncap;
  }
  return DCO_OK;
}

/* Append the low n bits of v, most significant first (WriteSingleValueToBitFileBuffer,
   bit_file_buffer.c:297-308).  n == 0 is a legal no-op (SEG uses it for p == 0). */
static int put_bits(dco_bits *b, uint64_t v, unsigned n)
{
  int ret;
  if (n == 0)
    return DCO_OK;
  if ((ret = bits_reserve(b, n)) != DCO_OK)
    return ret;
  while (n > 0)
  {
    const unsigned used = (unsigned)(b->nbits & 7);
    const unsigned room = 8 - used;
    const unsigned take = n < room ? n : room;
    const uint64_t chunk = (v >> (n - take)) & ((1u << take) - 1u);
    b->data[b->nbits >> 3] |= (uint8_t)(chunk << (room - take));
    b->nbits += take;
    n -= take;
  }
  return DCO_OK;
}

int dco_bits_from_bytes(dco_bits *b, const uint8_t *bytes, uint64_t nbits)
{
  int ret;
  const uint64_t nbytes = (nbits + 7) / 8;
  dco_bits_free(b);
  if ((ret = bits_reserve(b, nbits)) != DCO_OK)
    return ret;
  if (nbytes)
    memcpy(b->data, bytes, nbytes);
  if (nbits & 7) /* keep the padding bits zero so that later appends can OR into them */
    b->data[nbytes - 1] &= (uint8_t)(0xFF00u >> (nbits & 7));
  b->nbits = nbits;
  return DCO_OK;
}

typedef struct bit_reader
{
  const uint8_t *data;
  uint64_t nbits;
  uint64_t pos;
} bit_reader;

static void rd_init(bit_reader *r, const dco_bits *b)
{
  r->data = b->data;
  r->nbits = b->nbits;
  r->pos = 0;
}

/* EndOfBitFileBuffer, bit_file_buffer.c:53-62 */
static int rd_eof(const bit_reader *r)
{
  return r->pos >= r->nbits;
}

/* Reads n <= 64 bits, zero-extended (ReadSingleValueFromBitFileBuffer, bit_file_buffer.c:280-294).
   A short read is what READ_VALUE_BITS_CHECKED turns into ERROR_LIBRARY_CALL (io_macros.h:13-27,60-65). */
static int get_bits(bit_reader *r, unsigned n, uint64_t *v)
{
  uint64_t acc = 0;
  unsigned i;
  if (r->nbits - r->pos < n)
  {
    r->pos = r->nbits;
    return DCO_ERR_LIBRARY_CALL;
  }
  for (i = 0; i < n; i++)
  {
    const uint64_t p = r->pos++;
    acc = (acc << 1) | ((r->data[p >> 3] >> (7 - (p & 7))) & 1u);
  }
  *v = acc;
  return DCO_OK;
}

/* EXTEND_IO_INT_SIGN, DCLib/inc/io_macros.h:89-90 */
static int64_t sign_extend(uint64_t v, unsigned vs)
{
  if (vs == 64)
    return (int64_t)v;
  return (int64_t)(v << (64 - vs)) >> (64 - vs);
}

/* ------------------------------------------------------------------------------------------- */
/* normalize: DCLib/src/normalize.c */

static int64_t float_to_i64_x86(float v)
{
  /* (io_int_t)value compiles to cvttss2si: out-of-range and NaN give INT64_MIN. */
  if (!(v >= -9223372036854775808.0f && v < 9223372036854775808.0f))
    return INT64_MIN;
  return (int64_t)v;
}

int dco_normalize(const dco_bits *in, dco_bits *out, unsigned vs, float nf)
{
  bit_reader r;
  const float lo_bound = -(float)((uint64_t)1 << (vs - 1));         /* normalize.c:21, left operand  */
  const float hi_bound = (float)(((uint64_t)1 << (vs - 1)) - 1);    /* normalize.c:21, uint64 -> float */
  rd_init(&r, in);
  while (!rd_eof(&r))                                               /* normalize.c:11 */
  {
    uint8_t raw[4];
    uint64_t byte;
    float value;
    int64_t n;
    int i, ret;
    for (i = 0; i < 4; i++)                                         /* normalize.c:16, native LE float */
    {
      if ((ret = get_bits(&r, 8, &byte)) != DCO_OK)
        return ret;
      raw[i] = (uint8_t)byte;
    }
    memcpy(&value, raw, 4);
    if (value > 0)                                                  /* normalize.c:17-20 */
    {
      volatile float t = value * nf;                                /* two separately rounded ops */
      value = t + 0.5f;
    }
    else if (value < 0)
    {
      volatile float t = value * nf;
      value = t - 0.5f;
    }
    if (value < lo_bound || value > hi_bound)                       /* normalize.c:21-22 */
      return DCO_ERR_INVALID_VALUE;
    n = float_to_i64_x86(value);                                    /* normalize.c:23 */
    if ((ret = put_bits(out, (uint64_t)n, vs)) != DCO_OK)           /* normalize.c:24 */
      return ret;
  }
  return DCO_OK;
}

int dco_denormalize(const dco_bits *in, dco_bits *out, unsigned vs, float nf)
{
  bit_reader r;
  rd_init(&r, in);
  while (!rd_eof(&r))                                               /* normalize.c:31 */
  {
    uint64_t u;
    int64_t n;
    volatile float f;
    uint8_t raw[4];
    int i, ret;
    if ((ret = get_bits(&r, vs, &u)) != DCO_OK)                     /* normalize.c:36 */
      return ret;
    n = sign_extend(u, vs);                                         /* normalize.c:37 */
    f = (float)n;                                                   /* cvtsi2ss */
    f = f / nf;                                                     /* divss, normalize.c:38 */
    {
      float tmp = f;
      memcpy(raw, &tmp, 4);
    }
    for (i = 0; i < 4; i++)                                         /* normalize.c:39 */
      if ((ret = put_bits(out, raw[i], 8)) != DCO_OK)
        return ret;
  }
  return DCO_OK;
}

/* ------------------------------------------------------------------------------------------- */
/* diff: DCLib/src/diff.c */

int dco_diff_encode(const dco_bits *in, dco_bits *out, unsigned vs)
{
  bit_reader r;
  int64_t last = 0;                                                 /* diff.c:11 */
  rd_init(&r, in);
  while (!rd_eof(&r))
  {
    uint64_t u;
    int64_t value, d;
    int ret;
    if ((ret = get_bits(&r, vs, &u)) != DCO_OK)                     /* diff.c:15: NOT sign-extended */
      return ret;
    value = (int64_t)u;
    d = (int64_t)((uint64_t)value - (uint64_t)last);                /* diff.c:16 */
    if (vs != 64)                                                   /* diff.c:17-18 */
    {
      const int64_t lim = (int64_t)1 << (vs - 1);
      if (d < -lim || d > lim - 1)
        return DCO_ERR_INVALID_VALUE;
    }
    if ((ret = put_bits(out, (uint64_t)d, vs)) != DCO_OK)           /* diff.c:19 */
      return ret;
    last = value;                                                   /* diff.c:20 */
  }
  return DCO_OK;
}

int dco_diff_decode(const dco_bits *in, dco_bits *out, unsigned vs)
{
  bit_reader r;
  int64_t last = 0;                                                 /* diff.c:27 */
  rd_init(&r, in);
  while (!rd_eof(&r))
  {
    uint64_t u;
    int64_t value;
    int ret;
    if ((ret = get_bits(&r, vs, &u)) != DCO_OK)                     /* diff.c:31 */
      return ret;
    value = sign_extend(u, vs);                                     /* diff.c:32 */
    value = (int64_t)((uint64_t)value + (uint64_t)last);            /* diff.c:33, untruncated int64 */
    if ((ret = put_bits(out, (uint64_t)value, vs)) != DCO_OK)       /* diff.c:34 */
      return ret;
    last = value;                                                   /* diff.c:35 */
  }
  return DCO_OK;
}

/* ------------------------------------------------------------------------------------------- */
/* seg: DCLib/src/seg.c */

int dco_seg_encode(const dco_bits *in, dco_bits *out, unsigned vs)
{
  bit_reader r;
  rd_init(&r, in);
  while (!rd_eof(&r))                                               /* seg.c:34 */
  {
    uint64_t u, mapped, plus1, t;
    int64_t x;
    unsigned p = 0;
    int ret;
    if ((ret = get_bits(&r, vs, &u)) != DCO_OK)                     /* seg.c:37 */
      return ret;
    x = sign_extend(u, vs);                                         /* seg.c:38 */
    if (x > 0)                                                      /* seg.c:25-28 */
      mapped = 2 * (uint64_t)x - 1;
    else
      mapped = 2 * (uint64_t)(x < 0 ? -(uint64_t)x : 0);            /* llabs(INT64_MIN) wraps like the reference */
    plus1 = mapped + 1;                                             /* seg.c:13 */
    t = plus1;
    while ((t >>= 1) != 0)                                          /* seg.c:15-17 */
      p++;
    if ((ret = put_bits(out, 0, p)) != DCO_OK)                      /* seg.c:18: p zero bits */
      return ret;
    if ((ret = put_bits(out, plus1, p + 1)) != DCO_OK)              /* seg.c:19 */
      return ret;
  }
  return DCO_OK;
}

int dco_seg_decode(const dco_bits *in, dco_bits *out, unsigned vs)
{
  bit_reader r;
  const unsigned max_prefix = vs + 1 > 64 ? 64 : vs + 1;           /* seg.c:74 */
  rd_init(&r, in);
  while (!rd_eof(&r))                                               /* seg.c:84 */
  {
    unsigned p = 0;
    uint64_t bit = 0, rest = 0, m;
    int64_t x;
    int ret;
    while (bit == 0 && !rd_eof(&r))                                 /* seg.c:50-57 */
    {
      if ((ret = get_bits(&r, 1, &bit)) != DCO_OK)
        return ret;
      if (bit == 0)
        p++;
      if (p >= max_prefix)
        return DCO_ERR_INVALID_FORMAT;
    }
    if (rd_eof(&r) && p != 0)                                       /* seg.c:58-62: padding / end of stream */
      break;
    if ((ret = get_bits(&r, p, &rest)) != DCO_OK)                   /* seg.c:64 */
      return ret;
    m = (rest | ((uint64_t)1 << p)) - 1;                            /* seg.c:65-66 */
    x = (int64_t)((m + 1) / 2);                                     /* seg.c:76 */
    if ((m & 1) == 0)                                               /* seg.c:77-78 */
      x = -x;
    if ((ret = put_bits(out, (uint64_t)x, vs)) != DCO_OK)           /* seg.c:92 */
      return ret;
  }
  return DCO_OK;
}

/* ------------------------------------------------------------------------------------------- */
/* bac: DCLib/src/bac.c -- Witten/Neal/Cleary coder, RANGE_BITS = 16, symbols {0, 1, EOF} */

#define BAC_TOP      0xFFFFu          /* MAX_RANGE, bac.c:22 */
#define BAC_QUARTER  0x4000u          /* bac.c:23 */
#define BAC_HALF     0x8000u          /* bac.c:24 */
#define BAC_3QUARTER 0xC000u          /* bac.c:25 */
#define BAC_MAX_FREQ 0x3FFFu          /* bac.c:27 */
#define BAC_EOF_IDX  3                /* bac.c:30 */

typedef struct bac_model
{
  uint16_t freq[4];                   /* bac.c:36 */
  uint16_t cum[4];                    /* bac.c:37 */
  unsigned sym_to_idx[2];             /* bac.c:33 */
  unsigned idx_to_sym[4];             /* bac.c:34 */
} bac_model;

static void bac_model_init(bac_model *m)                            /* bac.c:39-52 */
{
  unsigned i;
  for (i = 0; i < 2; i++)
  {
    m->sym_to_idx[i] = i + 1;
    m->idx_to_sym[i + 1] = i;
  }
  m->idx_to_sym[0] = m->idx_to_sym[3] = 0;
  for (i = 0; i < 4; i++)
  {
    m->freq[i] = (uint16_t)(i == 0 ? 0 : 1);
    m->cum[i] = (uint16_t)(3 - i);
  }
}

static void bac_model_update(bac_model *m, unsigned idx)            /* bac.c:54-81 */
{
  unsigned i;
  if (m->cum[0] == BAC_MAX_FREQ)                                    /* bac.c:57-67 */
  {
    uint16_t run = 0;
    for (i = 4; i-- != 0;)
    {
      m->freq[i] = (uint16_t)((m->freq[i] + 1) / 2);
      m->cum[i] = run;
      run = (uint16_t)(run + m->freq[i]);
    }
  }
  for (i = idx; m->freq[i] == m->freq[i - 1]; i--)                  /* bac.c:68 */
    ;
  if (i < idx)                                                      /* bac.c:69-77 */
  {
    const unsigned cur = m->idx_to_sym[i];
    const unsigned last = m->idx_to_sym[idx];
    m->idx_to_sym[i] = last;
    m->idx_to_sym[idx] = cur;
    m->sym_to_idx[cur] = idx;
    m->sym_to_idx[last] = i;
  }
  m->freq[i]++;                                                     /* bac.c:78 */
  while (i-- > 0)                                                   /* bac.c:79-80 */
    m->cum[i]++;
}

typedef struct bac_enc
{
  uint16_t lo, hi;                                                  /* start_range / end_range, bac.c:83 */
  uint64_t pending;                                                 /* next_bits, bac.c:84 */
  dco_bits *out;
} bac_enc;

static int bac_emit(bac_enc *e, int bit)                            /* OutputNextBits, bac.c:93-105 */
{
  int ret;
  if ((ret = put_bits(e->out, (uint64_t)(bit ? 1 : 0), 1)) != DCO_OK)
    return ret;
  while (e->pending != 0)
  {
    if ((ret = put_bits(e->out, (uint64_t)(bit ? 0 : 1), 1)) != DCO_OK)
      return ret;
    e->pending--;
  }
  return DCO_OK;
}

static int bac_encode_symbol(bac_enc *e, const bac_model *m, unsigned idx)   /* bac.c:107-139 */
{
  const uint64_t range = (uint64_t)(uint16_t)(e->hi - e->lo) + 1;   /* bac.c:109 */
  int ret;
  e->hi = (uint16_t)(e->lo + (uint16_t)((range * m->cum[idx - 1]) / m->cum[0]) - 1);   /* bac.c:110 */
  e->lo = (uint16_t)(e->lo + (uint16_t)((range * m->cum[idx]) / m->cum[0]));           /* bac.c:111 */
  for (;;)                                                          /* bac.c:112-137 */
  {
    if (e->hi < BAC_HALF)
    {
      if ((ret = bac_emit(e, 0)) != DCO_OK)
        return ret;
    }
    else if (e->lo >= BAC_HALF)
    {
      if ((ret = bac_emit(e, 1)) != DCO_OK)
        return ret;
      e->lo = (uint16_t)(e->lo - BAC_HALF);
      e->hi = (uint16_t)(e->hi - BAC_HALF);
    }
    else if (e->lo >= BAC_QUARTER && e->hi < BAC_3QUARTER)
    {
      e->pending++;
      e->lo = (uint16_t)(e->lo - BAC_QUARTER);
      e->hi = (uint16_t)(e->hi - BAC_QUARTER);
    }
    else
      break;
    e->lo = (uint16_t)(e->lo * 2);
    e->hi = (uint16_t)(2 * e->hi + 1);
  }
  return DCO_OK;
}

int dco_bac_encode(const dco_bits *in, dco_bits *out, int adaptive)
{
  bit_reader r;
  bac_model m;
  bac_enc e;
  int ret;
  bac_model_init(&m);                                               /* bac.c:150 */
  e.lo = 0;                                                         /* bac.c:86-91 */
  e.hi = BAC_TOP;
  e.pending = 0;
  e.out = out;
  rd_init(&r, in);
  while (!rd_eof(&r))                                               /* bac.c:152-162 */
  {
    uint64_t bit;
    unsigned idx;
    if ((ret = get_bits(&r, 1, &bit)) != DCO_OK)
      return ret;
    idx = m.sym_to_idx[bit];
    if ((ret = bac_encode_symbol(&e, &m, idx)) != DCO_OK)
      return ret;
    if (adaptive)
      bac_model_update(&m, idx);
  }
  if ((ret = bac_encode_symbol(&e, &m, BAC_EOF_IDX)) != DCO_OK)     /* bac.c:163 */
    return ret;
  e.pending++;                                                      /* bac.c:141-145 */
  return bac_emit(&e, e.lo < BAC_QUARTER ? 0 : 1);
}

typedef struct bac_dec
{
  uint16_t lo, hi, value;                                           /* bac.c:83,168 */
  unsigned after_eof;                                               /* bac.c:169 */
  bit_reader r;
} bac_dec;

static int bac_next_bit(bac_dec *d, unsigned *bit)                  /* ReadBitSpecial, bac.c:171-186 */
{
  uint64_t b;
  int ret;
  if (rd_eof(&d->r))
  {
    if (d->after_eof > 0)
    {
      d->after_eof--;
      *bit = 0;
      return DCO_OK;
    }
    return DCO_ERR_INVALID_FORMAT;
  }
  if ((ret = get_bits(&d->r, 1, &b)) != DCO_OK)
    return ret;
  *bit = (unsigned)b;
  return DCO_OK;
}

int dco_bac_decode(const dco_bits *in, dco_bits *out, int adaptive)
{
  bac_model m;
  bac_dec d;
  unsigned i, bit;
  int ret;
  bac_model_init(&m);                                               /* bac.c:247 */
  rd_init(&d.r, in);
  d.value = 0;                                                      /* bac.c:188-204 */
  d.after_eof = 16 - 2;
  for (i = 0; i < 16; i++)
  {
    if ((ret = bac_next_bit(&d, &bit)) != DCO_OK)
      return ret;
    d.value = (uint16_t)(2 * d.value + bit);
  }
  d.lo = 0;
  d.hi = BAC_TOP;
  for (;;)                                                          /* bac.c:250-262 */
  {
    const uint64_t range = (uint64_t)(uint16_t)(d.hi - d.lo) + 1;  /* bac.c:208 */
    const uint16_t cf = (uint16_t)((((uint64_t)(uint16_t)(d.value - d.lo) + 1) * m.cum[0] - 1) / range);  /* bac.c:209 */
    unsigned idx;
    for (idx = 1; m.cum[idx] > cf; idx++)                           /* bac.c:210 */
      ;
    d.hi = (uint16_t)(d.lo + (uint16_t)((range * m.cum[idx - 1]) / m.cum[0]) - 1);  /* bac.c:211 */
    d.lo = (uint16_t)(d.lo + (uint16_t)((range * m.cum[idx]) / m.cum[0]));          /* bac.c:212 */
    for (;;)                                                        /* bac.c:213-240 */
    {
      if (d.hi < BAC_HALF)
      {
      }
      else if (d.lo >= BAC_HALF)
      {
        d.value = (uint16_t)(d.value - BAC_HALF);
        d.lo = (uint16_t)(d.lo - BAC_HALF);
        d.hi = (uint16_t)(d.hi - BAC_HALF);
      }
      else if (d.lo >= BAC_QUARTER && d.hi < BAC_3QUARTER)
      {
        d.value = (uint16_t)(d.value - BAC_QUARTER);
        d.lo = (uint16_t)(d.lo - BAC_QUARTER);
        d.hi = (uint16_t)(d.hi - BAC_QUARTER);
      }
      else
        break;
      d.lo = (uint16_t)(d.lo * 2);
      d.hi = (uint16_t)(2 * d.hi + 1);
      if ((ret = bac_next_bit(&d, &bit)) != DCO_OK)
        return ret;
      d.value = (uint16_t)(2 * d.value + bit);
    }
    if (idx == BAC_EOF_IDX)                                         /* bac.c:256 */
      break;
    if ((ret = put_bits(out, m.idx_to_sym[idx], 1)) != DCO_OK)      /* bac.c:258-259 */
      return ret;
    if (adaptive)                                                   /* bac.c:260-261 */
      bac_model_update(&m, idx);
  }
  return DCO_OK;
}

/* ------------------------------------------------------------------------------------------- */
/* lzmh: DCLib/src/lzmh.c -- LZ77 (128-byte window, length 3..274, 4 recent offsets) + adaptive-rank
   static prefix code for literals.  Restated on linear stream positions: the reference's 403-byte
   ring (lzmh.c:68) always holds min(pos,128) bytes of history and min(remaining,275) bytes of
   look-ahead, so ring indices and linear positions select the same bytes. */

#define LZ_WINDOW     128             /* lzmh.c:54 */
#define LZ_MAXLEN     274             /* lzmh.c:56 */
#define LZ_RING       (LZ_WINDOW + LZ_MAXLEN + 1)   /* lzmh.c:68 */
#define LZ_LIST       48              /* lzmh.c:61 */
#define LZ_TREE       19              /* lzmh.c:85 */

static const uint8_t lz_tree_code[LZ_TREE] = {                     /* lzmh.c:86-106 */
  0x0F, 0x0E, 0x0D, 0x0C, 0x17, 0x16, 0x15, 0x14, 0x13, 0x25, 0x24, 0x23, 0x22, 0x43, 0x42, 0x83, 0x82, 0x81, 0x80
};
static const uint8_t lz_tree_len[LZ_TREE] = {
  4, 4, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 8, 8, 8, 8
};

typedef struct lz_list
{
  uint8_t sym[LZ_LIST];
  int count[LZ_LIST];
} lz_list;

/* Literal bookkeeping shared by encoder (lzmh.c:297-322) and decoder (lzmh.c:463-477): returns the
   rank the byte had BEFORE the update, or -1 if it was not in the list. */
static int lz_list_touch(lz_list *l, uint8_t byte)
{
  int pos = 0, rank = -1;
  while (pos < LZ_LIST && l->count[pos] > 0)
  {
    if (l->sym[pos] == byte)
    {
      rank = pos;
      if (l->count[pos] < 65535)
      {
        const int c = l->count[pos] + 1;
        while (pos > 0 && c > l->count[pos - 1])                    /* bubble past strictly smaller counts */
        {
          l->sym[pos] = l->sym[pos - 1];
          l->count[pos] = l->count[pos - 1];
          pos--;
        }
        l->count[pos] = c;
        l->sym[pos] = byte;
      }
      return rank;
    }
    pos++;
  }
  if (pos < LZ_LIST)                                                /* lzmh.c:318-322 */
  {
    l->sym[pos] = byte;
    l->count[pos] = 1;
  }
  return rank;
}

int dco_lzmh_encode(const dco_bits *in, dco_bits *out)
{
  const uint64_t n = in->nbits / 8;
  const uint8_t *src = in->data;
  uint64_t pos = 0;
  int recent[4] = { 0, 0, 0, 0 };                                   /* lzmh.c:157-159 */
  lz_list list;
  int ret;
  memset(&list, 0, sizeof(list));                                   /* lzmh.c:151-154 */
  if (in->nbits & 7)                                                /* 8-bit reads, lzmh.c:165 */
    return DCO_ERR_LIBRARY_CALL;
  if (n == LZ_RING)                                                 /* lzmh.c:169-173: the initial fill wraps write==read at EOF */
    return DCO_OK;
  while (pos < n)                                                   /* lzmh.c:173 */
  {
    const int maxoff = pos < LZ_WINDOW ? (int)pos : LZ_WINDOW;      /* lzmh.c:175-179 */
    const int maxlen = (n - pos) > LZ_MAXLEN ? LZ_MAXLEN : (int)(n - pos);   /* lzmh.c:180-190 */
    int best = 2, bestoff = 0, off;                                 /* lzmh.c:192-193 */
    uint64_t code = 0;
    unsigned codelen = 0;
    for (off = 1; off <= maxoff && best < maxlen; off++)            /* lzmh.c:196-214 */
    {
      const uint8_t *cand = src + pos - off;
      if (cand[0] == src[pos] && cand[best] == src[pos + best])
      {
        int len = 1;
        while (len < maxlen && cand[len] == src[pos + len])
          len++;
        if (len > best)
        {
          best = len;
          bestoff = off;
        }
      }
    }
    if (best >= 3)                                                  /* lzmh.c:216-284 */
    {
      int k;
      for (k = 0; k < 4 && recent[k] != bestoff; k++)
        ;
      switch (k)
      {
        case 0: code = 0x06; codelen = 4; break;                    /* 0110 */
        case 1: code = 0x0E; codelen = 5; break;                    /* 01110 */
        case 2: code = 0x1E; codelen = 6; break;                    /* 011110 */
        case 3: code = 0x1F; codelen = 6; break;                    /* 011111 */
        default: code = 0x100 | (uint64_t)(bestoff - 1); codelen = 10; k = 3; break;   /* 010 + 7 bit */
      }
      for (; k > 0; k--)                                            /* move to front */
        recent[k] = recent[k - 1];
      recent[0] = bestoff;
      if (best < 11)                                                /* lzmh.c:263-278 */
      {
        code = (code << 4) | (uint64_t)(best - 3);
        codelen += 4;
      }
      else if (best < 19)
      {
        code = (code << 5) | 0x10 | (uint64_t)(best - 11);
        codelen += 5;
      }
      else
      {
        code = (code << 10) | 0x300 | (uint64_t)(best - 19);
        codelen += 10;
      }
      pos += (uint64_t)best;
    }
    else                                                            /* lzmh.c:285-335 */
    {
      const uint8_t byte = src[pos++];
      const int rank = lz_list_touch(&list, byte);
      if (rank >= 0 && rank < LZ_TREE)
      {
        code = lz_tree_code[rank];
        codelen = lz_tree_len[rank];
      }
      else
      {
        code = byte;                                                /* 00 + 8 bit */
        codelen = 10;
      }
    }
    if ((ret = put_bits(out, code, codelen)) != DCO_OK)             /* lzmh.c:337-339 */
      return ret;
  }
  return DCO_OK;
}

int dco_lzmh_decode(const dco_bits *in, dco_bits *out)
{
  bit_reader r;
  uint32_t window = 0;                                              /* code_sym, lzmh.c:393 */
  int have = 0;                                                     /* code_length (int8_t), lzmh.c:394 */
  int recent[4] = { 0, 0, 0, 0 };
  uint8_t hist[LZ_WINDOW];                                          /* lzmh.c:396 (static: stale bytes are zero on first use) */
  unsigned hpos = 0;
  lz_list list;
  int ret;
  memset(&list, 0, sizeof(list));
  memset(hist, 0, sizeof(hist));
  rd_init(&r, in);
  do
  {
    while (!rd_eof(&r) && (32 - have) >= 8)                         /* lzmh.c:409-413 */
    {
      uint64_t b;
      have = (int8_t)(have + 1);
      if ((ret = get_bits(&r, 1, &b)) != DCO_OK)
        return ret;
      window |= (uint32_t)b << (32 - have);
    }
    if (window & 0x80000000u)                                       /* lzmh.c:414-450 */
    {
      int i;
      for (i = 0; i < LZ_TREE; i++)
        if (have >= lz_tree_len[i] && (window >> (32 - lz_tree_len[i])) == lz_tree_code[i])
          break;
      if (i == LZ_TREE)                                             /* lzmh.c:447-449 */
        return DCO_OK;
      have = (int8_t)(have - lz_tree_len[i]);
      window <<= lz_tree_len[i];
      {
        const uint8_t byte = list.sym[i];
        if ((ret = put_bits(out, byte, 8)) != DCO_OK)
          return ret;
        hist[hpos] = byte;
        hpos = (hpos + 1) % LZ_WINDOW;
        if (list.count[i] < 65535)                                  /* lzmh.c:431-442 */
        {
          const int c = list.count[i] + 1;
          while (i > 0 && c > list.count[i - 1])
          {
            list.sym[i] = list.sym[i - 1];
            i--;
          }
          list.count[i] = (uint16_t)c;
          list.sym[i] = byte;
        }
      }
    }
    else if ((window & 0x40000000u) == 0)                           /* lzmh.c:451-479 */
    {
      const uint8_t byte = (uint8_t)((window >> 22) & 0xFF);
      int pos;
      have = (int8_t)(have - 10);
      window <<= 10;
      if ((ret = put_bits(out, byte, 8)) != DCO_OK)
        return ret;
      hist[hpos] = byte;
      hpos = (hpos + 1) % LZ_WINDOW;
      for (pos = 0; pos < LZ_LIST && list.count[pos] > 0 && list.sym[pos] != byte; pos++)   /* lzmh.c:463 */
        ;
      if (pos < LZ_LIST && list.count[pos] < 65535)                 /* lzmh.c:464-477 */
      {
        const int c = list.count[pos] + 1;
        while (pos > 0 && c > list.count[pos - 1])
        {
          list.sym[pos] = list.sym[pos - 1];
          list.count[pos] = list.count[pos - 1];
          pos--;
        }
        list.count[pos] = c;
        list.sym[pos] = byte;
      }
    }
    else                                                            /* lzmh.c:480-569 */
    {
      int off, len, k;
      have = (int8_t)(have - 2);
      window <<= 2;
      if ((window & 0x80000000u) == 0)                              /* explicit 7-bit offset, lzmh.c:488-497 */
      {
        off = (int)((window >> 24) & 0x7F) + 1;
        have = (int8_t)(have - 8);
        window <<= 8;
        k = 3;
      }
      else                                                          /* recent offsets, lzmh.c:498-536 */
      {
        have = (int8_t)(have - 1);
        window <<= 1;
        if ((window & 0x80000000u) == 0)
          k = 0;
        else
        {
          have = (int8_t)(have - 1);
          window <<= 1;
          if ((window & 0x80000000u) == 0)
            k = 1;
          else
          {
            have = (int8_t)(have - 1);
            window <<= 1;
            k = (window & 0x80000000u) == 0 ? 2 : 3;
          }
        }
        have = (int8_t)(have - 1);
        window <<= 1;
        off = recent[k];
      }
      for (; k > 0; k--)
        recent[k] = recent[k - 1];
      recent[0] = off;
      if ((window & 0x80000000u) == 0)                              /* lzmh.c:540-558 */
      {
        len = (int)((window >> 28) & 0x7) + 3;
        have = (int8_t)(have - 4);
        window <<= 4;
      }
      else
      {
        have = (int8_t)(have - 1);
        window <<= 1;
        if ((window & 0x80000000u) == 0)
        {
          len = (int)((window >> 28) & 0x7) + 11;
          have = (int8_t)(have - 4);
          window <<= 4;
        }
        else
        {
          len = (int)((window >> 23) & 0xFF) + 19;
          have = (int8_t)(have - 9);
          window <<= 9;
        }
      }
      for (k = 0; k < len; k++)                                     /* lzmh.c:560-568 */
      {
        int from = (int)hpos - off;
        uint8_t byte;
        while (from < 0)
          from += LZ_WINDOW;
        byte = hist[from % LZ_WINDOW];
        if ((ret = put_bits(out, byte, 8)) != DCO_OK)
          return ret;
        hist[hpos] = byte;
        hpos = (hpos + 1) % LZ_WINDOW;
      }
    }
  } while (!rd_eof(&r) || window > 0);                              /* lzmh.c:571 */
  return DCO_OK;
}

/* ------------------------------------------------------------------------------------------- */
/* Copy, DCLib/src/copy.c:11-29: blocks of block_bits bits are read and written unchanged until the input ends
   (copy.c:22-26); a short last block is a failed read (READ_BITS_CHECKED, io_macros.h:13-27), after the whole blocks
   before it have been written.  Blocks may be any size: they are moved 64 bits at a time here. */
int dco_copy(const dco_bits *in, dco_bits *out, unsigned block_bits)
{
  bit_reader r;
  if (block_bits == 0)
    return DCO_ERR_INVALID_VALUE;
  rd_init(&r, in);
  while (!rd_eof(&r))
  {
    uint64_t left = block_bits;
    if (r.nbits - r.pos < left)                                     /* copy.c:24 */
      return DCO_ERR_LIBRARY_CALL;
    while (left > 0)
    {
      const unsigned n = left > 64 ? 64u : (unsigned)left;
      uint64_t v;
      int ret;
      if ((ret = get_bits(&r, n, &v)) != DCO_OK || (ret = put_bits(out, v, n)) != DCO_OK)   /* copy.c:24-25 */
        return ret;
      left -= n;
    }
  }
  return DCO_OK;
}

static unsigned g_block_bits = 8;                                   /* enc_dec.c:187-197: blocksize defaults to 8 */
void dco_set_block_size(unsigned bits) { g_block_bits = bits ? bits : 8; }

/* Chains (DCCLI/src/cli.c:430-466: every stage runs to completion, exact-bit hand-off) */

static int run_stage(int id, const dco_bits *in, dco_bits *out, unsigned vs, float nf, int adaptive)
{
  switch (id)
  {
    case DCO_NORMALIZE:   return dco_normalize(in, out, vs, nf);
    case DCO_DENORMALIZE: return dco_denormalize(in, out, vs, nf);
    case DCO_DIFF_ENC:    return dco_diff_encode(in, out, vs);
    case DCO_DIFF_DEC:    return dco_diff_decode(in, out, vs);
    case DCO_SEG_ENC:     return dco_seg_encode(in, out, vs);
    case DCO_SEG_DEC:     return dco_seg_decode(in, out, vs);
    case DCO_BAC_ENC:     return dco_bac_encode(in, out, adaptive);
    case DCO_BAC_DEC:     return dco_bac_decode(in, out, adaptive);
    case DCO_LZMH_ENC:    return dco_lzmh_encode(in, out);
    case DCO_LZMH_DEC:    return dco_lzmh_decode(in, out);
    case DCO_COPY:        return dco_copy(in, out, g_block_bits);
    default:              return DCO_ERR_INVALID_VALUE;
  }
}

static int run_chain_bits(const int *stages, int nstages, dco_bits *cur, unsigned vs, float nf, int adaptive)
{
  int i, ret = DCO_OK;
  for (i = 0; i < nstages && ret == DCO_OK; i++)
  {
    dco_bits next;
    dco_bits_init(&next);
    ret = run_stage(stages[i], cur, &next, vs, nf, adaptive);
    dco_bits_free(cur);
    *cur = next;
  }
  return ret;
}

int dco_run_chain(const int *stages, int nstages, const uint8_t *in, uint64_t in_bits,
                  unsigned vs, float nf, int adaptive,
                  uint8_t *out, uint64_t out_cap, uint64_t *out_bits)
{
  dco_bits cur;
  int ret;
  dco_bits_init(&cur);
  if ((ret = dco_bits_from_bytes(&cur, in, in_bits)) != DCO_OK)
    return ret;
  ret = run_chain_bits(stages, nstages, &cur, vs, nf, adaptive);
  *out_bits = cur.nbits;
  if ((cur.nbits + 7) / 8 > out_cap)
  {
    dco_bits_free(&cur);
    return DCO_ERR_MEMORY;
  }
  if (cur.nbits)
    memcpy(out, cur.data, (size_t)((cur.nbits + 7) / 8));
  dco_bits_free(&cur);
  return ret;
}

int dco_dega_encode(const float *samples, uint64_t n, unsigned vs, float nf, int adaptive,
                    uint8_t *out, uint64_t out_cap, uint64_t *out_bits)
{
  static const int chain[4] = { DCO_NORMALIZE, DCO_DIFF_ENC, DCO_SEG_ENC, DCO_BAC_ENC };
  return dco_run_chain(chain, 4, (const uint8_t *)samples, n * 32, vs, nf, adaptive, out, out_cap, out_bits);
}

int dco_dega_decode(const uint8_t *in, uint64_t in_bits, unsigned vs, float nf, int adaptive,
                    float *out, uint64_t out_cap_samples, uint64_t *n_out)
{
  static const int chain[4] = { DCO_BAC_DEC, DCO_SEG_DEC, DCO_DIFF_DEC, DCO_DENORMALIZE };
  uint64_t bits = 0;
  const int ret = dco_run_chain(chain, 4, in, in_bits, vs, nf, adaptive, (uint8_t *)out, out_cap_samples * 4, &bits);
  *n_out = bits / 32;
  return ret;
}

/* csv.c:13-44.  One character per pass (csv.c:22); a field ends at the separator, at a newline or with the input
   (csv.c:23); the last character of the input joins the wanted column's text whatever it is (csv.c:25-26); the
   wanted column goes through strtof (csv.c:27-35); a newline resets the column counter (csv.c:41-42).  The
   reference's text buffer holds 48 characters; longer fields are undefined there and cut here. */
uint64_t dco_csv_decode(const uint8_t *text, uint64_t nbytes, unsigned column, char separator, float *out, uint64_t cap)
{
  char buffer[64];
  size_t used = 0;
  unsigned col = 1;
  uint64_t i, n = 0;
  for (i = 0; i < nbytes; i++)
  {
    const char ch = (char)text[i];
    const int at_end = i + 1 == nbytes;
    if (ch == separator || ch == '\n' || at_end)
    {
      if (at_end && col == column && used < sizeof(buffer) - 1)
        buffer[used++] = ch;
      if (col == column)
      {
        buffer[used] = '\0';
        if (n < cap)
          out[n] = strtof(buffer, NULL);
        n++;
        used = 0;
      }
      col++;
    }
    else if (col == column && used < sizeof(buffer) - 1)
      buffer[used++] = ch;
    if (ch == '\n')
      col = 1;
  }
  return n;
}

/* `encode csv column=.. separator_char=.. num_decimal_places=..` (DCLib/src/csv.c:46-65, WriteCSV): per value,
   column - 1 separator characters (csv.c:55-58), then printf's "%.*f" of the value with num_decimal_places digits and
   a newline (csv.c:59-61).  printf is the algorithm here as it is there: the exact binary value, rounded to the
   requested digits half-to-even, "-" for a set sign bit (also of zero and nan), "inf" / "nan" for the non-finite.
   Returns the number of bytes the text has; writes at most cap of them. */
uint64_t dco_csv_encode(const float *in, uint64_t n, unsigned column, char separator, unsigned decimals, uint8_t *out, uint64_t cap)
{
  uint64_t i, used = 0;
  for (i = 0; i < n; i++)
  {
    char buffer[64];
    unsigned k;
    int len, j;
    for (k = 1; k < column; k++, used++)
      if (used < cap)
        out[used] = (uint8_t)separator;
    len = snprintf(buffer, sizeof(buffer), "%.*f\n", (int)decimals, (double)in[i]);
    for (j = 0; j < len; j++, used++)
      if (used < cap)
        out[used] = (uint8_t)buffer[j];
  }
  return used;
}

/* aggregate.c:9-27.  The outer loop runs while input remains (aggregate.c:11); each pass starts a sum at 0
   (aggregate.c:13), adds up to num_values values in order (aggregate.c:16-20) and stops early when the input
   ends (aggregate.c:21-22).  volatile keeps the adds separately rounded and in order whatever the flags. */
uint64_t dco_aggregate(const float *in, uint64_t n, unsigned num_values, float *out)
{
  uint64_t i = 0, g = 0;
  while (i < n)
  {
    volatile float sum = 0.0f;
    unsigned k;
    for (k = 0; k < num_values && i < n; k++, i++)
      sum = sum + in[i];
    out[g++] = sum;
  }
  return g;
}

int64_t dco_dega_encode_batch(const float *samples, uint64_t nstreams, uint64_t len,
                              unsigned vs, float nf, int adaptive)
{
  static const int chain[4] = { DCO_NORMALIZE, DCO_DIFF_ENC, DCO_SEG_ENC, DCO_BAC_ENC };
  int64_t total = 0;
  uint64_t s;
  for (s = 0; s < nstreams; s++)
  {
    dco_bits cur;
    int ret;
    dco_bits_init(&cur);
    if ((ret = dco_bits_from_bytes(&cur, (const uint8_t *)(samples + s * len), len * 32)) != DCO_OK)
      return ret;
    ret = run_chain_bits(chain, 4, &cur, vs, nf, adaptive);
    total += (int64_t)((cur.nbits + 7) / 8);
    dco_bits_free(&cur);
    if (ret != DCO_OK)
      return ret;
  }
  return total;
}
