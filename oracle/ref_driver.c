/* ref_driver.c -- in-process driver for the UNMODIFIED reference codecs (test infrastructure only).
 *
 * Linked by oracle/Makefile against objects compiled directly from
 * /root/reference/DataCompressor/{DCIOLib,DCLib}/src, every .c file (release flags of
 * common/build/gcc/common.mak:11-20) into oracle/_ref/libdcref.so.  It replays what DCCLI's stage
 * loop does (DCCLI/src/cli.c:430-466): every stage runs to completion on in-memory bit buffers and
 * the write->read mode switch keeps the fractional last byte (cli.c:212-223), so stage N+1 sees the
 * exact bit count stage N wrote.  None of this file is copied from the reference; it only calls its
 * public API (DCLib/inc/enc_dec.h, DCIOLib/inc/bit_file_buffer.h, DCIOLib/inc/file_buffer.h).
 *
 * The reference codecs keep state in statics (bac.c:33-37,83-84; lzmh.c:139,143,396): one call at a
 * time per process.  Parallel CPU baselines fork one process per worker.
 */
#include "err_codes.h"
#include "enc_dec.h"

#include <stdlib.h>
#include <string.h>

#define REF_MAX_STAGES 16

typedef struct mem_bits
{
  file_buffer_t *fb;
  bit_file_buffer_t *bb;
} mem_bits;

static int mem_bits_open(mem_bits *m, file_buffer_mode_t mode)
{
  m->fb = AllocateFileBuffer();
  m->bb = AllocateBitFileBuffer();
  if (m->fb == NULL || m->bb == NULL)
    return ERROR_MEMORY;
  if (InitFileBufferInMemory(m->fb, mode, 2048) != NO_ERROR)
    return ERROR_LIBRARY_INIT;
  InitBitFileBuffer(m->bb, m->fb);
  return NO_ERROR;
}

static void mem_bits_close(mem_bits *m)
{
  if (m->fb != NULL)
  {
    UninitFileBuffer(m->fb);
    FreeFileBuffer(m->fb);
  }
  if (m->bb != NULL)
    FreeBitFileBuffer(m->bb);
  m->fb = NULL;
  m->bb = NULL;
}

static size_t g_block_size_bits = 8;   /* `blocksize` of the copy stages (options_t.block_size_bits), default as enc_dec.c:187-197 */
void ref_set_block_size(unsigned bits) { g_block_size_bits = bits ? bits : 8; }

/* Runs `nstages` reference stages back to back.
   names[i]   : encoder name as given on the DCCLI command line ("normalize", "diff", "seg", "bac", "lzmh", ...)
   encode[i]  : 1 = "encode <name>", 0 = "decode <name>"
   vs, nf, adaptive are applied to every stage (DCCLI takes them per stage; the tests always repeat them).
   Input: `in_bits` bits, MSB-first, in `in`.  Output: exact bit count in *out_bits, bytes zero padded.
   Returns the failing stage's (negative) code, or 0.  *failed_stage = index of the failing stage or -1. */
long long ref_run_chain(const char *const *names, const int *encode, int nstages,
                        const unsigned char *in, unsigned long long in_bits,
                        unsigned vs, float nf, int adaptive,
                        unsigned char *out, unsigned long long out_cap, unsigned long long *out_bits,
                        int *failed_stage)
{
  mem_bits a, b;
  mem_bits *rd = &a, *wr = &b;
  io_int_t ret = NO_ERROR;
  int i;
  *out_bits = 0;
  *failed_stage = -1;
  a.fb = b.fb = NULL;
  a.bb = b.bb = NULL;
  if (nstages < 1 || nstages > REF_MAX_STAGES)
    return ERROR_INVALID_VALUE;
  if (mem_bits_open(rd, FBM_WRITING) != NO_ERROR || mem_bits_open(wr, FBM_WRITING) != NO_ERROR)
  {
    mem_bits_close(&a);
    mem_bits_close(&b);
    return ERROR_LIBRARY_INIT;
  }
  /* Load the input with its exact bit count, then flip to reading (what SwitchTempBuffers does). */
  if (in_bits > 0 && WriteBitFileBuffer(rd->bb, in, (size_t)in_bits) != (io_int_t)in_bits)
  {
    mem_bits_close(&a);
    mem_bits_close(&b);
    return ERROR_LIBRARY_CALL;
  }
  SetBitFileBufferMode(rd->bb, FBM_READING);
  for (i = 0; i < nstages; i++)
  {
    options_t opt;
    const enc_dec_t *ed = GetEncoder(names[i]);
    enc_dec_function_t *fn;
    mem_bits *tmp;
    if (ed == NULL)
    {
      ret = ERROR_INVALID_VALUE;
      *failed_stage = i;
      break;
    }
    memset(&opt, 0, sizeof(opt));
    SetDefaultOptions(&opt);
    opt.error_log_file = NULL;              /* LOG tolerates NULL, common/inc/log.h:10-13 */
    opt.encode = encode[i];
    opt.encoder_decoder = ed;
    opt.value_size_bits = vs;
    opt.normalization_factor = nf;
    opt.adaptive = adaptive;
    opt.block_size_bits = g_block_size_bits;
    fn = encode[i] ? ed->encoder : ed->decoder;
    if (fn == NULL)
    {
      ret = ERROR_INVALID_MODE;
      *failed_stage = i;
      break;
    }
    ret = (*fn)(rd->bb, wr->bb, &opt);
    /* the written buffer becomes the next stage's input, also after an error (partial output) */
    SetBitFileBufferMode(wr->bb, FBM_READING);
    tmp = rd;
    rd = wr;
    wr = tmp;
    if (ret != NO_ERROR)
    {
      *failed_stage = i;
      break;
    }
    if (i + 1 < nstages)
      ResetBitFileBuffer(wr->bb, FBM_WRITING);
  }
  /* Drain the final buffer: bytes first, then the fractional byte. */
  {
    unsigned long long total = 0;
    io_int_t got;
    for (;;)
    {
      unsigned char chunk[4096];
      if (EndOfBitFileBuffer(rd->bb))
        break;
      got = ReadBitFileBuffer(rd->bb, chunk, sizeof(chunk) * 8);
      if (got <= 0)
        break;
      if ((total + (unsigned long long)got + 7) / 8 > out_cap)
      {
        ret = ERROR_MEMORY;
        break;
      }
      memcpy(out + total / 8, chunk, (size_t)((got + 7) / 8));   /* total is a multiple of 8 until the last chunk */
      total += (unsigned long long)got;
      if ((size_t)got != sizeof(chunk) * 8)
        break;
    }
    *out_bits = total;
  }
  mem_bits_close(&a);
  mem_bits_close(&b);
  return ret;
}

/* `encode aggregate num_values=k` (DCLib/src/aggregate.c:9-27) on n float32 values through the real codec.
   Returns the number of sums written to out (capacity out_cap floats) or a negative error code. */
long long ref_aggregate(const float *in, unsigned long long n, unsigned num_values, float *out, unsigned long long out_cap)
{
  mem_bits a, b;
  options_t opt;
  const enc_dec_t *ed = GetEncoder("aggregate");
  io_int_t ret;
  unsigned long long got_total = 0;
  a.fb = b.fb = NULL;
  a.bb = b.bb = NULL;
  if (ed == NULL || ed->encoder == NULL)
    return ERROR_INVALID_VALUE;
  if (mem_bits_open(&a, FBM_WRITING) != NO_ERROR || mem_bits_open(&b, FBM_WRITING) != NO_ERROR)
  {
    mem_bits_close(&a);
    mem_bits_close(&b);
    return ERROR_LIBRARY_INIT;
  }
  if (n > 0 && WriteBitFileBuffer(a.bb, (const unsigned char *)in, (size_t)(n * 32)) != (io_int_t)(n * 32))
  {
    mem_bits_close(&a);
    mem_bits_close(&b);
    return ERROR_LIBRARY_CALL;
  }
  SetBitFileBufferMode(a.bb, FBM_READING);
  memset(&opt, 0, sizeof(opt));
  SetDefaultOptions(&opt);
  opt.error_log_file = NULL;
  opt.encode = 1;
  opt.encoder_decoder = ed;
  opt.num_values = num_values;
  ret = (*ed->encoder)(a.bb, b.bb, &opt);
  SetBitFileBufferMode(b.bb, FBM_READING);
  while (ret == NO_ERROR && !EndOfBitFileBuffer(b.bb) && got_total < out_cap)
  {
    io_int_t got = ReadBitFileBuffer(b.bb, (unsigned char *)(out + got_total), 32);
    if (got != 32)
      break;
    got_total++;
  }
  mem_bits_close(&a);
  mem_bits_close(&b);
  return ret != NO_ERROR ? (long long)ret : (long long)got_total;
}

/* `encode csv column=.. separator_char=.. num_decimal_places=..` (DCLib/src/csv.c:46-65) on n floats through the real
   codec.  Returns the number of text bytes written to out (capacity out_cap) or a negative error code. */
long long ref_csv_encode(const float *values, unsigned long long n, unsigned column, char separator, unsigned decimals,
                         unsigned char *out, unsigned long long out_cap)
{
  mem_bits a, b;
  options_t opt;
  const enc_dec_t *ed = GetEncoder("csv");
  io_int_t ret;
  unsigned long long got_total = 0;
  a.fb = b.fb = NULL;
  a.bb = b.bb = NULL;
  if (ed == NULL || ed->encoder == NULL)
    return ERROR_INVALID_VALUE;
  if (mem_bits_open(&a, FBM_WRITING) != NO_ERROR || mem_bits_open(&b, FBM_WRITING) != NO_ERROR)
  {
    mem_bits_close(&a);
    mem_bits_close(&b);
    return ERROR_LIBRARY_INIT;
  }
  if (n > 0 && WriteBitFileBuffer(a.bb, (const unsigned char *)values, (size_t)(n * 32)) != (io_int_t)(n * 32))
  {
    mem_bits_close(&a);
    mem_bits_close(&b);
    return ERROR_LIBRARY_CALL;
  }
  SetBitFileBufferMode(a.bb, FBM_READING);
  memset(&opt, 0, sizeof(opt));
  SetDefaultOptions(&opt);
  opt.error_log_file = NULL;
  opt.encode = 1;
  opt.encoder_decoder = ed;
  opt.column = column;
  opt.separator_char = separator;
  opt.num_decimal_places = decimals;
  ret = (*ed->encoder)(a.bb, b.bb, &opt);
  SetBitFileBufferMode(b.bb, FBM_READING);
  while (ret == NO_ERROR && !EndOfBitFileBuffer(b.bb) && got_total < out_cap)
  {
    io_int_t got = ReadBitFileBuffer(b.bb, out + got_total, 8);
    if (got != 8)
      break;
    got_total++;
  }
  mem_bits_close(&a);
  mem_bits_close(&b);
  return ret != NO_ERROR ? (long long)ret : (long long)got_total;
}

/* `decode csv column=.. separator=..` (DCLib/src/csv.c:13-44) on nbytes of text through the real codec.
   Returns the number of floats written to out (capacity out_cap) or a negative error code. */
long long ref_csv_decode(const unsigned char *text, unsigned long long nbytes, unsigned column, char separator, float *out,
                         unsigned long long out_cap)
{
  mem_bits a, b;
  options_t opt;
  const enc_dec_t *ed = GetEncoder("csv");
  io_int_t ret;
  unsigned long long got_total = 0;
  a.fb = b.fb = NULL;
  a.bb = b.bb = NULL;
  if (ed == NULL || ed->decoder == NULL)
    return ERROR_INVALID_VALUE;
  if (mem_bits_open(&a, FBM_WRITING) != NO_ERROR || mem_bits_open(&b, FBM_WRITING) != NO_ERROR)
  {
    mem_bits_close(&a);
    mem_bits_close(&b);
    return ERROR_LIBRARY_INIT;
  }
  if (nbytes > 0 && WriteBitFileBuffer(a.bb, text, (size_t)(nbytes * 8)) != (io_int_t)(nbytes * 8))
  {
    mem_bits_close(&a);
    mem_bits_close(&b);
    return ERROR_LIBRARY_CALL;
  }
  SetBitFileBufferMode(a.bb, FBM_READING);
  memset(&opt, 0, sizeof(opt));
  SetDefaultOptions(&opt);
  opt.error_log_file = NULL;
  opt.encode = 0;
  opt.encoder_decoder = ed;
  opt.column = column;
  opt.separator_char = separator;
  ret = (*ed->decoder)(a.bb, b.bb, &opt);
  SetBitFileBufferMode(b.bb, FBM_READING);
  while (ret == NO_ERROR && !EndOfBitFileBuffer(b.bb) && got_total < out_cap)
  {
    io_int_t got = ReadBitFileBuffer(b.bb, (unsigned char *)(out + got_total), 32);
    if (got != 32)
      break;
    got_total++;
  }
  mem_bits_close(&a);
  mem_bits_close(&b);
  return ret != NO_ERROR ? (long long)ret : (long long)got_total;
}
