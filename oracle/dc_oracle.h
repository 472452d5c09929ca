/* dc_oracle.h -- CPU restatement ("oracle") of the Data Compressor hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the CPU baseline.  The CUDA path never calls into it.
 *
 * Every function restates one reference function on plain byte arrays; citations are to
 * /root/reference/DataCompressor/<path>:<lines>.  Parity of this file is pinned against the real
 * reference (oracle/_ref, built from the reference's own sources by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py and against the committed golden vectors in tests/golden/.
 *
 * Bit streams are MSB-first inside each byte (DCIOLib/src/bit_file_buffer.c:183,244,299-307).
 * A stage reads an exact bit count (the in-chain hand-off keeps the fractional last byte,
 * bit_file_buffer.c:135-141,169-176) and reports the exact bit count it wrote.
 */
#ifndef DC_ORACLE_H
#define DC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes: common/inc/err_codes.h:8-15,28-29 */
#define DCO_OK                    0
#define DCO_ERR_INVALID_VALUE    (-1)
#define DCO_ERR_INVALID_FORMAT   (-3)
#define DCO_ERR_MEMORY           (-6)
#define DCO_ERR_LIBRARY_CALL    (-11)

/* Growable MSB-first bit string. */
typedef struct dco_bits
{
  uint8_t *data;     /* zero-padded to a whole byte */
  size_t cap;        /* bytes allocated */
  uint64_t nbits;    /* exact number of valid bits */
} dco_bits;

void dco_bits_init(dco_bits *b);
void dco_bits_free(dco_bits *b);
int dco_bits_from_bytes(dco_bits *b, const uint8_t *bytes, uint64_t nbits);

/* One function per reference codec stage; all are bits -> bits like enc_dec_function_t
   (DCLib/inc/enc_dec.h:11).  Return DCO_OK or a negative code; on error `out` holds what the
   reference would have written so far. */
int dco_normalize(const dco_bits *in, dco_bits *out, unsigned vs, float nf);           /* normalize.c:9-27  */
int dco_denormalize(const dco_bits *in, dco_bits *out, unsigned vs, float nf);         /* normalize.c:29-42 */
int dco_diff_encode(const dco_bits *in, dco_bits *out, unsigned vs);                   /* diff.c:9-23       */
int dco_diff_decode(const dco_bits *in, dco_bits *out, unsigned vs);                   /* diff.c:25-38      */
int dco_seg_encode(const dco_bits *in, dco_bits *out, unsigned vs);                    /* seg.c:11-43       */
int dco_seg_decode(const dco_bits *in, dco_bits *out, unsigned vs);                    /* seg.c:45-95       */
int dco_bac_encode(const dco_bits *in, dco_bits *out, int adaptive);                   /* bac.c:86-166      */
int dco_bac_decode(const dco_bits *in, dco_bits *out, int adaptive);                   /* bac.c:171-264     */
int dco_lzmh_encode(const dco_bits *in, dco_bits *out);                                /* lzmh.c:130-370    */
int dco_lzmh_decode(const dco_bits *in, dco_bits *out);                                /* lzmh.c:383-574    */
int dco_copy(const dco_bits *in, dco_bits *out, unsigned block_bits);                  /* copy.c:11-29      */

/* Flat-buffer conveniences for ctypes.  `out` must hold out_cap bytes; *out_bits receives the exact
   bit count (bytes = ceil(bits/8), zero padded).  Return code as above; DCO_ERR_MEMORY if out_cap is
   too small. */
enum dco_stage_id
{
  DCO_NORMALIZE = 0, DCO_DENORMALIZE, DCO_DIFF_ENC, DCO_DIFF_DEC, DCO_SEG_ENC, DCO_SEG_DEC,
  DCO_BAC_ENC, DCO_BAC_DEC, DCO_LZMH_ENC, DCO_LZMH_DEC,
  DCO_COPY = 12   /* same number as DC_STAGE_COPY of include/dc_cuda.h (10 and 11 are that header's fused chains) */
};

/* `blocksize` of the copy stages of the chains run from now on (options_t.block_size_bits, default 8: enc_dec.c:187-197) */
void dco_set_block_size(unsigned bits);

/* Runs `nstages` stages back to back with exact-bit hand-off (what DCCLI's stage loop does,
   DCCLI/src/cli.c:430-466).  adaptive applies to the bac stages, vs/nf to the others. */
int dco_run_chain(const int *stages, int nstages, const uint8_t *in, uint64_t in_bits,
                  unsigned vs, float nf, int adaptive,
                  uint8_t *out, uint64_t out_cap, uint64_t *out_bits);

/* DEGA = normalize -> diff -> seg -> bac (readme.md:36); raw little-endian float32 in. */
int dco_dega_encode(const float *samples, uint64_t n, unsigned vs, float nf, int adaptive,
                    uint8_t *out, uint64_t out_cap, uint64_t *out_bits);
/* Inverse chain bac -> seg -> diff -> normalize; writes float32, *n_out = samples produced. */
int dco_dega_decode(const uint8_t *in, uint64_t in_bits, unsigned vs, float nf, int adaptive,
                    float *out, uint64_t out_cap_samples, uint64_t *n_out);

/* Batch helper for the CPU baseline: encodes `nstreams` uniform streams of `len` samples one after
   another (single thread, like one DCCLI process per stream) and returns the total payload bytes. */
/* csv.c:13-44 (`decode csv column=.. separator=..`, ReadCSV) on nbytes of text: writes the column's values to out
   (capacity cap) and returns how many there are. */
uint64_t dco_csv_decode(const uint8_t *text, uint64_t nbytes, unsigned column, char separator, float *out, uint64_t cap);
uint64_t dco_csv_encode(const float *in, uint64_t n, unsigned column, char separator, unsigned decimals, uint8_t *out, uint64_t cap);

/* aggregate.c:9-27 (`encode aggregate num_values=k`) on n float32 values: writes ceil(n / k) sums to out (sequential
   float32 adds from 0.0f, last group may be short) and returns that count. */
uint64_t dco_aggregate(const float *in, uint64_t n, unsigned num_values, float *out);

int64_t dco_dega_encode_batch(const float *samples, uint64_t nstreams, uint64_t len,
                              unsigned vs, float nf, int adaptive);

#ifdef __cplusplus
}
#endif

#endif
