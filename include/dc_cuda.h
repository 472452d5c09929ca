/* dc_cuda.h -- C ABI of the B200-native Data Compressor hot path (libdc_cuda.so).
 *
 * Plain C, no C++/torch types: this is what the reference's C90 host code (or a ctypes / cgo / JNI
 * stub) binds.  INTEGRATION.md shows the DCLib-side glue.  Each entry point names the reference
 * interface it replaces; paths are relative to /root/reference/DataCompressor/.
 *
 * Conventions (same as the reference):
 *   - return NO_ERROR (0) or a negative code of common/inc/err_codes.h:8-29
 *   - bit streams are MSB-first within a byte (DCIOLib/src/bit_file_buffer.c:183,244,299-307);
 *     a stream of `nbits` bits occupies ceil(nbits/8) bytes, zero padded, exactly the bytes DCCLI
 *     writes to its output file (bit_file_buffer.c:310-333)
 *   - a "stream" is one independent reference invocation (one DCCLI run): fresh BAC model,
 *     last_value = 0 (bac.c:150-151, diff.c:11).  A channel may be cut into resumable blocks; every
 *     (channel, block) is then one stream.
 *   - not re-entrant per context; one context per host thread / per GPU.
 */
#ifndef DC_CUDA_H
#define DC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* common/inc/err_codes.h:8-15,28-29 */
#define DC_OK                       0
#define DC_ERR_INVALID_VALUE       (-1)
#define DC_ERR_VALUE_TOO_LARGE     (-2)
#define DC_ERR_INVALID_FORMAT      (-3)
#define DC_ERR_INVALID_MODE        (-4)
#define DC_ERR_FILE_IO             (-5)
#define DC_ERR_MEMORY              (-6)
#define DC_ERR_LIBRARY_INIT       (-10)   /* CUDA context / allocation failures */
#define DC_ERR_LIBRARY_CALL       (-11)   /* kernel launch / copy failures, short reads */

#define DC_ABI_VERSION 1

typedef struct dc_ctx dc_ctx;   /* opaque: device, stream, scratch arena, division table */

/* Per-encoder options, the subset of options_t (DCLib/inc/enc_dec.h:28-41) the hot path reads. */
typedef struct dc_params
{
  uint32_t value_size_bits;       /* options_t.value_size_bits, 1..64, default 32 (enc_dec.c:187-197) */
  float    normalization_factor;  /* options_t.normalization_factor, default 100 */
  int32_t  adaptive;              /* options_t.adaptive (bac only), default 0 */
  uint32_t block_size_bits;       /* options_t.block_size_bits (copy only), default 8 (enc_dec.c:187-197); 0 = that default */
} dc_params;

/* Where the streams live inside one flat sample array (channel-major, one channel after another,
   like one raw float32 file per channel).
     uniform mode (sample_off == NULL): n_channels channels of channel_len samples, each cut into
       ceil(channel_len / block_len) blocks; stream s = channel s / nb, block s % nb;
       block_len == 0 means "whole channel".
     ragged mode: stream s covers samples [sample_off[s], sample_off[s+1]); n_streams + 1 entries.
   sample_off is a device pointer for the *_device calls and a host pointer for the *_host calls. */
typedef struct dc_layout
{
  uint64_t n_channels;
  uint64_t channel_len;
  uint64_t block_len;
  uint64_t n_streams;             /* ragged mode only; ignored in uniform mode */
  const uint64_t *sample_off;
} dc_layout;

/* ---- context ------------------------------------------------------------------------------- */

int  dc_abi_version(void);
/* Creates a context on CUDA device `device` (own non-blocking stream).  DC_ERR_LIBRARY_INIT when no
   usable GPU exists: there is no CPU fallback. */
int  dc_create(dc_ctx **ctx, int device);
void dc_destroy(dc_ctx *ctx);
const char *dc_last_error(const dc_ctx *ctx);
uint64_t dc_layout_streams(const dc_layout *layout);
/* Blocks until everything queued on the context's stream has finished. */
int  dc_synchronize(dc_ctx *ctx);
/* cudaStream_t of the context (as void*), for callers that time with CUDA events. */
void *dc_stream(dc_ctx *ctx);
/* Number of kernel launches issued through this context since creation. */
uint64_t dc_launch_count(const dc_ctx *ctx);
/* How many streams the arithmetic-coder kernels keep in flight at once on this device (SM count x resident threads
   per SM; one stream per thread).  A batch whose stream count is a multiple of this runs in whole waves; callers that
   are free to pick a block length (the reference has no block parameter at all, DCLib/src/enc_dec.c:51-60) can use it. */
uint64_t dc_coder_wave_streams(dc_ctx *ctx);

/* Per-kernel device timing (CUDA events on the context's stream around every launch).  The reference's
   analogue is the per-stage "Time elapsed" print of DCCLI (DCCLI/src/cli.c:338-352,444-459). */
enum dc_kernel_id
{
  DC_K_FRONT = 0,     /* normalize + diff + seg (or normalize only) */
  DC_K_BAC_ENC,       /* arithmetic coder, encode */
  DC_K_SCAN,          /* prefix sum over byte counts (3 launches) */
  DC_K_GATHER,        /* compaction into the dense payload */
  DC_K_BAC_DEC,       /* arithmetic coder, decode */
  DC_K_BACK,          /* seg parse + undiff + denormalize */
  DC_K_LZMH_ENC,
  DC_K_LZMH_DEC,
  DC_K_COUNT
};
int dc_set_profiling(dc_ctx *ctx, int enable);
/* ms[DC_K_COUNT]: device time of each kernel group during the LAST call (0 if it did not run).  Synchronises. */
int dc_get_kernel_ms(dc_ctx *ctx, float *ms);

/* ---- batch face: DEGA = normalize -> diff -> seg -> bac ----------------------------------- */
/* Replaces, for every stream at once, the DCCLI chain
     encode normalize # encode diff # encode seg # encode bac [adaptive]
   i.e. Normalize (DCLib/src/normalize.c:9-27), EncodeDifferential (diff.c:9-23),
   EncodeSEG (seg.c:31-43), EncodeBAC (bac.c:147-166).
   Device pointers.  Outputs:
     d_payload  : stream payloads, back to back, each ceil(bit_len/8) bytes (byte-identical to the
                  file DCCLI writes for that stream); capacity payload_cap bytes
     d_bit_len  : [n_streams]   exact compressed bit count per stream
     d_byte_off : [n_streams+1] exclusive prefix sum of the byte counts
     d_status   : [n_streams]   0 or the code the reference chain would have failed with
                  (range checks normalize.c:21-22, diff.c:17-18); a failed stream has bit_len 0
     total_bytes: host, = d_byte_off[n_streams] (the call synchronises to return it)
   Returns DC_ERR_MEMORY if payload_cap is too small (total_bytes is still valid). */
int dc_dega_encode_device(dc_ctx *ctx, const float *d_samples, const dc_layout *layout, const dc_params *params,
                          uint8_t *d_payload, uint64_t payload_cap,
                          uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status,
                          uint64_t *total_bytes);

/* Inverse chain  decode bac [adaptive] # decode seg # decode diff # decode normalize
   (DecodeBAC bac.c:244-264, DecodeSEG seg.c:82-95, DecodeDifferential diff.c:25-38,
   Denormalize normalize.c:29-42) for every stream.  `layout` gives where each stream's samples go
   and how many are expected (capacity); d_n_out[s] receives the number actually decoded. */
int dc_dega_decode_device(dc_ctx *ctx, const uint8_t *d_payload, uint64_t payload_bytes,
                          const uint64_t *d_bit_len, const uint64_t *d_byte_off,
                          const dc_layout *layout, const dc_params *params,
                          float *d_samples, uint64_t *d_n_out, int32_t *d_status);

/* Same two calls with HOST buffers (pinned or pageable); host<->device copies happen inside. */
int dc_dega_encode_host(dc_ctx *ctx, const float *samples, const dc_layout *layout, const dc_params *params,
                        uint8_t *payload, uint64_t payload_cap,
                        uint64_t *bit_len, uint64_t *byte_off, int32_t *status, uint64_t *total_bytes);
int dc_dega_decode_host(dc_ctx *ctx, const uint8_t *payload, uint64_t payload_bytes,
                        const uint64_t *bit_len, const uint64_t *byte_off,
                        const dc_layout *layout, const dc_params *params,
                        float *samples, uint64_t *n_out, int32_t *status);

/* Asynchronous forms of the two calls above: same arguments and results, but the call only queues the job and
   returns; dc_wait() blocks until every job queued on the context has finished and returns the first non-zero code
   among them, in submission order.  A context runs at most one encode and one decode at a time (further jobs wait
   their turn, in order); every buffer and *total_bytes must stay valid until dc_wait() returns (layout and params are
   copied), and no other call may be made on the context meanwhile.  The reason to use them: an encode is mostly host
   -> device traffic and a decode mostly device -> host, so an encode queued next to a decode keeps both directions of
   the PCIe link busy, where one call after the other uses one direction at a time.  The library issues the copies of
   the two jobs itself, interleaved on one stream per direction (csrc/api.cu, "the job engine"); two contexts with
   their own copy streams do NOT overlap on this hardware.  The reference has no analogue (one DCCLI process per
   stream, DCCLI/src/cli.c:407-473); its unit of overlap is the process. */
int dc_dega_encode_host_async(dc_ctx *ctx, const float *samples, const dc_layout *layout, const dc_params *params,
                              uint8_t *payload, uint64_t payload_cap,
                              uint64_t *bit_len, uint64_t *byte_off, int32_t *status, uint64_t *total_bytes);
int dc_dega_decode_host_async(dc_ctx *ctx, const uint8_t *payload, uint64_t payload_bytes,
                              const uint64_t *bit_len, const uint64_t *byte_off,
                              const dc_layout *layout, const dc_params *params,
                              float *samples, uint64_t *n_out, int32_t *status);
int dc_wait(dc_ctx *ctx);
/* Waits for the OLDEST asynchronous job not yet collected and returns its code (DC_OK when there is none).  For
   streaming use: keep queueing jobs and collect them in submission order, so that the copies and kernels of the next
   job start while the results of this one are being consumed. */
int dc_wait_next(dc_ctx *ctx);

/* ---- batch face: A-XDR = normalize only ---------------------------------------------------- */
/* `encode normalize` (normalize.c:9-27) for every stream: stream s becomes len(s) * value_size_bits
   bits, MSB-first, zero padded to a byte; records are laid back to back (d_byte_off as above). */
int dc_axdr_encode_device(dc_ctx *ctx, const float *d_samples, const dc_layout *layout, const dc_params *params,
                          uint8_t *d_payload, uint64_t payload_cap,
                          uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status,
                          uint64_t *total_bytes);
/* `decode normalize` (normalize.c:29-42). */
int dc_axdr_decode_device(dc_ctx *ctx, const uint8_t *d_payload, uint64_t payload_bytes,
                          const uint64_t *d_bit_len, const uint64_t *d_byte_off,
                          const dc_layout *layout, const dc_params *params,
                          float *d_samples, uint64_t *d_n_out, int32_t *d_status);

/* ---- batch face: LZMH ---------------------------------------------------------------------- */
/* `encode lzmh` / `decode lzmh` (EncodeLZMH lzmh.c:130-370, DecodeLZMH lzmh.c:383-574) over
   n_streams independent byte blocks.  d_in_off[n_streams+1] are byte offsets into d_in.
   Decode: d_out_off[n_streams+1] gives each block's output capacity window; d_out_len[s] = bytes
   produced. */
int dc_lzmh_encode_device(dc_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t n_streams,
                          uint8_t *d_payload, uint64_t payload_cap,
                          uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status,
                          uint64_t *total_bytes);
int dc_lzmh_decode_device(dc_ctx *ctx, const uint8_t *d_payload, uint64_t payload_bytes,
                          const uint64_t *d_bit_len, const uint64_t *d_byte_off, uint64_t n_streams,
                          uint8_t *d_out, const uint64_t *d_out_off, uint64_t *d_out_len, int32_t *d_status);

/* ---- batch face: aggregate (granularity reduction) ------------------------------------------ */

/* DCLib/src/aggregate.c:9-27 (`encode aggregate num_values=k`) for n_channels channels of channel_len
   float32 samples each (device pointers, channel-major): out[c][g] = in[c][g*k] + in[c][g*k+1] + ...
   added strictly left to right starting from 0.0f, like the reference's loop; the last group of a
   channel may be short (aggregate.c:21-22).  d_out holds n_channels * ceil(channel_len / k) floats.
   Used to derive 1 min / 15 min series from 1 s data on the device (BASELINE.json configs[4]). */
int dc_aggregate_device(dc_ctx *ctx, const float *d_samples, uint64_t n_channels, uint64_t channel_len,
                        uint32_t num_values, float *d_out);

/* ---- csv: text column -> float32 ------------------------------------------------------------ */

/* `decode csv [column=..] [separator=..]` (DCLib/src/csv.c:13-44, ReadCSV): the `column`-th field (counted
   from 1) of every line of `text` through strtof, in line order; lines without that field produce nothing,
   an unterminated last line counts.  Host buffers; line splitting, field search and the decimal -> float32
   conversion run on the GPU (correctly rounded, i.e. strtof's result, for decimals of up to 7 significant digits,
   with or without a small decimal exponent, and integers of up to 19 digits; what is left of strtof's grammar --
   hex floats, inf/nan, longer digit strings, large exponents -- is converted by strtof on the host).  On DC_ERR_MEMORY *n_out holds the count needed. */
int dc_csv_decode_host(dc_ctx *ctx, const uint8_t *text, uint64_t nbytes, uint32_t column, char separator,
                       float *out, uint64_t out_cap, uint64_t *n_out);

/* `encode csv column=c separator_char=s num_decimal_places=d` (DCLib/src/csv.c:46-65, WriteCSV) for n float32 values:
   per value c - 1 separator characters, then printf's "%.*f" with d <= 6 digits, then '\n'; the text is what the
   reference's sprintf writes (exact value, round half to even, "-0.00", "inf", "-nan").  *n_out = bytes of text;
   on DC_ERR_MEMORY it holds the size needed. */
int dc_csv_encode_host(dc_ctx *ctx, const float *values, uint64_t n, uint32_t column, char separator, uint32_t num_decimal_places,
                       uint8_t *text, uint64_t text_cap, uint64_t *n_out);

/* ---- container: one file for the streams of a batch encode (host code, no GPU needed) ------- */

/* The reference writes one stream per file (DCIOLib/src/bit_file_buffer.c:310-333 via DCCLI/src/cli.c:129-148).
   The container keeps every stream's bytes exactly as DCCLI would have written them, back to back, behind a
   header and a bit-length index; see data_compressor_b200/csrc/container.cpp for the layout. */
#define DC_CONTAINER_VERSION 1
#define DC_CHAIN_DEGA 1   /* normalize # diff # seg # bac */
#define DC_CHAIN_AXDR 2   /* normalize */
#define DC_CHAIN_LZMH 3   /* lzmh */

typedef struct dc_container_header
{
  char     magic[8];              /* "DCB200\0\1" */
  uint32_t version;
  uint32_t chain;                 /* DC_CHAIN_* */
  uint32_t value_size_bits;       /* the stage options the streams were coded with (enc_dec.h:28-41) */
  float    normalization_factor;
  int32_t  adaptive;
  uint32_t reserved0;
  uint64_t n_channels;            /* dc_layout of the samples (uniform mode); 0 for LZMH byte blocks */
  uint64_t channel_len;
  uint64_t block_len;
  uint64_t n_streams;
  uint64_t payload_bytes;
  uint64_t index_offset;          /* file offset of bit_len[n_streams], uint64 little endian */
  uint64_t payload_offset;        /* file offset of the first stream */
  uint64_t reserved[5];
} dc_container_header;            /* 128 bytes */

/* Fills a header (offsets included) for n_streams streams and payload_bytes bytes of payload. */
int dc_container_init_header(dc_container_header *h, uint32_t chain, const dc_params *params, const dc_layout *layout,
                             uint64_t n_streams, uint64_t payload_bytes);
/* Creates / truncates the file, writes the header and sizes the file for index and payload. */
int dc_container_create(const char *path, const dc_container_header *h);
/* Writes the index entries [first_stream, first_stream + n) and payload_bytes bytes of payload at byte_base
   (offsets relative to the container's payload).  Position based: the shards of a multi-GPU job
   write their own slices, in any order, after one of them has created the file. */
int dc_container_write_shard(const char *path, const dc_container_header *h, uint64_t first_stream, uint64_t n,
                             const uint64_t *bit_len, uint64_t byte_base, const uint8_t *payload, uint64_t payload_bytes);
int dc_container_read_header(const char *path, dc_container_header *h);
int dc_container_read_index(const char *path, const dc_container_header *h, uint64_t first_stream, uint64_t n, uint64_t *bit_len);
int dc_container_read_payload(const char *path, const dc_container_header *h, uint64_t byte_off, uint64_t nbytes, uint8_t *dst);

/* ---- plugin face: one stage, one stream, host buffers -------------------------------------- */
/* What an enc_dec_function_t body (DCLib/inc/enc_dec.h:11) calls after draining its input bit
   buffer: runs ONE reference stage on the GPU over `in_bits` bits and returns the exact output bit
   count.  Stage ids mirror the registry names of DCLib/src/enc_dec.c:51-60. */
enum dc_stage
{
  DC_STAGE_NORMALIZE = 0,   /* encode normalize  : Normalize            normalize.c:9-27  */
  DC_STAGE_DENORMALIZE,     /* decode normalize  : Denormalize          normalize.c:29-42 */
  DC_STAGE_DIFF_ENC,        /* encode diff       : EncodeDifferential   diff.c:9-23       */
  DC_STAGE_DIFF_DEC,        /* decode diff       : DecodeDifferential   diff.c:25-38      */
  DC_STAGE_SEG_ENC,         /* encode seg        : EncodeSEG            seg.c:31-43       */
  DC_STAGE_SEG_DEC,         /* decode seg        : DecodeSEG            seg.c:82-95       */
  DC_STAGE_BAC_ENC,         /* encode bac        : EncodeBAC            bac.c:147-166     */
  DC_STAGE_BAC_DEC,         /* decode bac        : DecodeBAC            bac.c:244-264     */
  DC_STAGE_LZMH_ENC,        /* encode lzmh       : EncodeLZMH           lzmh.c:130-370    */
  DC_STAGE_LZMH_DEC,        /* decode lzmh       : DecodeLZMH           lzmh.c:383-574    */
  DC_STAGE_DEGA_ENC,        /* fused normalize # diff # seg # bac                           */
  DC_STAGE_DEGA_DEC,        /* fused bac # seg # diff # normalize (decode)                  */
  DC_STAGE_COPY,            /* encode / decode copy [blocksize=..] : Copy    copy.c:11-29   */
  DC_STAGE_COUNT
};

/* out must hold out_cap bytes; on DC_ERR_MEMORY *out_bits holds the size that would have been
   needed.  On a codec error (negative return) *out_bits is 0. */
int dc_stage_run_host(dc_ctx *ctx, int stage, const uint8_t *in, uint64_t in_bits, const dc_params *params,
                      uint8_t *out, uint64_t out_cap, uint64_t *out_bits);

/* A chain of stages over one stream, `encode X # encode Y # ...` of the reference's command line (DCCLI/src/cli.c:430-466
   runs the stages one after the other, each reading what the one before wrote).  The intermediate results stay in device
   memory: one copy in, one copy out.  stages[i] is a DC_STAGE_* value, params[i] its options (the reference parses
   options per stage, DCCLI/src/params.c:248); at most 16 stages (DCCLI/inc/params.h:10).  Returns the code of the first
   stage that fails -- *failed_stage (may be NULL) then holds its index -- with the reference's meaning.  On DC_ERR_MEMORY
   from the last stage *out_bits holds the size that would have been needed. */
int dc_chain_run_host(dc_ctx *ctx, uint32_t n_stages, const int *stages, const dc_params *params, const uint8_t *in,
                      uint64_t in_bits, uint8_t *out, uint64_t out_cap, uint64_t *out_bits, uint32_t *failed_stage);
/* Upper bound, in bytes, of the output of `stage` for an input of in_bits bits. */
uint64_t dc_stage_out_bound(int stage, uint64_t in_bits, const dc_params *params);

#ifdef __cplusplus
}
#endif

#endif
