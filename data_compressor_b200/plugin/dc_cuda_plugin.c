/* dc_cuda_plugin.c -- DCLib-side glue: enc_dec_function_t bodies that run on the GPU.
 *
 * This is the file a Data Compressor maintainer adds to DCLib.  It is C90 like the rest of
 * the reference, includes only the reference's public headers (DCLib/inc/enc_dec.h, io_macros.h) and the C ABI of
 * libdc_cuda.so (include/dc_cuda.h).  Every function has the plugin signature of DCLib/inc/enc_dec.h:11:
 *
 *   io_int_t f(bit_file_buffer_t *in, bit_file_buffer_t *out, const options_t *options)
 *
 * and follows the contract of SURVEY.md section 8b: read until EndOfBitFileBuffer(in), write the exact number of
 * result bits, never flush/close the buffers, return NO_ERROR or a negative code of common/inc/err_codes.h.
 *
 * Two ways to register them:
 *   (a) new names next to the CPU codecs: "dega", "gpu_bac", ... (two-line edit of the table in enc_dec.c), or
 *   (b) -DDC_CUDA_REPLACE_REFERENCE: this file then defines Normalize, EncodeDifferential, EncodeSEG, EncodeBAC,
 *       EncodeLZMH and their decoders themselves, so linking it INSTEAD of normalize.c/diff.c/seg.c/bac.c/lzmh.c
 *       gives a DCCLI whose unchanged command lines run on the GPU.  oracle/Makefile builds that binary as
 *       oracle/_ref/DCCLI_cuda for the drop-in test.
 */
#include "err_codes.h"
#include "io_macros.h"
#include "enc_dec.h"

#include "dc_cuda.h"

#include <stdlib.h>
#include <string.h>

static dc_ctx *g_ctx = NULL;

static io_int_t GetContext(dc_ctx **ctx, FILE * const error_log_file)
{
  if (g_ctx == NULL)
  {
    const char * const dev = getenv("DC_CUDA_DEVICE");
    const int ret = dc_create(&g_ctx, dev != NULL ? atoi(dev) : 0);
    if (ret != DC_OK)
    {
      LOG_ERROR(error_log_file, "Could not create a CUDA context (no usable GPU?); there is no CPU fallback\n");
      g_ctx = NULL;
      return ERROR_LIBRARY_INIT;
    }
  }
  *ctx = g_ctx;
  return NO_ERROR;
}

/* Drains the whole input bit buffer; *bits receives the exact bit count (the in-chain hand-off keeps fractional
   bytes, DCIOLib/src/bit_file_buffer.c:135-141,169-176). */
static io_int_t DrainInput(bit_file_buffer_t * const in_bit_buf, uint8_t ** const data, io_uint_t * const bits)
{
  size_t capacity = 1 << 16, used = 0;
  uint8_t *buffer = (uint8_t*)malloc(capacity);
  *bits = 0;
  *data = NULL;
  if (buffer == NULL)
    return ERROR_MEMORY;
  while (!EndOfBitFileBuffer(in_bit_buf))
  {
    const size_t chunk = 1 << 15; /* bytes */
    io_int_t got;
    if (used + chunk > capacity)
    {
      uint8_t *bigger;
      capacity *= 2;
      if ((bigger = (uint8_t*)realloc(buffer, capacity)) == NULL)
      {
        free(buffer);
        return ERROR_MEMORY;
      }
      buffer = bigger;
    }
    got = ReadBitFileBuffer(in_bit_buf, buffer + used, 8 * chunk);
    if (got < 0)
    {
      free(buffer);
      return got;
    }
    *bits += (io_uint_t)got;
    used += (size_t)((got + 7) / 8);
    if ((size_t)got != 8 * chunk) /* Short read: end of input (possibly inside a byte) */
      break;
  }
  *data = buffer;
  return NO_ERROR;
}

static io_int_t RunStage(const int stage, bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options)
{
  dc_ctx *ctx;
  dc_params params;
  uint8_t *in_data, *out_data;
  io_uint_t in_bits;
  uint64_t out_bits = 0, out_capacity;
  io_int_t ret;
  int attempt;
  if ((ret = GetContext(&ctx, options->error_log_file)) != NO_ERROR)
    return ret;
  if ((ret = DrainInput(in_bit_buf, &in_data, &in_bits)) != NO_ERROR)
    return ret;
  params.value_size_bits = (uint32_t)options->value_size_bits;
  params.normalization_factor = options->normalization_factor;
  params.adaptive = options->adaptive;
  params.block_size_bits = (uint32_t)options->block_size_bits;
  out_capacity = dc_stage_out_bound(stage, in_bits, &params);
  out_data = NULL;
  for (attempt = 0; attempt < 2; attempt++)
  {
    if ((out_data = (uint8_t*)malloc((size_t)out_capacity + 16)) == NULL)
    {
      free(in_data);
      return ERROR_MEMORY;
    }
    ret = dc_stage_run_host(ctx, stage, in_data, in_bits, &params, out_data, out_capacity, &out_bits);
    if (ret != DC_ERR_MEMORY || out_bits == 0)
      break;
    free(out_data); /* The library reported the exact size it needs: retry once */
    out_data = NULL;
    out_capacity = (out_bits + 7) / 8 + 16;
  }
  free(in_data);
  if (ret != NO_ERROR)
  {
    if (out_data != NULL)
      free(out_data);
    if (ret == ERROR_LIBRARY_CALL || ret == ERROR_LIBRARY_INIT)
      LOG_ERROR(options->error_log_file, "CUDA stage %d failed: %s\n", stage, dc_last_error(ctx));
    return ret;
  }
  if (out_bits > 0)
  {
    WRITE_BITS_CHECKED_WITH_ACTION_ON_ERROR(out_data, (size_t)out_bits, out_bit_buf, options->error_log_file, free(out_data));
  }
  free(out_data);
  return NO_ERROR;
}

/* ---- fused chains: one plugin = the whole DEGA chain ------------------------------------------------------ */
/* `encode dega [adaptive] [valuesize=..] [normalization_factor=..]` == encode normalize # encode diff # encode seg # encode bac */
io_int_t EncodeDEGA_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options)
{
  return RunStage(DC_STAGE_DEGA_ENC, in_bit_buf, out_bit_buf, options);
}

/* `decode dega ...` == decode bac # decode seg # decode diff # decode normalize */
io_int_t DecodeDEGA_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options)
{
  return RunStage(DC_STAGE_DEGA_DEC, in_bit_buf, out_bit_buf, options);
}

/* ---- stage-by-stage plugins --------------------------------------------------------------------------------- */
#ifdef DC_CUDA_REPLACE_REFERENCE
  #define PLUGIN_NAME(cpu_name, gpu_name) cpu_name
#else
  #define PLUGIN_NAME(cpu_name, gpu_name) gpu_name
#endif

#define DEFINE_STAGE_PLUGIN(cpu_name, gpu_name, stage) \
  io_int_t PLUGIN_NAME(cpu_name, gpu_name)(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options) \
  { \
    return RunStage((stage), in_bit_buf, out_bit_buf, options); \
  }

DEFINE_STAGE_PLUGIN(Normalize, Normalize_CUDA, DC_STAGE_NORMALIZE)                             /* normalize.c:9  */
DEFINE_STAGE_PLUGIN(Denormalize, Denormalize_CUDA, DC_STAGE_DENORMALIZE)                       /* normalize.c:29 */
DEFINE_STAGE_PLUGIN(EncodeDifferential, EncodeDifferential_CUDA, DC_STAGE_DIFF_ENC)            /* diff.c:9       */
DEFINE_STAGE_PLUGIN(DecodeDifferential, DecodeDifferential_CUDA, DC_STAGE_DIFF_DEC)            /* diff.c:25      */
DEFINE_STAGE_PLUGIN(EncodeSEG, EncodeSEG_CUDA, DC_STAGE_SEG_ENC)                               /* seg.c:31       */
DEFINE_STAGE_PLUGIN(DecodeSEG, DecodeSEG_CUDA, DC_STAGE_SEG_DEC)                               /* seg.c:82       */
DEFINE_STAGE_PLUGIN(EncodeBAC, EncodeBAC_CUDA, DC_STAGE_BAC_ENC)                               /* bac.c:147      */
DEFINE_STAGE_PLUGIN(DecodeBAC, DecodeBAC_CUDA, DC_STAGE_BAC_DEC)                               /* bac.c:244      */
DEFINE_STAGE_PLUGIN(EncodeLZMH, EncodeLZMH_CUDA, DC_STAGE_LZMH_ENC)                            /* lzmh.c:130     */
DEFINE_STAGE_PLUGIN(DecodeLZMH, DecodeLZMH_CUDA, DC_STAGE_LZMH_DEC)                            /* lzmh.c:383     */
DEFINE_STAGE_PLUGIN(Copy, Copy_CUDA, DC_STAGE_COPY)                                            /* copy.c:11      */

/* ---- chains: several stages of one command line as ONE device-resident chain (dc_chain_run_host) ----------------- */
/* Which DC_STAGE_* value a registry function of this file runs; -1 for any other function (the csv plugins below, the
   reference's own codecs): those are stages a chain fuser leaves alone. */
int DcCudaStageOf(enc_dec_function_t * const function, const int encode)
{
  static const struct { enc_dec_function_t *function; int stage; } stages[] =
  {
    { PLUGIN_NAME(Normalize, Normalize_CUDA), DC_STAGE_NORMALIZE }, { PLUGIN_NAME(Denormalize, Denormalize_CUDA), DC_STAGE_DENORMALIZE },
    { PLUGIN_NAME(EncodeDifferential, EncodeDifferential_CUDA), DC_STAGE_DIFF_ENC }, { PLUGIN_NAME(DecodeDifferential, DecodeDifferential_CUDA), DC_STAGE_DIFF_DEC },
    { PLUGIN_NAME(EncodeSEG, EncodeSEG_CUDA), DC_STAGE_SEG_ENC }, { PLUGIN_NAME(DecodeSEG, DecodeSEG_CUDA), DC_STAGE_SEG_DEC },
    { PLUGIN_NAME(EncodeBAC, EncodeBAC_CUDA), DC_STAGE_BAC_ENC }, { PLUGIN_NAME(DecodeBAC, DecodeBAC_CUDA), DC_STAGE_BAC_DEC },
    { PLUGIN_NAME(EncodeLZMH, EncodeLZMH_CUDA), DC_STAGE_LZMH_ENC }, { PLUGIN_NAME(DecodeLZMH, DecodeLZMH_CUDA), DC_STAGE_LZMH_DEC },
    { PLUGIN_NAME(Copy, Copy_CUDA), DC_STAGE_COPY }, { EncodeDEGA_CUDA, DC_STAGE_DEGA_ENC }, { DecodeDEGA_CUDA, DC_STAGE_DEGA_DEC }
  };
  size_t i;
  (void)encode;
  for (i = 0; i < sizeof(stages) / sizeof(stages[0]); i++)
    if (stages[i].function == function)
      return stages[i].stage;
  return -1;
}

/* Runs `count` consecutive stages of a parsed command line (their options_t in order, DCCLI/inc/params.h:12-17) with one
   copy to the device and one copy back: drains `in_bit_buf`, writes the exact result bits to `out_bit_buf`.  Every stage
   must be one DcCudaStageOf() knows.  On failure *failed_stage is the index of the stage that failed, with the code the
   reference's stage would have returned. */
io_int_t DcCudaRunChain(options_t * const * const options, const size_t count, bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf,
                        size_t * const failed_stage)
{
  dc_ctx *ctx;
  dc_params params[16]; /* MAX_OPTIONS, DCCLI/inc/params.h:10 */
  int stages[16];
  uint8_t *in_data, *out_data = NULL;
  io_uint_t in_bits;
  uint64_t out_bits = 0, out_capacity;
  uint32_t failed = 0;
  io_int_t ret;
  size_t i;
  int attempt;
  FILE * const error_log_file = options[0]->error_log_file;
  *failed_stage = 0;
  if (count == 0 || count > 16)
    return ERROR_INVALID_VALUE;
  for (i = 0; i < count; i++)
  {
    enc_dec_function_t * const function = options[i]->encode ? options[i]->encoder_decoder->encoder : options[i]->encoder_decoder->decoder;
    if ((stages[i] = DcCudaStageOf(function, options[i]->encode)) < 0)
    {
      *failed_stage = i;
      return ERROR_INVALID_MODE;
    }
    params[i].value_size_bits = (uint32_t)options[i]->value_size_bits;
    params[i].normalization_factor = options[i]->normalization_factor;
    params[i].adaptive = options[i]->adaptive;
    params[i].block_size_bits = (uint32_t)options[i]->block_size_bits;
  }
  if ((ret = GetContext(&ctx, error_log_file)) != NO_ERROR)
    return ret;
  if ((ret = DrainInput(in_bit_buf, &in_data, &in_bits)) != NO_ERROR)
    return ret;
  out_capacity = (in_bits + 7) / 8 * 8 + (1 << 16);
  for (attempt = 0; attempt < 2; attempt++)
  {
    if ((out_data = (uint8_t*)malloc((size_t)out_capacity + 16)) == NULL)
    {
      free(in_data);
      return ERROR_MEMORY;
    }
    ret = dc_chain_run_host(ctx, (uint32_t)count, stages, params, in_data, in_bits, out_data, out_capacity, &out_bits, &failed);
    if (ret != DC_ERR_MEMORY || failed + 1 != count || out_bits == 0)
      break;
    free(out_data); /* The last stage reported the exact size it needs: once more with that */
    out_data = NULL;
    out_capacity = (out_bits + 7) / 8 + 16;
  }
  free(in_data);
  if (ret != NO_ERROR)
  {
    if (out_data != NULL)
      free(out_data);
    *failed_stage = failed;
    if (ret == ERROR_LIBRARY_CALL || ret == ERROR_LIBRARY_INIT)
      LOG_ERROR(error_log_file, "CUDA chain failed at its stage %lu: %s\n", (unsigned long)(failed + 1), dc_last_error(ctx));
    return ret;
  }
  if (out_bits > 0)
  {
    WRITE_BITS_CHECKED_WITH_ACTION_ON_ERROR(out_data, (size_t)out_bits, out_bit_buf, error_log_file, free(out_data));
  }
  free(out_data);
  return NO_ERROR;
}

/* ---- text input ------------------------------------------------------------------------------------------------ */
/* `decode csv [column=..] [separator_char=..]` (csv.c:13).  csv.c defines ReadCSV and WriteCSV in one file, so this one
   always carries its own name; a build that wants it behind the unchanged "csv" registry row compiles enc_dec.c with
   -DReadCSV=ReadCSV_CUDA (oracle/Makefile does that for oracle/_ref/DCCLI_cuda). */
io_int_t ReadCSV_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options)
{
  dc_ctx *ctx;
  uint8_t *in_data;
  float *values;
  io_uint_t in_bits, i;
  uint64_t capacity = 1, count = 0;
  io_int_t ret;
  if ((ret = GetContext(&ctx, options->error_log_file)) != NO_ERROR)
    return ret;
  if ((ret = DrainInput(in_bit_buf, &in_data, &in_bits)) != NO_ERROR)
    return ret;
  for (i = 0; i < in_bits / 8; i++) /* At most one value per line */
    capacity += in_data[i] == '\n';
  if ((values = (float*)malloc((size_t)capacity * sizeof(float))) == NULL)
  {
    free(in_data);
    return ERROR_MEMORY;
  }
  ret = dc_csv_decode_host(ctx, in_data, in_bits / 8, (uint32_t)options->column, options->separator_char, values, capacity, &count);
  free(in_data);
  if (ret != NO_ERROR)
  {
    free(values);
    if (ret == ERROR_LIBRARY_CALL || ret == ERROR_LIBRARY_INIT)
      LOG_ERROR(options->error_log_file, "CUDA csv reader failed: %s\n", dc_last_error(ctx));
    return ret;
  }
  if (count > 0)
  {
    WRITE_BITS_CHECKED_WITH_ACTION_ON_ERROR((const uint8_t*)values, (size_t)(32 * count), out_bit_buf, options->error_log_file, free(values));
  }
  free(values);
  return NO_ERROR;
}

/* `encode csv [column=..] [separator_char=..] [num_decimal_places=..]` (csv.c:46); same naming rule as ReadCSV_CUDA. */
io_int_t WriteCSV_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options)
{
  dc_ctx *ctx;
  uint8_t *in_data, *text;
  io_uint_t in_bits;
  uint64_t capacity, count = 0, n;
  io_int_t ret;
  int attempt;
  if ((ret = GetContext(&ctx, options->error_log_file)) != NO_ERROR)
    return ret;
  if ((ret = DrainInput(in_bit_buf, &in_data, &in_bits)) != NO_ERROR)
    return ret;
  n = in_bits / 32;
  if (in_bits % 32 != 0) /* The reference fails on a partial value: READ_BITS_CHECKED in csv.c:54 */
  {
    free(in_data);
    LOG_ERROR(options->error_log_file, "Could not read %lu bits from input\n", (unsigned long)32);
    return ERROR_LIBRARY_CALL;
  }
  capacity = n * (options->column + 15) + 64; /* Values below 10^7; the library reports the size when it needs more */
  text = NULL;
  for (attempt = 0; attempt < 2; attempt++)
  {
    if ((text = (uint8_t*)malloc((size_t)capacity)) == NULL)
    {
      free(in_data);
      return ERROR_MEMORY;
    }
    ret = dc_csv_encode_host(ctx, (const float*)in_data, n, (uint32_t)options->column, options->separator_char, (uint32_t)options->num_decimal_places,
                             text, capacity, &count);
    if (ret != DC_ERR_MEMORY || count <= capacity)
      break;
    free(text);
    text = NULL;
    capacity = count;
  }
  free(in_data);
  if (ret != NO_ERROR)
  {
    if (text != NULL)
      free(text);
    if (ret == ERROR_LIBRARY_CALL || ret == ERROR_LIBRARY_INIT)
      LOG_ERROR(options->error_log_file, "CUDA csv writer failed: %s\n", dc_last_error(ctx));
    return ret;
  }
  if (count > 0)
  {
    WRITE_BITS_CHECKED_WITH_ACTION_ON_ERROR(text, (size_t)(8 * count), out_bit_buf, options->error_log_file, free(text));
  }
  free(text);
  return NO_ERROR;
}
