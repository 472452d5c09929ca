/* dc_fused_main.c -- DCCLI's front door with a chain fuser in front of the stage loop.
 *
 * The reference's main() (DCCLI/src/cli.c:407-473) hands one stage at a time to the registry's functions and moves
 * every intermediate result through host memory (cli.c:430-466).  With GPU plugins behind the registry that is one copy
 * to the device and one copy back PER STAGE.  This file is what a maintainer puts in front of that loop: it parses the
 * command line with the reference's own ProcessParameters() (DCCLI/inc/params.h:19), groups every run of consecutive
 * stages that are GPU plugins (DcCudaStageOf) into ONE DcCudaRunChain() call -- one copy in, the stages back to back on
 * the device, one copy out -- and runs every other stage (csv, aggregate, anything the registry holds that is not a GPU
 * plugin) through its registry function exactly as cli.c does, with the same in-memory hand-off between the pieces
 * (SetBitFileBufferMode keeps the fractional last byte, DCIOLib/src/bit_file_buffer.c:135-141).
 *
 * oracle/Makefile links it into oracle/_ref/DCCLI_cuda together with the UNMODIFIED cli.c, whose main() is renamed on
 * the compiler's command line (-Dmain=DCCLIReferenceMain): a command line with nothing to fuse, or DC_CUDA_NO_FUSE in the
 * environment, goes to the reference's own main().  C90, reference headers and DCIOLib's public API only; no line of
 * cli.c is copied.  The in-tree equivalent is a call to DcCudaRunChain() from cli.c's stage loop.
 */
#include "err_codes.h"
#include "io_macros.h"
#include "enc_dec.h"
#include "params.h"

#include "dc_cuda_plugin.h"

#include <stdlib.h>

int DCCLIReferenceMain(const int argc, const char * const * const argv);

typedef struct mem_buffer_t
{
  file_buffer_t *file_buf;
  bit_file_buffer_t *bit_buf;
} mem_buffer_t;

static int OpenBuffer(mem_buffer_t * const buffer, FILE * const file, const file_buffer_mode_t mode)
{
  buffer->file_buf = AllocateFileBuffer();
  buffer->bit_buf = AllocateBitFileBuffer();
  if (buffer->file_buf == NULL || buffer->bit_buf == NULL)
    return ERROR_MEMORY;
  if ((file != NULL ? InitFileBuffer(buffer->file_buf, file, mode, 1 << 16) : InitFileBufferInMemory(buffer->file_buf, mode, 1 << 16)) != NO_ERROR)
    return ERROR_LIBRARY_INIT;
  InitBitFileBuffer(buffer->bit_buf, buffer->file_buf);
  return NO_ERROR;
}

static void CloseBuffer(mem_buffer_t * const buffer)
{
  if (buffer->bit_buf != NULL)
    UninitBitFileBuffer(buffer->bit_buf); /* Flushes a fractional last byte, zero padded (bit_file_buffer.c:310-333) */
  if (buffer->file_buf != NULL)
  {
    UninitFileBuffer(buffer->file_buf);
    FreeFileBuffer(buffer->file_buf);
  }
  if (buffer->bit_buf != NULL)
    FreeBitFileBuffer(buffer->bit_buf);
  buffer->file_buf = NULL;
  buffer->bit_buf = NULL;
}

static int StageOf(const options_t * const options)
{
  return DcCudaStageOf(options->encode ? options->encoder_decoder->encoder : options->encoder_decoder->decoder, options->encode);
}

int main(const int argc, const char * const * const argv)
{
  parameters_t parameters;
  FILE *in_file = NULL, *out_file = NULL;
  mem_buffer_t in_buffer = { NULL, NULL }, out_buffer = { NULL, NULL }, temp[2] = { { NULL, NULL }, { NULL, NULL } };
  size_t i, first, longest_run = 0, run = 0, pieces = 0;
  int ret, which = 0;
  io_int_t stage_ret = NO_ERROR;
  parameters.error_log_file = stderr;
  if (getenv("DC_CUDA_NO_FUSE") != NULL)
    return DCCLIReferenceMain(argc, argv);
  if ((ret = ProcessParameters(argc - 1, &argv[1], &parameters, &in_file, &out_file)) != NO_ERROR)
    return ret;
  for (i = 0; i < parameters.num_options; i++)
  {
    run = StageOf(parameters.options[i]) >= 0 ? run + 1 : 0;
    if (run > longest_run)
      longest_run = run;
  }
  if (longest_run < 2) /* Nothing to fuse: the reference's own loop does the job (it parses the command line again) */
  {
    if (in_file != NULL && in_file != stdin)
      fclose(in_file);
    if (out_file != NULL && out_file != stdout)
      fclose(out_file);
    return DCCLIReferenceMain(argc, argv);
  }
  if (OpenBuffer(&in_buffer, in_file, FBM_READING) != NO_ERROR || OpenBuffer(&out_buffer, out_file, FBM_WRITING) != NO_ERROR ||
      OpenBuffer(&temp[0], NULL, FBM_WRITING) != NO_ERROR || OpenBuffer(&temp[1], NULL, FBM_WRITING) != NO_ERROR)
  {
    LOG_ERROR(parameters.error_log_file, "Error initializing the I/O buffers\n");
    return ERROR_LIBRARY_INIT;
  }
  for (first = 0; first < parameters.num_options && stage_ret == NO_ERROR; pieces++)
  {
    size_t count = 1, failed = 0;
    bit_file_buffer_t *read_bit_buf, *write_bit_buf;
    if (StageOf(parameters.options[first]) >= 0)
      while (first + count < parameters.num_options && StageOf(parameters.options[first + count]) >= 0)
        count++;
    read_bit_buf = first == 0 ? in_buffer.bit_buf : temp[which].bit_buf;                                   /* cli.c:434 */
    write_bit_buf = first + count == parameters.num_options ? out_buffer.bit_buf : temp[1 - which].bit_buf; /* cli.c:435 */
    if (StageOf(parameters.options[first]) >= 0)
      stage_ret = DcCudaRunChain(&parameters.options[first], count, read_bit_buf, write_bit_buf, &failed);
    else
    {
      options_t * const options = parameters.options[first];
      stage_ret = (*(options->encode ? options->encoder_decoder->encoder : options->encoder_decoder->decoder))(read_bit_buf, write_bit_buf, options);
    }
    if (stage_ret != NO_ERROR)
    {
      LOG_ERROR(parameters.error_log_file, "%s while executing encoder/decoder %lu of %lu\n", ERROR_MESSAGE_STRING(stage_ret),
                (unsigned long)(first + failed + 1), (unsigned long)parameters.num_options);
      break;
    }
    first += count;
    if (first < parameters.num_options) /* What was written becomes the next piece's input, to the bit (cli.c:212-223) */
    {
      which = 1 - which;
      if (SetBitFileBufferMode(temp[which].bit_buf, FBM_READING) != NO_ERROR || ResetBitFileBuffer(temp[1 - which].bit_buf, FBM_WRITING) != NO_ERROR)
      {
        LOG_ERROR(parameters.error_log_file, "Error switching the temporary buffers\n");
        stage_ret = ERROR_LIBRARY_CALL;
      }
    }
  }
  CloseBuffer(&temp[0]);
  CloseBuffer(&temp[1]);
  CloseBuffer(&out_buffer);
  CloseBuffer(&in_buffer);
  if (getenv("DC_CUDA_FUSE_REPORT") != NULL)
    fprintf(stderr, "dc_fused_main: %lu stages in %lu pieces\n", (unsigned long)parameters.num_options, (unsigned long)pieces);
  return stage_ret != NO_ERROR ? ERROR_LIBRARY_CALL : 0;
}
