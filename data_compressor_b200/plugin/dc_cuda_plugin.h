/* dc_cuda_plugin.h -- declarations of the GPU plugins for DCLib/src/enc_dec.c */
#ifndef _DC_CUDA_PLUGIN_H
#define _DC_CUDA_PLUGIN_H

#include "enc_dec.h"

io_int_t EncodeDEGA_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t DecodeDEGA_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);

#ifndef DC_CUDA_REPLACE_REFERENCE
io_int_t Normalize_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t Denormalize_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t EncodeDifferential_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t DecodeDifferential_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t EncodeSEG_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t DecodeSEG_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t EncodeBAC_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t DecodeBAC_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t EncodeLZMH_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t DecodeLZMH_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
io_int_t Copy_CUDA(bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf, const options_t * const options);
#endif

/* For a chain fuser on the CLI side (dc_fused_main.c) */
int DcCudaStageOf(enc_dec_function_t * const function, const int encode);
io_int_t DcCudaRunChain(options_t * const * const options, const size_t count, bit_file_buffer_t * const in_bit_buf, bit_file_buffer_t * const out_bit_buf,
                        size_t * const failed_stage);

#endif
