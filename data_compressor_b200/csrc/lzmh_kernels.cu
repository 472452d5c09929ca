// lzmh_kernels.cu -- LZMH (DCLib/src/lzmh.c) on sm_100a.  Placeholder until the kernels land: every entry point
// fails loudly (there is no CPU fallback).
#include <cuda_runtime.h>

#include "../../include/dc_cuda.h"
#include "kernels.h"

extern "C" {

int dc_lzmh_stage_host(dc_ctx *, int, const uint8_t *, uint64_t, uint8_t *, uint64_t, uint64_t *) { return DC_ERR_INVALID_MODE; }

int dc_lzmh_encode_device(dc_ctx *, const uint8_t *, const uint64_t *, uint64_t, uint8_t *, uint64_t, uint64_t *, uint64_t *, int32_t *,
                          uint64_t *)
{
  return DC_ERR_INVALID_MODE;
}

int dc_lzmh_decode_device(dc_ctx *, const uint8_t *, uint64_t, const uint64_t *, const uint64_t *, uint64_t, uint8_t *, const uint64_t *,
                          uint64_t *, int32_t *)
{
  return DC_ERR_INVALID_MODE;
}
}
