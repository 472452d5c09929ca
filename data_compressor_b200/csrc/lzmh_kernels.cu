// lzmh_kernels.cu -- __global__ wrappers and launchers of the LZMH kernels (bodies: lzmh_device.cuh)
#include <cuda_runtime.h>
#include <stdint.h>

#include "lzmh_device.cuh"

namespace dc {

void count_launch(int n);

static const int kLzEncWarps = 8;

__global__ void __launch_bounds__(kLzEncWarps * 32) lzmh_encode_kernel(const LzmhArgs a)
{
  __shared__ __align__(16) uint8_t win[kLzEncWarps][kLzWin];
  __shared__ uint8_t lsym[kLzEncWarps][kLzList];
  __shared__ uint32_t lcnt[kLzEncWarps][kLzList];
  const int wid = threadIdx.x >> 5;
  const uint64_t s = (uint64_t)blockIdx.x * kLzEncWarps + wid;
  if (s < a.n_streams)
    lzmh_encode_warp(a, s, win[wid], lsym[wid], lcnt[wid]);
}

__global__ void __launch_bounds__(64) lzmh_decode_kernel(const LzmhArgs a)
{
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < a.n_streams)
    lzmh_decode_body(a, s);
}

void launch_lzmh_encode(const LzmhArgs &a, cudaStream_t st)
{
  if (a.n_streams == 0)
    return;
  const unsigned grid = (unsigned)((a.n_streams + kLzEncWarps - 1) / kLzEncWarps);
  lzmh_encode_kernel<<<grid, kLzEncWarps * 32, 0, st>>>(a);
  count_launch(1);
}

void launch_lzmh_decode(const LzmhArgs &a, cudaStream_t st)
{
  if (a.n_streams == 0)
    return;
  const unsigned block = 64;
  const unsigned grid = (unsigned)((a.n_streams + block - 1) / block);
  lzmh_decode_kernel<<<grid, block, 0, st>>>(a);
  count_launch(1);
}

}  // namespace dc
