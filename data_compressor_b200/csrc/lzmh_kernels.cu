// lzmh_kernels.cu -- __global__ wrappers and launchers of the LZMH kernels (bodies: lzmh_device.cuh)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "lzmh_device.cuh"

namespace dc {

void count_launch(int n);

static const int kLzEncWarps = 8;

__global__ void __launch_bounds__(kLzEncWarps * 32) lzmh_encode_kernel(const LzmhArgs a)
{
  __shared__ __align__(16) uint8_t win[kLzEncWarps][kLzWin + kLzWinPad];
  __shared__ uint8_t lsym[kLzEncWarps][kLzList];
  __shared__ uint32_t lcnt[kLzEncWarps][kLzList];
  const int wid = threadIdx.x >> 5;
  const uint64_t s = (uint64_t)blockIdx.x * kLzEncWarps + wid;
  if (s < a.n_streams)
    lzmh_encode_warp(a, s, win[wid], lsym[wid], lcnt[wid]);
}

// The decoder is one sequential, branchy parse per block.  With 32 blocks per warp the lanes diverge (9 of 32 active
// on average) and a 4096-block job is 128 warps on 148 SMs, i.e. bound by the latency of single warps.  So only the
// first `lanes` lanes of every warp take a block: fewer lanes per warp = less divergence per warp and more warps to
// hide latency with, at the price of more warp instructions in total.  The launcher picks `lanes` so that the job is
// about one full wave of warps (model and measurements: profiles/r1_lzmh_lanes.txt).
__global__ void __launch_bounds__(128) lzmh_decode_kernel(const LzmhArgs a, const uint32_t lanes)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t s = warp * lanes + lane;
  if (lane < lanes && s < a.n_streams)
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
  // lanes per warp: a power of two such that the job is at most ~48 warps per SM (148 SMs)
  uint32_t lanes = 1;
  const char *force = getenv("DC_LZMH_DECODE_LANES");
  if (force && atoi(force) >= 1 && atoi(force) <= 32)
    lanes = (uint32_t)atoi(force);
  else
    while (lanes < 32 && (a.n_streams + lanes - 1) / lanes > 148ull * 48)
      lanes *= 2;
  const unsigned block = 128;
  const uint64_t warps = (a.n_streams + lanes - 1) / lanes;
  const unsigned grid = (unsigned)((warps * 32 + block - 1) / block);
  lzmh_decode_kernel<<<grid, block, 0, st>>>(a, lanes);
  count_launch(1);
}

}  // namespace dc
