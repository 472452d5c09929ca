// lzmh_kernels.cu -- __global__ wrappers and launchers of the LZMH kernels (bodies: lzmh_device.cuh)
#include <cuda_runtime.h>
#include <stdint.h>

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

static const int kLzDecWarps = 4;

__global__ void __launch_bounds__(kLzDecWarps * 32) lzmh_decode_kernel(const LzmhArgs a)
{
  __shared__ __align__(16) uint8_t ring[kLzDecWarps][kLzOutRing];
  __shared__ __align__(4) uint8_t lsym[kLzDecWarps][kLzList];
  __shared__ uint32_t lcnt[kLzDecWarps][kLzList];
  __shared__ uint8_t modtab[kLzModTab];
  for (uint32_t i = threadIdx.x; i < kLzModTab; i += blockDim.x)
  {
    const uint32_t back = i < 32 * 32 ? i >> 5 : i - 32 * 32, x = i < 32 * 32 ? i & 31 : 32;
    modtab[i] = back ? (uint8_t)(x % back) : 0;
  }
  __syncthreads();
  const int wid = threadIdx.x >> 5;
  const uint64_t s = (uint64_t)blockIdx.x * kLzDecWarps + wid;
  if (s < a.n_streams)
    lzmh_decode_warp(a, s, (uint32_t)__cvta_generic_to_shared(ring[wid]), (uint32_t)__cvta_generic_to_shared(lsym[wid]),
                     (uint32_t)__cvta_generic_to_shared(lcnt[wid]), (uint32_t)__cvta_generic_to_shared(modtab));
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
  const unsigned grid = (unsigned)((a.n_streams + kLzDecWarps - 1) / kLzDecWarps);
  lzmh_decode_kernel<<<grid, kLzDecWarps * 32, 0, st>>>(a);
  count_launch(1);
}

}  // namespace dc
