// Micro-benchmark: dependent-issue latency (cycles) of the instructions on the arithmetic coder's critical path,
// sm_100a, ONE warp per SM, one dependent chain.   nvcc -arch=sm_100a -o latency latency.cu && ./latency
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITER 2048

#define CHAIN(NAME, ...)                                                                                   \
  __global__ void __launch_bounds__(32, 1) NAME(uint32_t *out, uint32_t seed, unsigned long long *cyc)     \
  {                                                                                                        \
    __shared__ uint2 tab[256];                                                                             \
    for (int i = threadIdx.x; i < 256; i += 32) tab[i] = make_uint2((i * 37u + seed) & 255u, i);           \
    __syncwarp();                                                                                          \
    uint32_t a = seed + threadIdx.x * 2654435761u, k = (seed + threadIdx.x) | 1, z = 0;                    \
    const unsigned long long t0 = clock64();                                                               \
    _Pragma("unroll 1") for (int i = 0; i < ITER; i++)                                                     \
    {                                                                                                      \
      _Pragma("unroll") for (int j = 0; j < 8; j++) { __VA_ARGS__ }                                        \
    }                                                                                                      \
    const unsigned long long t1 = clock64();                                                               \
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + z;                                                    \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                       \
  }

CHAIN(l_imad, asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a) : "r"(k));)
CHAIN(l_imadhi, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(k | 0x80000000u));)
CHAIN(l_iadd, asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(k));)
CHAIN(l_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(k), "r"(z));)
CHAIN(l_shf, asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(k), "r"(k & 7));)
CHAIN(l_shl, asm volatile("shl.b32 %0, %0, %1; or.b32 %0, %0, 1;" : "+r"(a) : "r"(k & 3));)
CHAIN(l_flo, asm volatile("bfind.shiftamt.u32 %0, %0; or.b32 %0, %0, %1;" : "+r"(a) : "r"(k));)
CHAIN(l_sel, asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(a) : "r"(k), "r"(seed));)
CHAIN(l_lds, { const uint2 e = tab[a & 255]; a = e.x; z += e.y; })
CHAIN(l_imadwide, { unsigned long long x = ((unsigned long long)a << 32) | k; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(a), "r"(k)); a = (uint32_t)(x >> 32) ^ (uint32_t)x; })

typedef void (*kern_t)(uint32_t *, uint32_t, unsigned long long *);

int main()
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t *out;
  unsigned long long *cyc, h[1024];
  cudaMalloc(&out, sms * 32 * 4);
  cudaMalloc(&cyc, sms * 8);
  struct { const char *name; kern_t k; int per; } tests[] = {
    { "IMAD", l_imad, 1 }, { "IMAD.HI", l_imadhi, 1 }, { "IADD", l_iadd, 1 }, { "LOP3", l_lop3, 1 }, { "SHF (funnel)", l_shf, 1 },
    { "SHL + OR", l_shl, 2 }, { "FLO + OR", l_flo, 2 }, { "SETP + SELP", l_sel, 2 }, { "LDS.64 + AND", l_lds, 2 }, { "IMAD.WIDE + XOR", l_imadwide, 2 },
  };
  for (auto &t : tests)
  {
    t.k<<<sms, 32>>>(out, 12345, cyc);
    t.k<<<sms, 32>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    printf("%-20s %7.2f cycles per chain step (%d dependent instruction(s) per step)  (%s)\n", t.name, avg / (ITER * 8.0), t.per, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
