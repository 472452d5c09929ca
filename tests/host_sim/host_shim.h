// host_shim.h -- lets g++ compile the per-thread device bodies (data_compressor_b200/csrc/*.cuh) for the HOST.
//
// TEST HARNESS ONLY.  It exists so that `pytest -m "not gpu"` can step the exact per-thread code of the kernels
// against the oracle in a container without a GPU.  It is never linked into libdc_cuda.so and is not a fallback:
// the product fails with DC_ERR_LIBRARY_INIT when no GPU is present.
#pragma once

#define __CUDA_RUNTIME_H__   // neutralise <cuda_runtime.h>
#define DC_HOST_SIM 1

#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(x)

typedef void *cudaStream_t;
struct float4
{
  float x, y, z, w;
};

static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
  const uint64_t v = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++)
    r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)((v << (sh & 31)) >> 32);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (sh & 31));
}
template <typename T>
static inline T __ldg(const T *p)
{
  return *p;
}
static inline float __fmul_rn(float a, float b)
{
  volatile float r = a * b;
  return r;
}
static inline float __fadd_rn(float a, float b)
{
  volatile float r = a + b;
  return r;
}
static inline float __fsub_rn(float a, float b)
{
  volatile float r = a - b;
  return r;
}
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fdiv_rn(float a, float b)
{
  volatile float r = a / b;
  return r;
}
static inline long long __float2ll_rz(float v)
{
  if (v != v)
    return 0;
  if (v >= 9223372036854775808.0f)
    return INT64_MAX;
  if (v < -9223372036854775808.0f)
    return INT64_MIN;
  return (long long)v;
}
static inline float __ll2float_rn(long long v) { return (float)v; }
static inline float __ull2float_rn(unsigned long long v) { return (float)v; }
static inline float __int2float_rn(int v) { return (float)v; }
static inline unsigned int __float2uint_rz(float v)
{
  if (!(v > 0.0f))
    return 0u;
  if (v >= 4294967296.0f)
    return 0xFFFFFFFFu;
  return (unsigned int)v;
}
static inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v)
{
  const unsigned long long o = *p;
  *p = v;
  return o;
}
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
