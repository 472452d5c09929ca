// dega_kernels.cu -- sm_100a kernels of the DEGA chain (normalize -> diff -> seg -> bac) and its inverse.
//
// Work decomposition: one independent stream (= one reference invocation, see include/dc_cuda.h) per thread.
// The arithmetic coder is inherently sequential per stream (bac.c:152-162), so parallelism comes from the
// number of streams; the kernels are integer-issue bound, not HBM bound.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "dega_device.cuh"

namespace dc {

static std::atomic<uint64_t> g_launches{0};   // contexts may run on several host threads (dc_*_host_async)
uint64_t launch_counter() { return g_launches.load(); }
void count_launch(int n) { g_launches += (uint64_t)n; }

// Work distribution of the stream-per-thread kernels: a stream is one long sequential task (up to milliseconds), so
// what matters is that the last, partial round of tasks is spread over all SMs instead of filling a few of them:
// resident CTAs, and every warp fetches its next 32 streams from a device counter.  Warps of a thin last round then
// run with few neighbours and finish several times faster than in a full SM, which takes most of the sting out of
// stream counts that are not a multiple of what the GPU holds at once (and lets each kernel pick its occupancy by
// register need, not by wave arithmetic).
template <typename Body>
__device__ __forceinline__ void warp_tasks(unsigned int *work, uint64_t n_streams, Body body)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t n_tasks = (n_streams + 31) / 32;
  for (;;)
  {
    uint32_t w = 0;
    if (lane == 0)
      w = atomicAdd(work, 1u);
    w = __shfl_sync(0xFFFFFFFFu, w, 0);
    if (w >= n_tasks)
      break;
    const uint64_t s = (uint64_t)w * 32 + lane;
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, s < n_streams);   // the lanes that run a stream in this task
    if (s < n_streams)
      body(s, mask);
    __syncwarp();
  }
}

// SM count of the current device (cached per device: a process may drive several)
int sm_count()
{
  static int cached[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64)
    dev = 0;
  if (cached[dev] == 0)
  {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;   // B200
    cached[dev] = n;
  }
  return cached[dev];
}

// grid of a resident launch: ctas_per_sm CTAs on every SM, fewer when there is less work
static unsigned resident_grid(uint64_t n_streams, unsigned block, unsigned ctas_per_sm)
{
  const uint64_t want = (n_streams + block - 1) / block;
  const uint64_t cap = (uint64_t)sm_count() * ctas_per_sm;
  return (unsigned)(want < cap ? want : cap);
}

__global__ void __launch_bounds__(128) front_kernel(const FrontArgs a)
{
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < a.lay.n_streams)
    front_body(a, s);
}

#ifndef FRONT_MIN_CTAS
#define FRONT_MIN_CTAS 6
#endif
// Resident CTAs of 128 threads fed by the warp-task scheduler; measured 0.93 ms at 6 CTAs per SM (48 registers, no
// spills), 1.02 at 10, 1.12 at 12, 1.30 at 16 (spills): the kernel is issue bound, extra warps only add pressure
__global__ void __launch_bounds__(128, FRONT_MIN_CTAS) front_dega_kernel(const FrontArgs a)
{
  warp_tasks(a.work, a.lay.n_streams, [&](uint64_t s, unsigned) { front_dega_body(a, s); });
}

void launch_front(const FrontArgs &a, cudaStream_t st)
{
  if (a.lay.n_streams == 0)
    return;
  const unsigned block = 128;
  const unsigned grid = (unsigned)((a.lay.n_streams + block - 1) / block);
  if (a.in_is_float && a.do_norm && a.do_diff && a.out_seg)
  {
    cudaMemsetAsync(a.work, 0, sizeof(unsigned int), st);
    front_dega_kernel<<<resident_grid(a.lay.n_streams, block, FRONT_MIN_CTAS), block, 0, st>>>(a);
  }
  else
    front_kernel<<<grid, block, 0, st>>>(a);
  g_launches++;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) back_kernel(const BackArgs a)
{
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < a.n_streams)
    back_body(a, s);
}

// measured (final kernel): 1.235 ms at 12 CTAs per SM, 1.245 at 11, 1.255 at 10, 1.29 at 8, 1.39+ at 13 and more (32 registers: spills)
#ifndef BACK_MIN_CTAS
#define BACK_MIN_CTAS 12
#endif
// CTAS: what the build is for.  BACK_MIN_CTAS gets 40 registers; 8 gets 48 and is 4 % faster where 10 or fewer CTAs per SM are
// wanted anyway (launch_back picks the count by the shape of the job).
template <int MODE, int CTAS = BACK_MIN_CTAS>
__global__ void __launch_bounds__(128, CTAS) back_dega_kernel(const BackArgs a)
{
  __shared__ float stage[128 * kBackStage];   // decoded samples per thread, leaving as 128-bit stores of whole sectors
  float *mine = stage + threadIdx.x * kBackStage;
  warp_tasks(a.work, a.n_streams, [&](uint64_t s, unsigned) { back_dega_body<MODE>(a, s, mine); });
}

void launch_back(const BackArgs &a, cudaStream_t st)
{
  if (a.n_streams == 0)
    return;
  const unsigned block = 128;
  const unsigned grid = (unsigned)((a.n_streams + block - 1) / block);
  // the specialised kernel reads the 4-byte aligned SEG scratch (fixed stride); anything else takes the general one
  if (a.in_seg && a.do_undiff && a.do_denorm && a.vs >= 17 && a.in.byte_off == nullptr && (a.in.stride_bytes & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(a.in.base) & 3) == 0)
  {
    cudaMemsetAsync(a.work, 0, sizeof(unsigned int), st);
    // How many CTAs per SM: the kernel is built for BACK_MIN_CTAS, and from 8 up its throughput hardly depends on the
    // number, but a last round that is a third full does cost (303 104 streams: 1.15 ms at 12 CTAs per SM = 1.33 rounds,
    // 1.05 ms at 8 = 2 rounds; 348 160 streams: 10 = 1.84 rounds).  So: the count whose rounds waste the fewest thread slots, the larger one on a tie.
    unsigned ctas = BACK_MIN_CTAS;
    uint64_t best = ~0ull;
    for (unsigned c = BACK_MIN_CTAS; c >= 8; c--)
    {
      const uint64_t per_round = (uint64_t)sm_count() * c * block;
      const uint64_t slots = (a.n_streams + per_round - 1) / per_round * c;
      if (slots < best)
      {
        best = slots;
        ctas = c;
      }
    }
    const unsigned rgrid = resident_grid(a.n_streams, block, ctas);
    const bool narrow = a.vs <= 32 && a.in.stride_bytes < (1ull << 28);   // every stream shorter than 2^31 bits
    if (narrow && a.vs == 32 && a.in.interleave && ctas <= 10)
      back_dega_kernel<2, 8><<<rgrid, block, 0, st>>>(a);
    else if (narrow && a.vs == 32 && a.in.interleave)
      back_dega_kernel<2><<<rgrid, block, 0, st>>>(a);
    else if (narrow)
      back_dega_kernel<1><<<rgrid, block, 0, st>>>(a);
    else
      back_dega_kernel<0><<<rgrid, block, 0, st>>>(a);
  }
  else
    back_kernel<<<grid, block, 0, st>>>(a);
  g_launches++;
}

// ---------------------------------------------------------------------------------------------------------
// A-XDR fast path (normalize only, normalize.c:9-42) for equal-length streams with word-aligned records.
// Memory bound: 4 B read + vs/8 B written per sample, every access coalesced.
// One 32-bit word of a record: every vs-bit field overlapping stream bits [32*wi, 32*wi + 32)
__device__ __forceinline__ uint32_t axdr_word(const float *src, uint32_t len, uint32_t wi, uint32_t vs, uint64_t mask, float nf, float lo_bound,
                                              float hi_bound, bool &ok)
{
  const uint32_t bit0 = wi * 32;
  uint32_t word = 0;
  for (uint32_t i = bit0 / vs; i * vs < bit0 + 32 && i < len; i++)
  {
    int64_t n;
    if (!normalize_one(__ldg(src + i), nf, lo_bound, hi_bound, n))
    {
      ok = false;
      return 0;
    }
    const uint64_t u = (uint64_t)n & mask;
    const int shift = (int)(i * vs + vs) - (int)(bit0 + 32);   // > 0: the field ends after this word
    word |= shift >= 0 ? (uint32_t)(u >> shift) : (uint32_t)(u << (-shift));
  }
  return bswap32(word);
}

// Four consecutive output words per thread (one 128-bit store).  VS32: a word is exactly one sample, so the four
// inputs come from one 128-bit load.
template <bool VS32>
__global__ void __launch_bounds__(256) axdr_encode_kernel(const AxdrArgs a)
{
  const uint32_t vs = a.vs;
  const uint64_t mask = low_mask(vs);
  const float lo_bound = -(float)(1ull << (vs - 1));
  const float hi_bound = (float)((1ull << (vs - 1)) - 1);
  const uint32_t wps = (uint32_t)a.words_per_stream, len = (uint32_t)a.len;
  const uint64_t total_words = a.n_streams * a.words_per_stream;
  const uint64_t total_quads = (total_words + 3) / 4;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint64_t g = q * 4;
    uint64_t s = g / wps;
    uint32_t wi = (uint32_t)(g - s * wps);
    uint32_t out[4];
    bool ok = true;
    if (VS32 && g + 4 <= total_words)
    {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(a.samples_in + g));   // wps == len: same flat index
      const float f[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        int64_t n = 0;
        const bool okj = normalize_one(f[j], a.nf, lo_bound, hi_bound, n);
        out[j] = bswap32((uint32_t)((uint64_t)n & mask));
        if (!okj)
        {
          a.status[(g + j) / wps] = DC_ERR_INVALID_VALUE_D;
          ok = false;
        }
      }
    }
    else
    {
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        out[j] = 0;
        if (g + j < total_words)
        {
          bool okj = true;
          out[j] = axdr_word(a.samples_in + s * len, len, wi, vs, mask, a.nf, lo_bound, hi_bound, okj);
          if (!okj)
          {
            a.status[s] = DC_ERR_INVALID_VALUE_D;
            ok = false;
          }
          if (wi == 0)
          {
            a.bit_len[s] = (uint64_t)len * vs;
            a.byte_off[s] = s * wps * 4ull;
          }
          if (++wi == wps)
          {
            wi = 0;
            s++;
          }
        }
      }
    }
    if (VS32 && g + 4 <= total_words)   // index arrays for the vector path: a stream starts inside this quad at most once per wps words
    {
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        if (wi == 0)
        {
          a.bit_len[s] = (uint64_t)len * vs;
          a.byte_off[s] = s * wps * 4ull;
        }
        if (++wi == wps)
        {
          wi = 0;
          s++;
        }
      }
    }
    if (!ok)
      atomicExch(a.err_flag, 1ull);
    if (g + 4 <= total_words)
      *reinterpret_cast<uint4 *>(a.payload_words + g) = make_uint4(out[0], out[1], out[2], out[3]);
    else
      for (int j = 0; j < 4 && g + j < total_words; j++)
        a.payload_words[g + j] = out[j];
    if (q == 0)
      a.byte_off[a.n_streams] = total_words * 4ull;
  }
}

// Four consecutive samples per thread (one 128-bit store)
__global__ void __launch_bounds__(256) axdr_decode_kernel(const AxdrArgs a)
{
  const uint32_t vs = a.vs;
  const uint64_t mask = low_mask(vs);
  const uint32_t len = (uint32_t)a.len, wps = (uint32_t)a.words_per_stream;
  const uint64_t total = a.n_streams * a.len;
  const uint32_t *words = reinterpret_cast<const uint32_t *>(a.payload);
  const uint64_t total_quads = (total + 3) / 4;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint64_t g = q * 4;
    uint64_t s = g / len;
    uint32_t i = (uint32_t)(g - s * len);
    float f[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      if (g + j < total)
      {
        const uint32_t bit = i * vs;
        const uint32_t *w = words + s * wps + (bit >> 5);
        const uint32_t sh = bit & 31;
        // up to 96 stream bits starting at a word boundary cover any field of at most 64 bits
        const uint64_t w0 = bswap32(__ldg(w));
        const uint64_t w1 = sh + vs > 32 ? bswap32(__ldg(w + 1)) : 0;
        const uint64_t w2 = sh + vs > 64 ? bswap32(__ldg(w + 2)) : 0;
        const uint64_t hi = (w0 << 32) | w1;
        uint64_t u = sh ? ((hi << sh) | (w2 >> (32 - sh))) : hi;   // field left-aligned
        u = (u >> (64 - vs)) & mask;
        f[j] = __fdiv_rn(__ll2float_rn(sext(u, vs)), a.nf);   // normalize.c:36-39
        if (i == 0)
        {
          a.n_out[s] = len;
          a.status[s] = 0;
        }
        if (++i == len)
        {
          i = 0;
          s++;
        }
      }
    }
    if (g + 4 <= total)
      *reinterpret_cast<float4 *>(a.samples_out + g) = make_float4(f[0], f[1], f[2], f[3]);
    else
      for (int j = 0; j < 4 && g + j < total; j++)
        a.samples_out[g + j] = f[j];
  }
}

// ---- A-XDR, value size 32: one sample = one word.  Pure streaming (4 B in, 4 B out per sample), so the kernels are
// written for instruction economy: 32-bit indices, one 128-bit load and store per thread per trip, the stream index
// by a multiply-high instead of a 64-bit division, index arrays written by the thread that holds a stream's first word.
struct Magic32
{
  uint32_t mul, shift;   // x / d == umulhi(x, mul) >> shift for every 32-bit x (d >= 2; round-up reciprocal with a 33rd bit folded below)
  uint32_t add;          // 1 when the 33-bit form is needed: q = (((x - t) >> 1) + t) >> (shift - 1), t = umulhi(x, mul)
};

static Magic32 make_magic32(uint32_t d)
{
  // Granlund-Montgomery: l = ceil(log2 d), m = floor(2^32 * (2^l - d) / d) + 1, q = (t + ((x - t) >> 1)) >> (l - 1)
  Magic32 g;
  uint32_t l = 0;
  while ((1ull << l) < d)
    l++;
  g.mul = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
  g.shift = l;
  g.add = 1;
  return g;
}

__device__ __forceinline__ uint32_t div_magic32(uint32_t x, const Magic32 g)
{
  const uint32_t t = __umulhi(x, g.mul);
  return (t + ((x - t) >> 1)) >> (g.shift - 1);
}

// normalize.c:16-23 for value size 32; returns false when the range check fails
__device__ __forceinline__ bool normalize32(float v, float nf, uint32_t &n)
{
  if (v > 0.0f)
    v = __fadd_rn(__fmul_rn(v, nf), 0.5f);
  else if (v < 0.0f)
    v = __fsub_rn(__fmul_rn(v, nf), 0.5f);
  // -(float)2^31 <= v <= (float)(2^31 - 1) == 2^31 (normalize.c:21-22); NaN fails
  const bool ok = v >= -2147483648.0f && v <= 2147483648.0f;
  n = v >= 2147483648.0f ? 0x80000000u : (uint32_t)__float2int_rz(v);   // low 32 bits of the int64 truncation
  return ok;
}

__global__ void __launch_bounds__(256) axdr_encode32_kernel(const AxdrArgs a, const Magic32 by_wps, const uint32_t total_quads)
{
  const uint32_t wps = (uint32_t)a.words_per_stream;
  const float4 *in = reinterpret_cast<const float4 *>(a.samples_in);
  uint4 *out = reinterpret_cast<uint4 *>(a.payload_words);
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += gridDim.x * blockDim.x)
  {
    const float4 v = __ldg(in + q);
    uint32_t n0, n1, n2, n3;
    const bool ok0 = normalize32(v.x, a.nf, n0), ok1 = normalize32(v.y, a.nf, n1);
    const bool ok2 = normalize32(v.z, a.nf, n2), ok3 = normalize32(v.w, a.nf, n3);
    out[q] = make_uint4(bswap32(n0), bswap32(n1), bswap32(n2), bswap32(n3));
    const uint32_t g = q * 4;
    const uint32_t s = div_magic32(g, by_wps);
    const uint32_t wi = g - s * wps;
    // a stream starts inside this quad at most once (wps >= 4): word 0 of s, or word (wps - wi) of s + 1
    if (wi == 0 || wps - wi < 4)
    {
      const uint32_t t = wi == 0 ? s : s + 1;
      a.bit_len[t] = (uint64_t)wps * 32;
      a.byte_off[t] = (uint64_t)t * wps * 4;
    }
    if (!(ok0 && ok1 && ok2 && ok3))   // rare: remember which stream, the caller redoes the batch on the general path
    {
      const bool okj[4] = { ok0, ok1, ok2, ok3 };
      for (uint32_t j = 0; j < 4; j++)
        if (!okj[j])
          a.status[(g + j) / wps] = DC_ERR_INVALID_VALUE_D;
      atomicExch(a.err_flag, 1ull);
    }
    if (q == 0)
      a.byte_off[a.n_streams] = (uint64_t)total_quads * 16;
  }
}

__global__ void __launch_bounds__(256) axdr_decode32_kernel(const AxdrArgs a, const Magic32 by_len, const uint32_t total_quads)
{
  const uint32_t len = (uint32_t)a.len;
  const uint4 *in = reinterpret_cast<const uint4 *>(a.payload);
  float4 *out = reinterpret_cast<float4 *>(a.samples_out);
  const float nf = a.nf;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += gridDim.x * blockDim.x)
  {
    const uint4 w = __ldg(in + q);
    float4 f;   // normalize.c:36-39: sign-extended value as float32, IEEE division
    f.x = __fdiv_rn(__int2float_rn((int32_t)bswap32(w.x)), nf);
    f.y = __fdiv_rn(__int2float_rn((int32_t)bswap32(w.y)), nf);
    f.z = __fdiv_rn(__int2float_rn((int32_t)bswap32(w.z)), nf);
    f.w = __fdiv_rn(__int2float_rn((int32_t)bswap32(w.w)), nf);
    out[q] = f;
    const uint32_t g = q * 4;
    const uint32_t s = div_magic32(g, by_len);
    const uint32_t i = g - s * len;
    if (i == 0 || len - i < 4)
    {
      const uint32_t t = i == 0 ? s : s + 1;
      a.n_out[t] = len;
      a.status[t] = 0;
    }
  }
}

// ---- A-XDR, value sizes below 32: the records of equal-length streams are whole words, so the payload is one flat
// string of vs-bit fields and a tile of 1024 samples is exactly 32 * vs words whatever the stream length.  A CTA loads a
// tile with 128-bit loads, normalises every sample once into shared memory, then every thread assembles whole output
// words from it: all global traffic is coalesced and each sample is read once (the general kernel above loads a sample
// once per word it touches, scalar).  Memory bound: 4 B in + vs / 8 B out per sample.
static const int kAxdrTile = 1024;

// normalize.c:16-23 for a value size below 32; n = the low 32 bits of the truncated value
__device__ __forceinline__ bool normalize_small(float v, float nf, float lo_bound, float hi_bound, uint32_t &n)
{
  if (v > 0.0f)
    v = __fadd_rn(__fmul_rn(v, nf), 0.5f);
  else if (v < 0.0f)
    v = __fsub_rn(__fmul_rn(v, nf), 0.5f);
  const bool ok = v >= lo_bound && v <= hi_bound;   // |v| <= 2^30 when it holds (NaN fails)
  n = ok ? (uint32_t)__float2int_rz(v) : 0u;
  return ok;
}

__global__ void __launch_bounds__(256) axdr_encode_tiled_kernel(const AxdrArgs a, const uint64_t total_samples)
{
  __shared__ uint32_t vals[kAxdrTile];
  const uint32_t vs = a.vs, mask = (1u << vs) - 1u;
  const float lo_bound = -(float)(1u << (vs - 1));
  const float hi_bound = (float)((1u << (vs - 1)) - 1);
  const uint32_t tid = threadIdx.x;
  const uint64_t gsize = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + tid; s < a.n_streams; s += gsize)   // the index: equal records
  {
    a.bit_len[s] = a.len * vs;
    a.byte_off[s] = s * a.words_per_stream * 4ull;
  }
  if (blockIdx.x == 0 && tid == 0)
    a.byte_off[a.n_streams] = a.n_streams * a.words_per_stream * 4ull;
  const uint64_t n_tiles = (total_samples + kAxdrTile - 1) / kAxdrTile;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const uint64_t base = tile * kAxdrTile;
    const uint32_t here = (uint32_t)min((uint64_t)kAxdrTile, total_samples - base);
    float f[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    const uint32_t i = 4 * tid;
    if (i + 4 <= here)
    {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(a.samples_in + base) + tid);
      f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    else
      for (uint32_t j = 0; j < 4; j++)
        if (i + j < here)
          f[j] = __ldg(a.samples_in + base + i + j);
    bool ok = true;
#pragma unroll
    for (uint32_t j = 0; j < 4; j++)
    {
      uint32_t n;
      const bool okj = normalize_small(f[j], a.nf, lo_bound, hi_bound, n);
      vals[i + j] = n & mask;
      if (!okj && i + j < here)
      {
        a.status[(base + i + j) / a.len] = DC_ERR_INVALID_VALUE_D;
        ok = false;
      }
    }
    if (!ok)
      atomicExch(a.err_flag, 1ull);   // the caller redoes the batch on the general path (a failed stream changes the layout)
    __syncthreads();
    const uint32_t n_words = here * vs / 32;   // exact: tiles and the whole batch are whole words
    uint32_t *out = a.payload_words + base / 32 * vs;
    for (uint32_t w = tid; w < n_words; w += 256)
    {
      const uint32_t bit0 = w * 32;
      uint32_t k = bit0 / vs;
      uint32_t nb = vs - (bit0 - k * vs);            // bits of field k that belong to this word (and the ones after it)
      uint64_t acc = vals[k++] & ((1u << nb) - 1u);
      while (nb < 32)
      {
        acc = (acc << vs) | vals[k++];
        nb += vs;
      }
      out[w] = bswap32((uint32_t)(acc >> (nb - 32)));
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) axdr_decode_tiled_kernel(const AxdrArgs a, const uint64_t total_samples)
{
  __shared__ uint32_t words[kAxdrTile + 1];   // a tile of 1024 samples is 32 * vs <= 992 words; one more for the last field's window
  const uint32_t vs = a.vs;
  const uint32_t tid = threadIdx.x;
  const uint64_t gsize = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + tid; s < a.n_streams; s += gsize)
  {
    a.n_out[s] = a.len;
    a.status[s] = 0;
  }
  const uint32_t *in = reinterpret_cast<const uint32_t *>(a.payload);
  const uint64_t n_tiles = (total_samples + kAxdrTile - 1) / kAxdrTile;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const uint64_t base = tile * kAxdrTile;
    const uint32_t here = (uint32_t)min((uint64_t)kAxdrTile, total_samples - base);
    const uint32_t n_words = here * vs / 32;
    const uint32_t *src = in + base / 32 * vs;
    for (uint32_t w = tid; w < n_words; w += 256)
      words[w] = bswap32(__ldg(src + w));
    if (tid == 0)
      words[n_words] = 0;
    __syncthreads();
    const uint32_t i = 4 * tid;
    float f[4];
#pragma unroll
    for (uint32_t j = 0; j < 4; j++)
    {
      const uint32_t bit = (i + j) * vs, w = min(bit >> 5, n_words - (n_words ? 1u : 0u)), sh = bit & 31;
      const uint64_t win = ((uint64_t)words[w] << 32) | words[w + 1];
      const uint32_t u = (uint32_t)((win << sh) >> (64 - vs));
      const int32_t v = (int32_t)(u << (32 - vs)) >> (32 - vs);   // EXTEND_IO_INT_SIGN, io_macros.h:89-90
      f[j] = __fdiv_rn(__int2float_rn(v), a.nf);                  // normalize.c:36-39
    }
    if (i + 4 <= here)
      reinterpret_cast<float4 *>(a.samples_out + base)[tid] = make_float4(f[0], f[1], f[2], f[3]);
    else
      for (uint32_t j = 0; j < 4; j++)
        if (i + j < here)
          a.samples_out[base + i + j] = f[j];
    __syncthreads();
  }
}

// ---- A-XDR, value size 24 (the other size the configs name): four samples are exactly three words, so a thread turns
// one 128-bit load into three word stores (the 32 threads of a warp write 384 contiguous bytes) and back; no shared
// memory, no barrier, same structure as the value-size-32 kernels.
__global__ void __launch_bounds__(256) axdr_encode24_kernel(const AxdrArgs a, const uint32_t total_quads)
{
  const float lo_bound = -8388608.0f, hi_bound = 8388607.0f;   // -(float)2^23, (float)(2^23 - 1): normalize.c:21-22
  const float4 *in = reinterpret_cast<const float4 *>(a.samples_in);
  const uint64_t gsize = (uint64_t)gridDim.x * blockDim.x, gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t s = gtid; s < a.n_streams; s += gsize)   // the index: equal records
  {
    a.bit_len[s] = a.len * 24;
    a.byte_off[s] = s * a.words_per_stream * 4ull;
  }
  if (gtid == 0)
    a.byte_off[a.n_streams] = a.n_streams * a.words_per_stream * 4ull;
  for (uint32_t q = (uint32_t)gtid; q < total_quads; q += (uint32_t)gsize)
  {
    const float4 v = __ldg(in + q);
    uint32_t n0, n1, n2, n3;
    const bool ok0 = normalize_small(v.x, a.nf, lo_bound, hi_bound, n0), ok1 = normalize_small(v.y, a.nf, lo_bound, hi_bound, n1);
    const bool ok2 = normalize_small(v.z, a.nf, lo_bound, hi_bound, n2), ok3 = normalize_small(v.w, a.nf, lo_bound, hi_bound, n3);
    n0 &= 0xFFFFFFu; n1 &= 0xFFFFFFu; n2 &= 0xFFFFFFu; n3 &= 0xFFFFFFu;
    uint32_t *out = a.payload_words + 3ull * q;
    out[0] = bswap32((n0 << 8) | (n1 >> 16));
    out[1] = bswap32((n1 << 16) | (n2 >> 8));
    out[2] = bswap32((n2 << 24) | n3);
    if (!(ok0 && ok1 && ok2 && ok3))   // rare: remember which stream, the caller redoes the batch on the general path
    {
      const bool okj[4] = { ok0, ok1, ok2, ok3 };
      for (uint32_t j = 0; j < 4; j++)
        if (!okj[j])
          a.status[(4ull * q + j) / a.len] = DC_ERR_INVALID_VALUE_D;
      atomicExch(a.err_flag, 1ull);
    }
  }
}

__global__ void __launch_bounds__(256) axdr_decode24_kernel(const AxdrArgs a, const uint32_t total_quads)
{
  const uint32_t *in = reinterpret_cast<const uint32_t *>(a.payload);
  float4 *out = reinterpret_cast<float4 *>(a.samples_out);
  const float nf = a.nf;
  const uint64_t gsize = (uint64_t)gridDim.x * blockDim.x, gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t s = gtid; s < a.n_streams; s += gsize)
  {
    a.n_out[s] = a.len;
    a.status[s] = 0;
  }
  for (uint32_t q = (uint32_t)gtid; q < total_quads; q += (uint32_t)gsize)
  {
    const uint32_t *w = in + 3ull * q;
    const uint32_t w0 = bswap32(__ldg(w)), w1 = bswap32(__ldg(w + 1)), w2 = bswap32(__ldg(w + 2));
    // four 24-bit fields, sign-extended (EXTEND_IO_INT_SIGN, io_macros.h:89-90): shift the field to the top, then down again
    const int32_t v0 = (int32_t)w0 >> 8;
    const int32_t v1 = (int32_t)((w0 << 24) | (w1 >> 8)) >> 8;
    const int32_t v2 = (int32_t)((w1 << 16) | (w2 >> 16)) >> 8;
    const int32_t v3 = (int32_t)(w2 << 8) >> 8;
    float4 f;   // normalize.c:36-39
    f.x = __fdiv_rn(__int2float_rn(v0), nf);
    f.y = __fdiv_rn(__int2float_rn(v1), nf);
    f.z = __fdiv_rn(__int2float_rn(v2), nf);
    f.w = __fdiv_rn(__int2float_rn(v3), nf);
    out[q] = f;
  }
}

void launch_axdr_encode(const AxdrArgs &a, cudaStream_t st)
{
  const uint64_t total = (a.n_streams * a.words_per_stream + 3) / 4;
  if (total == 0)
    return;
  const unsigned block = 256;
  const uint64_t want = (total + block - 1) / block;
  const uint64_t cap = (uint64_t)sm_count() * 16;                                // grid-stride over a multiple of the SM count
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const uint64_t total_words = a.n_streams * a.words_per_stream;
  if (a.vs == 24 && (reinterpret_cast<uintptr_t>(a.samples_in) & 15) == 0 && (a.n_streams * a.len) % 4 == 0 && a.n_streams * a.len < (1ull << 33))
    axdr_encode24_kernel<<<grid, block, 0, st>>>(a, (uint32_t)(a.n_streams * a.len / 4));
  else if (a.vs < 32 && (reinterpret_cast<uintptr_t>(a.samples_in) & 15) == 0)
  {
    const uint64_t total_samples = a.n_streams * a.len, tiles = (total_samples + kAxdrTile - 1) / kAxdrTile;
    axdr_encode_tiled_kernel<<<(unsigned)(tiles < cap ? tiles : cap), block, 0, st>>>(a, total_samples);
  }
  else if (a.vs == 32 && (reinterpret_cast<uintptr_t>(a.samples_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.payload_words) & 15) == 0 &&
      a.words_per_stream >= 4 && total_words % 4 == 0 && total_words < (1ull << 32))
    axdr_encode32_kernel<<<grid, block, 0, st>>>(a, make_magic32((uint32_t)a.words_per_stream), (uint32_t)(total_words / 4));
  else if (a.vs == 32 && (reinterpret_cast<uintptr_t>(a.samples_in) & 15) == 0)
    axdr_encode_kernel<true><<<grid, block, 0, st>>>(a);
  else
    axdr_encode_kernel<false><<<grid, block, 0, st>>>(a);
  g_launches++;
}

void launch_axdr_decode(const AxdrArgs &a, cudaStream_t st)
{
  const uint64_t total = (a.n_streams * a.len + 3) / 4;
  if (total == 0)
    return;
  const unsigned block = 256;
  const uint64_t want = (total + block - 1) / block;
  const uint64_t cap = (uint64_t)sm_count() * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const uint64_t total_samples = a.n_streams * a.len;
  if (a.vs == 24 && (reinterpret_cast<uintptr_t>(a.samples_out) & 15) == 0 && total_samples % 4 == 0 && total_samples < (1ull << 33))
    axdr_decode24_kernel<<<grid, block, 0, st>>>(a, (uint32_t)(total_samples / 4));
  else if (a.vs < 32 && (reinterpret_cast<uintptr_t>(a.samples_out) & 15) == 0)
  {
    const uint64_t tiles = (total_samples + kAxdrTile - 1) / kAxdrTile;
    axdr_decode_tiled_kernel<<<(unsigned)(tiles < cap ? tiles : cap), block, 0, st>>>(a, total_samples);
  }
  else if (a.vs == 32 && (reinterpret_cast<uintptr_t>(a.payload) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.samples_out) & 15) == 0 &&
      a.len >= 4 && total_samples % 4 == 0 && total_samples < (1ull << 32))
    axdr_decode32_kernel<<<grid, block, 0, st>>>(a, make_magic32((uint32_t)a.len), (uint32_t)(total_samples / 4));
  else
    axdr_decode_kernel<<<grid, block, 0, st>>>(a);
  g_launches++;
}

// ---------------------------------------------------------------------------------------------------------
// Aggregate (DCLib/src/aggregate.c:9-27): float32 sums of num_values consecutive samples per channel, added strictly
// left to right starting from 0.0f (the reference's `sum += value` loop), the last group of a channel may be short.
// One thread per output value; the groups of a warp are contiguous in memory, and each lane re-uses every 32-byte
// sector it touches for eight samples out of L1.  Memory bound: 4 B read per sample, 4 B written per group.
__global__ void __launch_bounds__(256) aggregate_kernel(const float *in, uint64_t n_channels, uint64_t channel_len, uint32_t num_values,
                                                        uint64_t groups_per_channel, float *out)
{
  const uint64_t total = n_channels * groups_per_channel;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint64_t c = g / groups_per_channel;
    const uint64_t first = (g - c * groups_per_channel) * num_values;
    const uint64_t count = channel_len - first < num_values ? channel_len - first : num_values;
    const float *src = in + c * channel_len + first;
    float sum = 0.0f;
    for (uint64_t i = 0; i < count; i++)
      sum = __fadd_rn(sum, __ldg(src + i));
    out[g] = sum;
  }
}

void launch_aggregate(const float *in, uint64_t n_channels, uint64_t channel_len, uint32_t num_values, float *out, cudaStream_t st)
{
  const uint64_t groups = (channel_len + num_values - 1) / num_values;
  const uint64_t total = n_channels * groups;
  if (total == 0)
    return;
  const unsigned block = 256;
  const uint64_t want = (total + block - 1) / block;
  const uint64_t cap = (uint64_t)sm_count() * 8;
  aggregate_kernel<<<(unsigned)(want < cap ? want : cap), block, 0, st>>>(in, n_channels, channel_len, num_values, groups, out);
  g_launches++;
}

// ---------------------------------------------------------------------------------------------------------
// Arithmetic coder kernels.  The 128 KiB division table lives in shared memory (one copy per CTA, one CTA of
// kBacThreads threads per SM); before the first halving every lane of a warp has coded the same number of
// symbols, so `total` is warp-uniform and the table read is a broadcast.
#ifndef DC_BAC_THREADS
#define DC_BAC_THREADS 1024
#endif
static const int kBacThreads = DC_BAC_THREADS;

int bac_smem_bytes() { return kDivTableEntries * (int)sizeof(DivEntry); }
int bac_threads_per_sm() { return kBacThreads; }

__device__ __forceinline__ void load_div_table(DivEntry *smem, const DivEntry *gmem)
{
  const uint4 *g = reinterpret_cast<const uint4 *>(gmem);
  uint4 *d = reinterpret_cast<uint4 *>(smem);
  for (int i = threadIdx.x; i < kDivTableEntries / 2; i += blockDim.x)
    d[i] = __ldg(g + i);
  __syncthreads();
}

// (Lane-level work distribution -- a lane that finishes its stream fetches the next one on its own, so that the 3 of 32
// lanes a warp loses to streams of unequal length stay busy -- was built and measured: 5.2 -> 8.1 ms.  The lanes of a warp
// then stand at different symbol counts, `total` differs from lane to lane, and the division-table read of every symbol
// turns from one broadcast into a many-way shared-memory bank conflict.  The warp-task form keeps the lanes in step.
// Sorting the streams by SEG length so that a warp's 32 streams are equally long keeps the broadcast but was slower too,
// 5.2 -> 6.4 ms plus 0.13 ms of sorting: the lanes then read 32 different regions of the interleaved scratch, i.e. one
// 32-byte sector per 4 bytes used, and the random-sector traffic costs more than the idle lanes did.)
template <bool ADAPTIVE>
__global__ void __launch_bounds__(kBacThreads, 1) bac_encode_kernel(const BacArgs a)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  DivEntry *tab = reinterpret_cast<DivEntry *>(smem_raw);
  load_div_table(tab, a.div_table);
  warp_tasks(a.work, a.n_streams, [&](uint64_t s, unsigned mask) { bac_encode_body<ADAPTIVE>(a, s, tab, mask); });
}

template <bool ADAPTIVE>
__global__ void __launch_bounds__(kBacThreads, 1) bac_decode_kernel(const BacArgs a)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  DivEntry *tab = reinterpret_cast<DivEntry *>(smem_raw);
  load_div_table(tab, a.div_table);
  warp_tasks(a.work, a.n_streams, [&](uint64_t s, unsigned mask) { bac_decode_body<ADAPTIVE>(a, s, tab, mask); });
}

template <typename K>
static void launch_bac(K kernel, const BacArgs &a, cudaStream_t st)
{
  if (a.n_streams == 0)
    return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bac_smem_bytes());
  cudaMemsetAsync(a.work, 0, sizeof(unsigned int), st);
  // one CTA per SM as soon as there are that many warp tasks: with few streams every SM gets a few warps (which then
  // run several times faster than in a full SM) instead of a few SMs getting 32 each
  const uint64_t n_tasks = (a.n_streams + 31) / 32;
  const unsigned grid = (unsigned)(n_tasks < (uint64_t)sm_count() ? n_tasks : (uint64_t)sm_count());
  kernel<<<grid, kBacThreads, bac_smem_bytes(), st>>>(a);
  g_launches++;
}

void launch_bac_encode(const BacArgs &a, cudaStream_t st)
{
  if (a.adaptive)
    launch_bac(bac_encode_kernel<true>, a, st);
  else
    launch_bac(bac_encode_kernel<false>, a, st);
}

void launch_bac_decode(const BacArgs &a, cudaStream_t st)
{
  if (a.adaptive)
    launch_bac(bac_decode_kernel<true>, a, st);
  else
    launch_bac(bac_decode_kernel<false>, a, st);
}

// ---------------------------------------------------------------------------------------------------------
// Compaction: device-wide exclusive prefix sum over the per-stream byte counts, then a gather into the dense
// payload.  Three small kernels: tile-local scan (4096 streams per CTA), scan of the tile totals, add-back.
static const int kScanThreads = 1024;
static const int kScanItems = 4;
static const int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *warp_sums, uint64_t &block_total)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d)
      inc += t;
  }
  if (lane == 31)
    warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0)
  {
    uint64_t ws = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, ws, d);
      if (lane >= d)
        ws += t;
    }
    warp_sums[lane] = ws;   // inclusive
  }
  __syncthreads();
  const uint64_t warp_base = wid ? warp_sums[wid - 1] : 0;
  block_total = warp_sums[(blockDim.x >> 5) - 1];
  return warp_base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(const uint64_t *nbits, uint64_t n, uint64_t *byte_off,
                                                                  uint64_t *tile_sums)
{
  __shared__ uint64_t warp_sums[32];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint64_t v[kScanItems];
  uint64_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++)
  {
    const uint64_t i = base + k;
    v[k] = i < n ? (nbits[i] + 7) >> 3 : 0;
    sum += v[k];
  }
  uint64_t total;
  uint64_t run = block_exclusive_scan(sum, warp_sums, total);
#pragma unroll
  for (int k = 0; k < kScanItems; k++)
  {
    const uint64_t i = base + k;
    if (i < n)
      byte_off[i] = run;
    run += v[k];
  }
  if (threadIdx.x == 0)
    tile_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the tile totals in place; tile_sums[n_tiles] receives the grand total
__global__ void __launch_bounds__(kScanThreads) scan_totals_kernel(uint64_t *tile_sums, uint64_t n_tiles)
{
  __shared__ uint64_t warp_sums[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0)
    carry_s = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n_tiles; base += kScanThreads)
  {
    const uint64_t i = base + threadIdx.x;
    const uint64_t v = i < n_tiles ? tile_sums[i] : 0;
    uint64_t total;
    const uint64_t ex = block_exclusive_scan(v, warp_sums, total);
    const uint64_t carry = carry_s;
    if (i < n_tiles)
      tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0)
      carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0)
    tile_sums[n_tiles] = carry_s;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(uint64_t *byte_off, uint64_t n, const uint64_t *tile_sums,
                                                                uint64_t n_tiles)
{
  const uint64_t add = tile_sums[blockIdx.x];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; k++)
  {
    const uint64_t i = base + k;
    if (i < n)
      byte_off[i] += add;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    byte_off[n] = tile_sums[n_tiles];
}

void launch_byte_scan(const uint64_t *nbits, uint64_t n, uint64_t *byte_off, uint64_t *tile_tmp, cudaStream_t st)
{
  const uint64_t n_tiles = (n + kScanTile - 1) / kScanTile;
  if (n == 0)
  {
    cudaMemsetAsync(byte_off, 0, sizeof(uint64_t), st);
    return;
  }
  scan_tiles_kernel<<<(unsigned)n_tiles, kScanThreads, 0, st>>>(nbits, n, byte_off, tile_tmp);
  scan_totals_kernel<<<1, kScanThreads, 0, st>>>(tile_tmp, n_tiles);
  scan_add_kernel<<<(unsigned)n_tiles, kScanThreads, 0, st>>>(byte_off, n, tile_tmp, n_tiles);
  g_launches += 3;
}

// Decoder-side check of a stream index that came from outside (a container file, a caller's arrays): a stream whose
// byte range does not lie inside the payload is marked ERROR_INVALID_FORMAT, and the decode kernels skip it instead
// of reading past the allocation.  Overflow-safe (bit_len and byte_off are arbitrary 64-bit values).
__global__ void __launch_bounds__(256) validate_index_kernel(const uint64_t *bit_len, const uint64_t *byte_off, uint64_t n, uint64_t payload_bytes,
                                                             int32_t *status)
{
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
    return;
  const uint64_t bits = bit_len[s], off = byte_off[s];
  const uint64_t nbytes = (bits >> 3) + ((bits & 7) != 0);
  if (off > payload_bytes || nbytes > payload_bytes - off)
    status[s] = DC_ERR_INVALID_FORMAT_D;
}

void launch_validate_index(const uint64_t *bit_len, const uint64_t *byte_off, uint64_t n, uint64_t payload_bytes, int32_t *status, cudaStream_t st)
{
  if (n == 0)
    return;
  validate_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bit_len, byte_off, n, payload_bytes, status);
  g_launches++;
}

// Hands two device scalars (the coder's overflow flag, the payload total) to the host through mapped pinned memory: no
// copy-engine operation, so the read cannot queue behind the bulk copies of another job (see the job engine in api.cu).
__global__ void publish_kernel(const uint64_t *flag, const uint64_t *total, volatile uint64_t *pub)
{
  pub[0] = *flag;
  pub[1] = *total;
  __threadfence_system();
}

void launch_publish(const uint64_t *d_flag, const uint64_t *d_total, uint64_t *pub, cudaStream_t st)
{
  publish_kernel<<<1, 1, 0, st>>>(d_flag, d_total, pub);
  g_launches++;
}

// one warp per stream, byte-granular destination
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t *scratch, uint64_t stride_bytes, const uint64_t *nbits,
                                                     const uint64_t *byte_off, uint64_t n, uint8_t *payload,
                                                     uint64_t payload_cap)
{
  const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (s >= n)
    return;
  const uint64_t nbytes = (nbits[s] + 7) >> 3;
  const uint64_t off = byte_off[s];
  if (off + nbytes > payload_cap)
    return;
  const uint8_t *src = scratch + s * stride_bytes;
  uint8_t *dst = payload + off;
  // head bytes up to 4-byte alignment of dst, then words assembled from the (aligned) source, then the tail
  const uint64_t head = min((uint64_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3), nbytes);
  if (lane < head)
    dst[lane] = src[lane];
  const uint64_t nwords = (nbytes - head) >> 2;
  const uint32_t *src_w = reinterpret_cast<const uint32_t *>(src);
  uint32_t *dst_w = reinterpret_cast<uint32_t *>(dst + head);
  const uint32_t sh = (uint32_t)head * 8;   // source words are read shifted by `head` bytes (little-endian memory)
  // four trips' loads first, then the stores: the copy is latency bound (a warp moves ~1.6 KB), so what counts is how many
  // loads a lane has in flight
  uint64_t k = lane;
  for (; k + 96 < nwords; k += 128)
  {
    uint32_t a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      a[j] = __ldg(src_w + k + 32 * j);
      b[j] = sh ? __ldg(src_w + k + 32 * j + 1) : 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
      dst_w[k + 32 * j] = sh ? __funnelshift_r(a[j], b[j], sh) : a[j];
  }
  for (; k < nwords; k += 32)
  {
    const uint32_t a = __ldg(src_w + k);
    const uint32_t v = sh ? __funnelshift_r(a, __ldg(src_w + k + 1), sh) : a;
    dst_w[k] = v;
  }
  const uint64_t done = head + nwords * 4;
  if (done + lane < nbytes)
    dst[done + lane] = src[done + lane];
}

void launch_gather(const uint8_t *scratch, uint64_t stride_bytes, const uint64_t *nbits, const uint64_t *byte_off, uint64_t n,
                   uint8_t *payload, uint64_t payload_cap, cudaStream_t st)
{
  if (n == 0)
    return;
  const unsigned block = 256;
  const unsigned grid = (unsigned)((n * 32 + block - 1) / block);
  gather_kernel<<<grid, block, 0, st>>>(scratch, stride_bytes, nbits, byte_off, n, payload, payload_cap);
  g_launches++;
}

}  // namespace dc
