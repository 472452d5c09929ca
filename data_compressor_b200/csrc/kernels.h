// kernels.h -- host-callable launch wrappers of the sm_100a kernels (internal; the public ABI is include/dc_cuda.h)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bitio.cuh"

namespace dc {

// Stream geometry inside one flat sample array (see dc_layout in include/dc_cuda.h)
struct Layout
{
  uint64_t n_streams;
  uint64_t channel_len;
  uint64_t block_len;            // uniform mode: samples per block (>= 1)
  uint64_t blocks_per_channel;   // uniform mode
  const uint64_t *sample_off;    // ragged mode (device pointer) or nullptr
};

__host__ __device__ inline void stream_extent(const Layout &L, uint64_t s, uint64_t &off, uint64_t &len)
{
  if (L.sample_off)
  {
    off = L.sample_off[s];
    len = L.sample_off[s + 1] - off;
  }
  else
  {
    const uint64_t c = s / L.blocks_per_channel;
    const uint64_t b = s - c * L.blocks_per_channel;
    const uint64_t start = b * L.block_len;
    off = c * L.channel_len + start;
    len = L.channel_len - start < L.block_len ? L.channel_len - start : L.block_len;
  }
}

// Division table for the arithmetic coder: for 3 <= d < 16384, entry d = {m, mb} with
//   floor(x / d) == umulhi(x, m) >> div_shift(d)     for every x < 2^30      (rng * cum[1], bac.c:110-111)
//   floor(x / d) == umulhi(x, mb)                     for every x <= 2^16     (rng itself: no shift needed)
// Exact round-up reciprocals (Granlund-Montgomery): l = ceil(log2 d), m = ceil(2^(30+l) / d) < 2^31, shift l - 2;
// mb = ceil(2^32 / d): its error term x * (mb * d - 2^32) stays below 2^16 * 2^14 < 2^32.  The shift depends on d through
// l only, i.e. it changes where d passes a power of two, so the unrolled symbol groups keep it in a register.
// Entries 0..2 are unused (the coder's total frequency is at least 3).
struct DivEntry
{
  uint32_t m;
  uint32_t mb;
};
static const int kDivTableEntries = 16384;
inline uint32_t div_shift_host(uint32_t d)   // l - 2
{
  uint32_t l = 0;
  while ((1u << l) < d)
    l++;
  return l - 2;
}
inline void build_div_table(DivEntry *t)
{
  for (uint32_t d = 0; d < (uint32_t)kDivTableEntries; d++)
  {
    t[d].m = 0;
    t[d].mb = 0;
    if (d < 3)
      continue;
    const uint32_t l = div_shift_host(d) + 2;
    t[d].m = (uint32_t)((((uint64_t)1 << (30 + l)) + d - 1) / d);
    t[d].mb = (uint32_t)((((uint64_t)1 << 32) + d - 1) / d);
  }
}

// ---- front end: [normalize] -> [diff] -> {seg | fixed-width pack} ---------------------------------
struct FrontArgs
{
  Layout lay;
  const float *samples;     // float source (in_is_float)
  BitSrc in;                // vs-bit integer source (!in_is_float)
  int in_is_float;
  int do_norm;
  int do_diff;
  int out_seg;              // 1: signed Exp-Golomb, 0: vs-bit fields
  uint32_t vs;
  float nf;
  BitDst out;
  int32_t *status;
  unsigned int *work;       // device counter of the warp-task scheduler (specialised kernel; zeroed by the launcher)
};
void launch_front(const FrontArgs &a, cudaStream_t st);

// ---- A-XDR fast path: equal-length streams whose records are whole 32-bit words ------------------------------
// One thread per 32-bit output word (encode) / per sample (decode): fully coalesced, no scratch, no scan, no gather.
struct AxdrArgs
{
  uint64_t n_streams;
  uint64_t len;              // samples per stream
  uint64_t words_per_stream; // len * vs / 32
  uint32_t vs;
  float nf;
  const float *samples_in;   // encode
  float *samples_out;        // decode
  uint32_t *payload_words;   // encode: dense payload (4-byte aligned)
  const uint8_t *payload;    // decode
  uint64_t *bit_len;         // encode out
  uint64_t *byte_off;        // encode out [n+1]
  uint64_t *n_out;           // decode out
  int32_t *status;
  unsigned long long *err_flag;   // encode: set when any sample fails the range check (caller falls back)
};
void launch_axdr_encode(const AxdrArgs &a, cudaStream_t st);
void launch_axdr_decode(const AxdrArgs &a, cudaStream_t st);

// ---- back end: {seg parse | fixed-width unpack} -> [undiff] -> [denormalize] ------------------------
struct BackArgs
{
  Layout lay;               // destination geometry for float output (capacity per stream)
  uint64_t n_streams;
  BitSrc in;
  int in_seg;
  int do_undiff;
  int do_denorm;            // 1: float32 out into `samples`, 0: vs-bit fields into `out`
  uint32_t vs;
  float nf;
  float *samples;
  BitDst out;
  uint64_t *n_out;          // values produced per stream
  int32_t *status;
  unsigned int *work;       // device counter of the warp-task scheduler (specialised kernel; zeroed by the launcher)
};
void launch_back(const BackArgs &a, cudaStream_t st);

// ---- binary arithmetic coder -----------------------------------------------------------------------
struct BacArgs
{
  uint64_t n_streams;
  BitSrc in;
  BitDst out;
  int adaptive;
  const DivEntry *div_table;   // device copy
  int32_t *status;             // in/out: streams with status != 0 are skipped (out bits = 0)
  unsigned long long *overflow_flag;   // set to 1 when any stream outgrew its output stride (may be nullptr)
  unsigned int *work;          // device counter of the warp-task scheduler (zeroed by the launcher)
};
void launch_bac_encode(const BacArgs &a, cudaStream_t st);
void launch_bac_decode(const BacArgs &a, cudaStream_t st);
int bac_smem_bytes();
int bac_threads_per_sm();   // streams one SM holds in flight in the coder kernels

// ---- compaction: byte counts -> exclusive scan -> dense payload ---------------------------------------
// byte_off[n+1] = exclusive scan of ceil(nbits[s] / 8); tile_tmp needs ceil(n / 4096) + 1 uint64.
void launch_byte_scan(const uint64_t *nbits, uint64_t n, uint64_t *byte_off, uint64_t *tile_tmp, cudaStream_t st);
// copies stream s (scratch, fixed stride) to payload + byte_off[s]; streams that do not fit are skipped
void launch_gather(const uint8_t *scratch, uint64_t stride_bytes, const uint64_t *nbits, const uint64_t *byte_off,
                   uint64_t n, uint8_t *payload, uint64_t payload_cap, cudaStream_t st);

// marks streams whose byte range [byte_off[s], + ceil(bit_len[s] / 8)) leaves the payload with ERROR_INVALID_FORMAT
void launch_validate_index(const uint64_t *bit_len, const uint64_t *byte_off, uint64_t n, uint64_t payload_bytes, int32_t *status, cudaStream_t st);

// writes {*d_flag, *d_total} to pub[0..1] (mapped pinned host memory)
void launch_publish(const uint64_t *d_flag, const uint64_t *d_total, uint64_t *pub, cudaStream_t st);

// ---- aggregate (granularity reduction): out[c][g] = sum of in[c][g*k .. g*k+k), strictly sequential float32 adds
void launch_aggregate(const float *in, uint64_t n_channels, uint64_t channel_len, uint32_t num_values, float *out, cudaStream_t st);

// ---- csv (decode csv: one text column -> float32), see csv_kernels.cu
uint64_t csv_tiles(uint64_t n);
void launch_csv_count(const uint8_t *text, uint64_t n, uint64_t *tile_bits, cudaStream_t st);
void launch_csv_lines(const uint8_t *text, uint64_t n, const uint64_t *tile_off, uint64_t *line_end, cudaStream_t st);
void launch_csv_parse(const uint8_t *text, uint64_t n, const uint64_t *line_end, uint64_t n_newlines, uint64_t n_lines, uint32_t column, uint8_t sep,
                      float *value, uint64_t *flag8, uint64_t *field_lo, uint64_t *field_hi, unsigned long long *n_hard, uint64_t *hard_list,
                      uint64_t hard_cap, cudaStream_t st);
void launch_csv_measure(const float *values, uint64_t n, uint32_t decimals, uint32_t prefix, uint64_t *len, cudaStream_t st);
void launch_csv_write(const float *values, uint64_t n, uint32_t decimals, uint32_t prefix, uint8_t sep, const uint64_t *pos, uint8_t *text,
                      cudaStream_t st);
void launch_csv_hard(const uint64_t *hard_list, uint64_t n_hard, const uint64_t *field_lo, const uint64_t *field_hi, const uint64_t *pos,
                     uint64_t *packed, cudaStream_t st);
void launch_csv_compact(const float *value, const uint64_t *flag8, const uint64_t *pos, uint64_t n_lines, float *out, uint64_t out_cap, cudaStream_t st);

// ---- LZMH ------------------------------------------------------------------------------------------------
struct LzmhArgs
{
  uint64_t n_streams;
  const uint8_t *in;
  const uint64_t *in_off;      // [n+1] byte offsets (encode) / nullptr
  BitSrc bits_in;              // decode input
  BitDst bits_out;             // encode output
  uint8_t *out;                // decode output bytes
  const uint64_t *out_off;     // [n+1] capacity windows (decode)
  uint64_t *out_len;           // decode: bytes produced
  int32_t *status;
};
void launch_lzmh_encode(const LzmhArgs &a, cudaStream_t st);
void launch_lzmh_decode(const LzmhArgs &a, cudaStream_t st);

// multiprocessor count of the current device
int sm_count();

// number of kernel launches issued so far by the launch_* wrappers (process-wide)
uint64_t launch_counter();

}  // namespace dc
