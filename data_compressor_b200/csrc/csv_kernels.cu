// csv_kernels.cu -- the csv codec on the device.  `decode csv` (DCLib/src/csv.c:13-44, ReadCSV): one column of a text
// file -> float32, described next; `encode csv` (csv.c:46-65, WriteCSV): float32 -> text, at the end of the file and in
// csv_device.cuh.
//
//
// ReadCSV walks the text character by character: fields end at the separator, at '\n' or at the end of the input,
// '\n' resets the column counter, and every time the wanted column ends its text goes through strtof (csv.c:22-35).
// Lines that do not have the column produce nothing.  Here:
//   1. csv_count_kernel / csv_lines_kernel: positions of the line ends (per-tile counts, a scan, a ranked scatter);
//   2. csv_parse_kernel: one thread per line finds the wanted field and converts it;
//   3. the values of the lines that have the column are compacted in line order (scan over the flags).
// Conversion: strtof is a correctly rounded decimal -> binary32 conversion.  A field of the form
// [space]* [+-]? digits [. digits]  whose digits, after dropping leading zeros and trailing fraction zeros, make an
// integer M < 2^24 with at most ten fraction digits k is converted as float(M) / 10^k: both operands are exact in
// binary32, so the IEEE division is the correctly rounded result, i.e. strtof's.  Integers without a fraction up to
// 19 digits go through the correctly rounded integer -> float conversion; a decimal exponent shifts k (M * 10^e with e <= 10
// is one correctly rounded multiplication).  What is left of strtof's grammar (hex floats, inf/nan, more digits, large
// exponents) is flagged and converted by the caller with strtof itself; meter readings never get there.
#include <cuda_runtime.h>
#include <stdint.h>

#include "csv_device.cuh"
#include "kernels.h"

namespace dc {

void count_launch(int n);

static const int kCsvTile = 4096;      // bytes per CTA
static const int kCsvThreads = 256;    // 16 bytes per thread

__global__ void __launch_bounds__(kCsvThreads) csv_count_kernel(const uint8_t *text, uint64_t n, uint64_t *tile_bits)
{
  __shared__ uint32_t total;
  if (threadIdx.x == 0)
    total = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kCsvTile + (uint64_t)threadIdx.x * 16;
  uint32_t cnt = 0;
  for (int k = 0; k < 16; k++)
    if (base + k < n && text[base + k] == '\n')
      cnt++;
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt)
    atomicAdd(&total, cnt);
  __syncthreads();
  if (threadIdx.x == 0)
    tile_bits[blockIdx.x] = (uint64_t)total * 8;   // launch_byte_scan sums ceil(bits / 8)
}

// line_end[r] = position of the r-th '\n'; tile_off = exclusive scan of the tile counts
__global__ void __launch_bounds__(kCsvThreads) csv_lines_kernel(const uint8_t *text, uint64_t n, const uint64_t *tile_off, uint64_t *line_end)
{
  __shared__ uint32_t warp_base[kCsvThreads / 32];
  const uint64_t base = (uint64_t)blockIdx.x * kCsvTile + (uint64_t)threadIdx.x * 16;
  uint32_t mask = 0;
  for (int k = 0; k < 16; k++)
    if (base + k < n && text[base + k] == '\n')
      mask |= 1u << k;
  const uint32_t cnt = __popc(mask);
  uint32_t inc = cnt;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d)
      inc += t;
  }
  if (lane == 31)
    warp_base[wid] = inc;
  __syncthreads();
  uint32_t before = 0;
  for (int w = 0; w < wid; w++)
    before += warp_base[w];
  uint64_t rank = tile_off[blockIdx.x] + before + inc - cnt;
  while (mask)
  {
    const int k = __ffs((int)mask) - 1;
    mask &= mask - 1;
    line_end[rank++] = base + (uint64_t)k;
  }
}

// value[j], flag8[j] (8 = the line has the column, 0 = it does not), and the indices of fields strtof has to convert
__global__ void __launch_bounds__(256) csv_parse_kernel(const uint8_t *text, uint64_t n, const uint64_t *line_end, uint64_t n_newlines,
                                                        uint64_t n_lines, uint32_t column, uint8_t sep, float *value, uint64_t *flag8,
                                                        uint64_t *field_lo, uint64_t *field_hi, unsigned long long *n_hard,
                                                        uint64_t *hard_list, uint64_t hard_cap)
{
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_lines; j += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint64_t start = j == 0 ? 0 : line_end[j - 1] + 1;
    const uint64_t end = j < n_newlines ? line_end[j] : n;   // position of the '\n', or the end of the input
    // the wanted field: csv.c:22-40, columns are counted from 1
    uint64_t fs = start, fe = end;
    uint32_t col = 1;
    bool have = false;
    for (uint64_t i = start; i <= end; i++)
    {
      if (i == end || text[i] == sep)
      {
        if (col == column)
        {
          fe = i;
          // an EMPTY field that ends with the input (unterminated last line ending in a separator) is never seen by
          // the reference: its loop stops with the last character (csv.c:20)
          have = !(i == end && end == n && fs == end);
          break;
        }
        col++;
        fs = i + 1;
      }
    }
    flag8[j] = have ? 8 : 0;
    field_lo[j] = fs;
    field_hi[j] = fe;
    if (!have)
      continue;
    bool hard;
    value[j] = csv_to_float(text, fs, fe, hard);   // strtof on text[fs, fe)
    if (hard)
    {
      const unsigned long long h = atomicAdd(n_hard, 1ull);
      if (h < hard_cap)
        hard_list[h] = j;
    }
  }
}

__global__ void __launch_bounds__(256) csv_compact_kernel(const float *value, const uint64_t *flag8, const uint64_t *pos, uint64_t n_lines, float *out,
                                                          uint64_t out_cap)
{
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_lines; j += (uint64_t)gridDim.x * blockDim.x)
    if (flag8[j] && pos[j] < out_cap)
      out[pos[j]] = value[j];
}

static unsigned csv_grid(uint64_t n, unsigned block)
{
  const uint64_t want = (n + block - 1) / block;
  const uint64_t cap = (uint64_t)sm_count() * 16;
  return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

void launch_csv_count(const uint8_t *text, uint64_t n, uint64_t *tile_bits, cudaStream_t st)
{
  csv_count_kernel<<<(unsigned)((n + kCsvTile - 1) / kCsvTile), kCsvThreads, 0, st>>>(text, n, tile_bits);
  count_launch(1);
}

void launch_csv_lines(const uint8_t *text, uint64_t n, const uint64_t *tile_off, uint64_t *line_end, cudaStream_t st)
{
  csv_lines_kernel<<<(unsigned)((n + kCsvTile - 1) / kCsvTile), kCsvThreads, 0, st>>>(text, n, tile_off, line_end);
  count_launch(1);
}

void launch_csv_parse(const uint8_t *text, uint64_t n, const uint64_t *line_end, uint64_t n_newlines, uint64_t n_lines, uint32_t column, uint8_t sep,
                      float *value, uint64_t *flag8, uint64_t *field_lo, uint64_t *field_hi, unsigned long long *n_hard, uint64_t *hard_list,
                      uint64_t hard_cap, cudaStream_t st)
{
  csv_parse_kernel<<<csv_grid(n_lines, 256), 256, 0, st>>>(text, n, line_end, n_newlines, n_lines, column, sep, value, flag8, field_lo, field_hi, n_hard,
                                                            hard_list, hard_cap);
  count_launch(1);
}

// lo / hi / output position of the fields the host converts, packed in hard-list order (three bulk copies instead of
// three per field)
__global__ void __launch_bounds__(256) csv_hard_kernel(const uint64_t *hard_list, uint64_t n_hard, const uint64_t *field_lo, const uint64_t *field_hi,
                                                       const uint64_t *pos, uint64_t *packed)
{
  for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < n_hard; h += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint64_t j = hard_list[h];
    packed[h] = field_lo[j];
    packed[n_hard + h] = field_hi[j];
    packed[2 * n_hard + h] = pos[j];
  }
}

void launch_csv_hard(const uint64_t *hard_list, uint64_t n_hard, const uint64_t *field_lo, const uint64_t *field_hi, const uint64_t *pos,
                     uint64_t *packed, cudaStream_t st)
{
  if (n_hard == 0)
    return;
  csv_hard_kernel<<<csv_grid(n_hard, 256), 256, 0, st>>>(hard_list, n_hard, field_lo, field_hi, pos, packed);
  count_launch(1);
}

void launch_csv_compact(const float *value, const uint64_t *flag8, const uint64_t *pos, uint64_t n_lines, float *out, uint64_t out_cap, cudaStream_t st)
{
  csv_compact_kernel<<<csv_grid(n_lines, 256), 256, 0, st>>>(value, flag8, pos, n_lines, out, out_cap);
  count_launch(1);
}

// ---- `encode csv` (csv.c:46-65): float32 -> text ----------------------------------------------------------------
// One thread per value, two passes around a scan: the length of every line (prefix separators of the empty columns
// before the wanted one, csv.c:55-58, plus the printed value), then the text itself at the scanned byte position.
// Neighbouring threads write neighbouring byte ranges, so the byte stores of a warp merge into a few sectors in L2.
__global__ void __launch_bounds__(256) csv_measure_kernel(const uint32_t *values, uint64_t n, uint32_t decimals, uint32_t prefix, uint64_t *len_bits)
{
  // in bits: the scan shared with the coder (launch_byte_scan) turns bit counts into byte offsets
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
    len_bits[j] = 8 * ((uint64_t)prefix + csv_print<false>(values[j], decimals, nullptr));
}

__global__ void __launch_bounds__(256) csv_write_kernel(const uint32_t *values, uint64_t n, uint32_t decimals, uint32_t prefix, uint8_t sep,
                                                        const uint64_t *pos, uint8_t *text)
{
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
  {
    uint8_t *dst = text + pos[j];
    for (uint32_t k = 0; k < prefix; k++)
      dst[k] = sep;
    csv_print<true>(values[j], decimals, dst + prefix);
  }
}

void launch_csv_measure(const float *values, uint64_t n, uint32_t decimals, uint32_t prefix, uint64_t *len, cudaStream_t st)
{
  csv_measure_kernel<<<csv_grid(n, 256), 256, 0, st>>>(reinterpret_cast<const uint32_t *>(values), n, decimals, prefix, len);
  count_launch(1);
}

void launch_csv_write(const float *values, uint64_t n, uint32_t decimals, uint32_t prefix, uint8_t sep, const uint64_t *pos, uint8_t *text,
                      cudaStream_t st)
{
  csv_write_kernel<<<csv_grid(n, 256), 256, 0, st>>>(reinterpret_cast<const uint32_t *>(values), n, decimals, prefix, sep, pos, text);
  count_launch(1);
}

uint64_t csv_tiles(uint64_t n) { return (n + kCsvTile - 1) / kCsvTile; }

}  // namespace dc
