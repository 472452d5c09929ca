// sim.cpp -- host-compiled copy of the kernels' per-thread bodies, for logic tests without a GPU.
// TEST HARNESS ONLY (see host_shim.h).  Mirrors dc_stage_run_host() for one stream and the batch DEGA
// encode/decode for uniform layouts, using plain host memory instead of device buffers.
#include "host_shim.h"

#include <algorithm>
#include <vector>

#include "../../data_compressor_b200/csrc/dega_device.cuh"
#include "../../data_compressor_b200/csrc/lzmh_device.cuh"
#include "../../data_compressor_b200/csrc/csv_device.cuh"
#include "../../include/dc_cuda.h"

using namespace dc;


static const DivEntry *div_table()
{
  static std::vector<DivEntry> tab;
  if (tab.empty())
  {
    tab.resize(kDivTableEntries);
    build_div_table(tab.data());
  }
  return tab.data();
}

static uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

extern "C" {

// `decode csv`'s conversion of one field (csv_to_float); *hard = 1 when the field is left to strtof
float sim_csv_field(const uint8_t *text, uint64_t n, int *hard)
{
  bool h = false;
  const float f = dc::csv_to_float(text, 0, n, h);
  *hard = h ? 1 : 0;
  return f;
}

// the two passes of `encode csv` (csv_measure_kernel / csv_write_kernel) over n values, sequentially; returns the text length
uint64_t sim_csv_encode(const uint32_t *values, uint64_t n, uint32_t column, char sep, uint32_t decimals, uint8_t *text, uint64_t cap)
{
  uint64_t pos = 0;
  for (uint64_t j = 0; j < n; j++)
  {
    const uint64_t len = (uint64_t)(column - 1) + dc::csv_print<false>(values[j], decimals, nullptr);
    if (pos + len <= cap)
    {
      for (uint32_t k = 0; k + 1 < column; k++)
        text[pos + k] = (uint8_t)sep;
      if (dc::csv_print<true>(values[j], decimals, text + pos + column - 1) + (column - 1) != len)
        return ~0ull;
    }
    pos += len;
  }
  return pos;
}

// The two forms of the encoder's interval update side by side: the top-aligned (a0, LO) pair of narrow_fast() and the one-register
// form of narrow_w(), on random windows, increments and ranges (pending carries included: lo + dlo may leave the window).  Returns the
// number of disagreements in the doublings, the new range, the new window or the outgoing bits.
uint64_t sim_check_narrow_w(uint64_t trials, uint64_t seed)
{
  uint64_t bad = 0, x = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&]() {
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    return x;
  };
  for (uint64_t t = 0; t < trials; t++)
  {
    const uint32_t lo16 = (uint32_t)rnd() & 0xFFFFu;
    uint32_t r1 = 1 + (uint32_t)(rnd() % 65536);
    uint32_t dlo = (uint32_t)(rnd() % 65536);
    if (t % 4 == 1)
      r1 = 1 + (uint32_t)(rnd() % 8);          // narrow intervals: many doublings
    if (t % 4 == 2)
      dlo = 0x8000u - (lo16 & 0x7FFFu) - (uint32_t)(rnd() % 4);   // around the midpoint: E3 runs
    dc::BacEncoder a, b;
    a.init();
    b.init();
    a.LO = lo16 << 16;
    uint32_t kg1 = 0;
    a.narrow_fast(dlo & 0xFFFFu, r1, kg1);
    const uint32_t n = kg1 - 1;
    uint32_t W = dc::BacEncoder::to_w(lo16 << 16), qs = 0;
    b.narrow_w(W, dlo & 0xFFFFu, r1, qs);
    const uint32_t n_w = 15u - qs;
    bool ok = n == n_w && a.rng == b.rng && a.LO == dc::BacEncoder::from_w(W);
    if (ok && n <= 14)
      ok = a.a0 == (W >> 17);
    bad += ok ? 0 : 1;
  }
  return bad;
}

// DenseReader on its own: the stream of `nbits` bits that starts `byte_off` bytes into buf (an allocation of buf_bytes,
// a multiple of 4); after every step the 32 bits at the current position are recorded, then the position moves on by steps[i]
// (<= 32).  Returns the position (bits consumed, counted past the end too).
uint32_t sim_dense_reader(const uint8_t *buf, uint64_t buf_bytes, uint32_t byte_off, uint32_t nbits, const uint32_t *steps, uint32_t n_steps,
                          uint32_t *peeks)
{
  std::vector<uint8_t> copy(buf, buf + buf_bytes);   // the reader reads whole aligned words: keep them inside a known allocation
  dc::DenseReader r;
  r.init(copy.data() + byte_off, nbits);
  for (uint32_t i = 0; i < n_steps; i++)
  {
    peeks[i] = r.peek32();
    r.advance(steps[i]);
  }
  return r.pos;
}

// CarryWriter on its own: pushes (v[i], k[i]) (k <= 31, bit k of v = a carry into what was appended before), then finish();
// returns the number of bits, the words land in out (cap_bytes).
uint64_t sim_carry_writer(const uint32_t *v, const uint32_t *k, uint32_t n, uint8_t *out, uint64_t cap_bytes, int *overflow)
{
  dc::CarryWriter w;
  w.init(out, cap_bytes);
  for (uint32_t i = 0; i < n; i++)
    w.push(v[i], k[i]);
  const uint64_t bits = w.finish();
  *overflow = w.overflow ? 1 : 0;
  return bits;
}

// how often the arithmetic coder's unrolled groups ran out of room and were redone (which = 0 encoder, 1 decoder)
uint64_t sim_bac_redo_count(int which) { return g_bac_redo[which & 1]; }

// exhaustive-ish check of the division table: returns the number of mismatches
uint64_t sim_check_div_table(uint32_t stride)
{
  const DivEntry *t = div_table();
  uint64_t bad = 0;
  for (uint32_t d = 3; d < (uint32_t)kDivTableEntries; d++)
  {
    for (uint64_t x = 0; x < (1ull << 30); x += stride)
      bad += div_magic((uint32_t)x, t[d], div_shift(d)) != (uint32_t)(x / d);
    const uint64_t edge[] = { (1ull << 30) - 1, (1ull << 29), (1ull << 29) - 1, 65536ull * 8192, 65535ull * 8191 };
    for (uint64_t x : edge)
      bad += div_magic((uint32_t)x, t[d], div_shift(d)) != (uint32_t)(x / d);
    for (uint64_t k = 1; k * d < (1ull << 30); k += 1 + stride / 64)   // multiples of d and their predecessors
    {
      bad += div_magic((uint32_t)(k * d), t[d], div_shift(d)) != (uint32_t)k;
      bad += div_magic((uint32_t)(k * d - 1), t[d], div_shift(d)) != (uint32_t)(k - 1);
    }
    bad += div_shift(d) != div_shift_host(d);
    for (uint32_t x = 0; x <= 65536; x++)   // the shift-free form, every value the range can take (bac.c:109)
      bad += div_small(x, t[d]) != x / d;
  }
  return bad;
}

// Same contract as dc_stage_run_host (include/dc_cuda.h), executed by the host-compiled device bodies.
int sim_stage_run(int stage, const uint8_t *in, uint64_t in_bits, uint32_t vs, float nf, int adaptive, uint8_t *out, uint64_t out_cap,
                  uint64_t *out_bits)
{
  *out_bits = 0;
  const uint64_t in_bytes = (in_bits + 7) / 8;
  const bool float_in = stage == DC_STAGE_NORMALIZE || stage == DC_STAGE_DEGA_ENC;
  const bool float_out = stage == DC_STAGE_DENORMALIZE || stage == DC_STAGE_DEGA_DEC;
  const uint64_t n_in = float_in ? in_bits / 32 : in_bits / vs;

  std::vector<uint8_t> h_in(round_up(in_bytes, 16) + 4096, 0);
  if (in_bytes)
    memcpy(h_in.data(), in, in_bytes);
  uint64_t meta[8] = { in_bits, 0, 0, 0, 0, 0, 0, 0 };
  int32_t *status = reinterpret_cast<int32_t *>(&meta[4]);

  Layout L;
  L.n_streams = 1;
  L.channel_len = n_in;
  L.block_len = n_in ? n_in : 1;
  L.blocks_per_channel = 1;
  L.sample_off = nullptr;

  BitSrc src_in = { h_in.data(), nullptr, 0, &meta[0] };
  for (unsigned boost = 1;; boost *= 8)
  {
    uint64_t mid_cap = 0, out_dev = 0;
    const uint64_t seg_cap = round_up((n_in * (2ull * vs + 1) + 7) / 8 + 8, 16);
    switch (stage)
    {
      case DC_STAGE_NORMALIZE:
      case DC_STAGE_DIFF_ENC:
      case DC_STAGE_DIFF_DEC: out_dev = round_up((std::max(n_in, in_bits / vs + 1) * vs + 7) / 8 + 8, 16); break;
      case DC_STAGE_SEG_ENC: out_dev = seg_cap; break;
      case DC_STAGE_SEG_DEC: out_dev = round_up(((in_bits + 1) * vs + 7) / 8 + 8, 16); break;
      case DC_STAGE_DENORMALIZE: out_dev = round_up((in_bits / vs + 1) * 4, 16); break;
      case DC_STAGE_BAC_ENC: out_dev = round_up((in_bytes + 24) * 13 / 8 * boost + 64, 16); break;
      case DC_STAGE_BAC_DEC: out_dev = round_up((in_bytes * 4 + 4096) * boost, 16); break;
      case DC_STAGE_DEGA_ENC: mid_cap = seg_cap; out_dev = round_up(seg_cap * 13 / 8 * boost + 64, 16); break;
      case DC_STAGE_DEGA_DEC: mid_cap = round_up((in_bytes * 4 + 4096) * boost, 16); out_dev = round_up(mid_cap * 8 * 4 + 16, 16); break;
      default: return DC_ERR_INVALID_VALUE;
    }
    std::vector<uint8_t> h_mid(mid_cap + 4096, 0), h_out(out_dev + 16, 0);   // the back end reads one word ahead
    memset(meta + 1, 0, 7 * sizeof(uint64_t));

    const bool is_front = stage == DC_STAGE_NORMALIZE || stage == DC_STAGE_DIFF_ENC || stage == DC_STAGE_SEG_ENC || stage == DC_STAGE_DEGA_ENC;
    const bool is_back = stage == DC_STAGE_DENORMALIZE || stage == DC_STAGE_DIFF_DEC || stage == DC_STAGE_SEG_DEC || stage == DC_STAGE_DEGA_DEC;
    if (is_front)
    {
      FrontArgs fa;
      memset(&fa, 0, sizeof(fa));
      fa.lay = L;
      fa.samples = reinterpret_cast<const float *>(h_in.data());
      fa.in = src_in;
      fa.in_is_float = float_in;
      fa.do_norm = float_in;
      fa.do_diff = stage == DC_STAGE_DIFF_ENC || stage == DC_STAGE_DEGA_ENC;
      fa.out_seg = stage == DC_STAGE_SEG_ENC || stage == DC_STAGE_DEGA_ENC;
      fa.vs = vs;
      fa.nf = nf;
      fa.out.base = stage == DC_STAGE_DEGA_ENC ? h_mid.data() : h_out.data();
      fa.out.stride_bytes = stage == DC_STAGE_DEGA_ENC ? mid_cap : out_dev;
      fa.out.nbits = stage == DC_STAGE_DEGA_ENC ? &meta[1] : &meta[2];
      fa.status = status;
      if (fa.in_is_float && fa.do_norm && fa.do_diff && fa.out_seg)
        front_dega_body(fa, 0);
      else
        front_body(fa, 0);
    }
    if (stage == DC_STAGE_BAC_ENC || stage == DC_STAGE_DEGA_ENC || stage == DC_STAGE_BAC_DEC || stage == DC_STAGE_DEGA_DEC)
    {
      const bool enc = stage == DC_STAGE_BAC_ENC || stage == DC_STAGE_DEGA_ENC;
      BacArgs ba;
      memset(&ba, 0, sizeof(ba));
      ba.n_streams = 1;
      if (stage == DC_STAGE_DEGA_ENC)
        ba.in = BitSrc{ h_mid.data(), nullptr, mid_cap, &meta[1] };
      else
        ba.in = src_in;
      const bool to_mid = stage == DC_STAGE_DEGA_DEC;
      ba.out.base = to_mid ? h_mid.data() : h_out.data();
      ba.out.stride_bytes = to_mid ? mid_cap : out_dev;
      ba.out.nbits = to_mid ? &meta[1] : &meta[2];
      ba.adaptive = adaptive;
      ba.div_table = div_table();
      ba.status = status;
      ba.overflow_flag = reinterpret_cast<unsigned long long *>(&meta[5]);
      if (enc)
      {
        if (adaptive)
          bac_encode_body<true>(ba, 0, div_table());
        else
          bac_encode_body<false>(ba, 0, div_table());
      }
      else
      {
        if (adaptive)
          bac_decode_body<true>(ba, 0, div_table());
        else
          bac_decode_body<false>(ba, 0, div_table());
      }
    }
    if (is_back)
    {
      BackArgs ka;
      memset(&ka, 0, sizeof(ka));
      Layout LO = L;
      const uint64_t cap_samples = stage == DC_STAGE_DEGA_DEC ? mid_cap * 8 : in_bits / vs + 1;
      LO.channel_len = cap_samples;
      LO.block_len = cap_samples;
      ka.lay = LO;
      ka.n_streams = 1;
      if (stage == DC_STAGE_DEGA_DEC)
        ka.in = BitSrc{ h_mid.data(), nullptr, mid_cap, &meta[1] };
      else
        ka.in = src_in;
      ka.in_seg = stage == DC_STAGE_SEG_DEC || stage == DC_STAGE_DEGA_DEC;
      ka.do_undiff = stage == DC_STAGE_DIFF_DEC || stage == DC_STAGE_DEGA_DEC;
      ka.do_denorm = float_out;
      ka.vs = vs;
      ka.nf = nf;
      ka.samples = reinterpret_cast<float *>(h_out.data());
      ka.out.base = h_out.data();
      ka.out.stride_bytes = out_dev;
      ka.out.nbits = &meta[2];
      ka.n_out = &meta[3];
      ka.status = status;
      if (ka.in_seg && ka.do_undiff && ka.do_denorm && ka.vs >= 17)
      {
        float stage4[kBackStage];
        if (ka.vs <= 32)
          back_dega_body<1>(ka, 0, stage4);
        else
          back_dega_body<0>(ka, 0, stage4);
      }
      else
        back_body(ka, 0);
    }
    if (meta[5] != 0 && boost < 4096)
      continue;
    if (*status != 0)
      return *status;
    const uint64_t bits = float_out ? meta[3] * 32 : meta[2];
    if (float_in && (in_bits % 32) != 0)
      return DC_ERR_LIBRARY_CALL;
    *out_bits = bits;
    const uint64_t nbytes = (bits + 7) / 8;
    if (nbytes > out_cap)
      return DC_ERR_MEMORY;
    if (nbytes)
      memcpy(out, h_out.data(), nbytes);
    return DC_OK;
  }
}

// Batch DEGA encode over a uniform layout on host memory; writes per-stream payloads back to back (dense) the way
// scan + gather do.  Returns 0 or the first negative code; status/bit_len/byte_off as in dc_dega_encode_device.
int sim_dega_encode_batch(const float *samples, uint64_t n_channels, uint64_t channel_len, uint64_t block_len, uint32_t vs, float nf,
                          int adaptive, uint8_t *payload, uint64_t payload_cap, uint64_t *bit_len, uint64_t *byte_off, int32_t *status,
                          uint64_t *total_bytes)
{
  Layout L;
  const uint64_t bl = block_len == 0 ? std::max<uint64_t>(channel_len, 1) : block_len;
  L.blocks_per_channel = channel_len == 0 ? 1 : (channel_len + bl - 1) / bl;
  L.n_streams = n_channels * L.blocks_per_channel;
  L.channel_len = channel_len;
  L.block_len = bl;
  L.sample_off = nullptr;
  const uint64_t n = L.n_streams;
  const uint64_t max_len = std::min(bl, channel_len);
  const uint64_t seg_cap = round_up((max_len * (2ull * vs + 1) + 7) / 8 + 8, 16);
  const uint64_t bac_cap = round_up(seg_cap * 13 / 8 + 64, 16);
  std::vector<uint8_t> seg(round_up(n, 32) * seg_cap + 4096), bac(n * bac_cap + 16);   // SEG scratch: interleaved per warp
  std::vector<uint64_t> seg_bits(n);
  unsigned long long flag = 0;
  for (uint64_t s = 0; s < n; s++)
    status[s] = 0;
  FrontArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.lay = L;
  fa.samples = samples;
  fa.in_is_float = fa.do_norm = fa.do_diff = fa.out_seg = 1;
  fa.vs = vs;
  fa.nf = nf;
  fa.out = BitDst{ seg.data(), seg_cap, seg_bits.data(), 1 };
  fa.status = status;
  BacArgs ba;
  memset(&ba, 0, sizeof(ba));
  ba.n_streams = n;
  ba.in = BitSrc{ seg.data(), nullptr, seg_cap, seg_bits.data(), 1 };
  ba.out = BitDst{ bac.data(), bac_cap, bit_len, 0 };
  ba.adaptive = adaptive;
  ba.div_table = div_table();
  ba.status = status;
  ba.overflow_flag = &flag;
  uint64_t run = 0;
  for (uint64_t s = 0; s < n; s++)
  {
    front_dega_body(fa, s);
    if (adaptive)
      bac_encode_body<true>(ba, s, div_table());
    else
      bac_encode_body<false>(ba, s, div_table());
    byte_off[s] = run;
    const uint64_t nb = (bit_len[s] + 7) / 8;
    if (run + nb <= payload_cap)
      memcpy(payload + run, bac.data() + s * bac_cap, nb);
    run += nb;
  }
  byte_off[n] = run;
  *total_bytes = run;
  if (flag)
    return DC_ERR_MEMORY;
  return run > payload_cap ? DC_ERR_MEMORY : DC_OK;
}

int sim_dega_decode_batch(const uint8_t *payload, const uint64_t *bit_len, const uint64_t *byte_off, uint64_t n_channels,
                          uint64_t channel_len, uint64_t block_len, uint32_t vs, float nf, int adaptive, float *samples, uint64_t *n_out,
                          int32_t *status)
{
  Layout L;
  const uint64_t bl = block_len == 0 ? std::max<uint64_t>(channel_len, 1) : block_len;
  L.blocks_per_channel = channel_len == 0 ? 1 : (channel_len + bl - 1) / bl;
  L.n_streams = n_channels * L.blocks_per_channel;
  L.channel_len = channel_len;
  L.block_len = bl;
  L.sample_off = nullptr;
  const uint64_t n = L.n_streams;
  const uint64_t max_len = std::min(bl, channel_len);
  const uint64_t seg_cap = round_up((max_len * (2ull * vs + 1) + 7) / 8 + 8, 16);
  std::vector<uint8_t> seg(round_up(n, 32) * seg_cap + 4096);   // interleaved per warp
  std::vector<uint64_t> seg_bits(n);
  // the payload is copied into a padded buffer: the reader may touch the aligned word after the last byte
  std::vector<uint8_t> pay(byte_off[n] + 16, 0);
  memcpy(pay.data(), payload, byte_off[n]);
  unsigned long long flag = 0;
  for (uint64_t s = 0; s < n; s++)
    status[s] = 0;
  BacArgs ba;
  memset(&ba, 0, sizeof(ba));
  ba.n_streams = n;
  ba.in = BitSrc{ pay.data(), byte_off, 0, bit_len };
  ba.out = BitDst{ seg.data(), seg_cap, seg_bits.data(), 1 };
  ba.adaptive = adaptive;
  ba.div_table = div_table();
  ba.status = status;
  ba.overflow_flag = &flag;
  BackArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.lay = L;
  ka.n_streams = n;
  ka.in = BitSrc{ seg.data(), nullptr, seg_cap, seg_bits.data(), 1 };
  ka.in_seg = ka.do_undiff = ka.do_denorm = 1;
  ka.vs = vs;
  ka.nf = nf;
  ka.samples = samples;
  ka.n_out = n_out;
  ka.status = status;
  for (uint64_t s = 0; s < n; s++)
  {
    if (adaptive)
      bac_decode_body<true>(ba, s, div_table());
    else
      bac_decode_body<false>(ba, s, div_table());
    if (vs >= 17)
    {
      float stage4[kBackStage];
      if (vs == 32)
        back_dega_body<2>(ka, s, stage4);     // the decode pipeline's configuration: interleaved scratch, value size 32
      else if (vs <= 32)
        back_dega_body<1>(ka, s, stage4);
      else
        back_dega_body<0>(ka, s, stage4);
    }
    else
      back_body(ka, s);
  }
  return DC_OK;
}
}
