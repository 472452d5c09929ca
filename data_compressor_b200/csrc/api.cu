// api.cu -- host side of the C ABI declared in include/dc_cuda.h.
//
// No CPU fallback anywhere in this file: every codec result is produced by the kernels in dega_kernels.cu /
// lzmh_kernels.cu.  Without a usable GPU dc_create() fails with DC_ERR_LIBRARY_INIT.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dc_cuda.h"
#include "kernels.h"

using namespace dc;

static const int kMaxHostChunks = 32;

struct dc_ctx
{
  int device;
  cudaStream_t stream;
  DivEntry *d_div;
  unsigned int *d_work;  // work counter of the coder kernels' warp-task scheduler
  uint8_t *arena;        // kernel scratch (SEG / coder strides), grown on demand
  uint64_t arena_size;
  uint8_t *stage;        // device staging of the host-buffer calls (samples, payload, index), grown on demand
  uint64_t stage_size;
  uint8_t *chain[2];     // intermediate results of dc_chain_run_host (ping-pong), grown on demand
  uint64_t chain_size[2];
  cudaStream_t s_h2d, s_d2h;          // copy engines of the host-buffer calls (overlap with the compute stream)
  cudaEvent_t ev_in[kMaxHostChunks], ev_out[kMaxHostChunks];    // per-chunk "input landed" / "results ready"
  uint64_t *h_pinned;   // small pinned mailbox for device->host scalars
  uint64_t launches_base;
  int profiling;
  cudaEvent_t ev[2 * DC_K_COUNT];
  uint32_t ev_mask;   // kernels timed during the last call
  std::thread *worker;   // the *_host_async call in flight (at most one per context), joined by dc_wait
  int worker_rc;
  char err[256];
};

namespace {

void set_err(dc_ctx *c, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(c->err, sizeof(c->err), fmt, ap);
  va_end(ap);
}

#define CU_TRY(ctx, call, code)                                                                          \
  do                                                                                                      \
  {                                                                                                       \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess)                                                                               \
    {                                                                                                     \
      set_err((ctx), "%s failed: %s", #call, cudaGetErrorString(e__));                                    \
      return (code);                                                                                      \
    }                                                                                                     \
  } while (0)

uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// RAII: CUDA events around one kernel group when profiling is on
struct Timed
{
  dc_ctx *c;
  int id;
  Timed(dc_ctx *ctx, int kernel_id) : c(ctx), id(kernel_id)
  {
    if (c->profiling)
      cudaEventRecord(c->ev[2 * id], c->stream);
  }
  ~Timed()
  {
    if (c->profiling)
    {
      cudaEventRecord(c->ev[2 * id + 1], c->stream);
      c->ev_mask |= 1u << id;
    }
  }
};

// Bump allocator over one device block; grown (with a synchronise) when a call needs more.
int arena_reserve(dc_ctx *c, uint64_t bytes)
{
  if (bytes <= c->arena_size)
    return DC_OK;
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  if (c->arena)
    cudaFree(c->arena);
  c->arena = nullptr;
  c->arena_size = 0;
  const uint64_t want = round_up(bytes + bytes / 8, 1 << 20);
  if (cudaMalloc(&c->arena, want) != cudaSuccess)
  {
    cudaGetLastError();
    if (cudaMalloc(&c->arena, round_up(bytes, 1 << 20)) != cudaSuccess)
    {
      cudaGetLastError();
      set_err(c, "device scratch allocation of %llu bytes failed", (unsigned long long)bytes);
      return DC_ERR_MEMORY;
    }
    c->arena_size = round_up(bytes, 1 << 20);
    return DC_OK;
  }
  c->arena_size = want;
  return DC_OK;
}

// same policy for the staging block of the *_host calls
int stage_reserve(dc_ctx *c, uint64_t bytes)
{
  if (bytes <= c->stage_size)
    return DC_OK;
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  if (c->stage)
    cudaFree(c->stage);
  c->stage = nullptr;
  c->stage_size = 0;
  const uint64_t want = round_up(bytes, 1 << 20);
  if (cudaMalloc(&c->stage, want) != cudaSuccess)
  {
    cudaGetLastError();
    set_err(c, "device staging allocation of %llu bytes failed", (unsigned long long)bytes);
    return DC_ERR_MEMORY;
  }
  c->stage_size = want;
  return DC_OK;
}

struct Bump
{
  uint8_t *base;
  uint64_t used;
  explicit Bump(uint8_t *b) : base(b), used(0) {}
  template <typename T>
  T *take(uint64_t count)
  {
    T *p = reinterpret_cast<T *>(base + used);
    used += round_up(count * sizeof(T), 256);
    return p;
  }
  static uint64_t size_of(uint64_t count, uint64_t elem) { return round_up(count * elem, 256); }
};

bool params_ok(const dc_params *p) { return p && p->value_size_bits >= 1 && p->value_size_bits <= 64; }

// dc_layout -> device Layout; max_len = longest stream (needs sample_off on the host for ragged mode)
int make_layout(dc_ctx *c, const dc_layout *in, const uint64_t *host_off, Layout &L, uint64_t &max_len, uint64_t &total_samples)
{
  if (!in)
    return DC_ERR_INVALID_VALUE;
  if (in->sample_off == nullptr)
  {
    const uint64_t bl = in->block_len == 0 ? std::max<uint64_t>(in->channel_len, 1) : in->block_len;
    const uint64_t bpc = in->channel_len == 0 ? 1 : (in->channel_len + bl - 1) / bl;
    L.n_streams = in->n_channels * bpc;
    L.channel_len = in->channel_len;
    L.block_len = bl;
    L.blocks_per_channel = bpc;
    L.sample_off = nullptr;
    max_len = std::min(bl, in->channel_len);
    total_samples = in->n_channels * in->channel_len;
    return DC_OK;
  }
  L.n_streams = in->n_streams;
  L.channel_len = 0;
  L.block_len = 0;
  L.blocks_per_channel = 1;
  L.sample_off = in->sample_off;
  max_len = 0;
  total_samples = 0;
  if (host_off)
  {
    for (uint64_t s = 0; s < in->n_streams; s++)
    {
      if (host_off[s + 1] < host_off[s])
      {
        set_err(c, "sample_off is not monotonic at stream %llu", (unsigned long long)s);
        return DC_ERR_INVALID_VALUE;
      }
      max_len = std::max(max_len, host_off[s + 1] - host_off[s]);
    }
    total_samples = in->n_streams ? host_off[in->n_streams] : 0;
  }
  return DC_OK;
}

// ragged device layouts: fetch the offsets once to size the scratch (not the performance path)
int fetch_offsets(dc_ctx *c, const dc_layout *in, std::vector<uint64_t> &host_off)
{
  host_off.resize(in->n_streams + 1);
  CU_TRY(c, cudaMemcpyAsync(host_off.data(), in->sample_off, (in->n_streams + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream),
         DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  return DC_OK;
}

uint64_t seg_cap_bytes(uint64_t max_len, uint32_t vs) { return round_up((max_len * (2ull * vs + 1) + 7) / 8 + 8, 16); }
uint64_t fixed_cap_bytes(uint64_t max_len, uint32_t vs) { return round_up((max_len * (uint64_t)vs + 7) / 8 + 8, 16); }
uint64_t bac_cap_bytes(uint64_t in_cap_bytes, int adaptive, unsigned boost)
{
  // adaptive frequency counting costs at most ~1 bit per input bit plus O(log n); the static 1/3-1/3-1/3
  // model costs log2(3) bits per input bit.  `boost` multiplies the estimate after a reported overflow.
  const uint64_t est = adaptive ? in_cap_bytes + in_cap_bytes / 8 : in_cap_bytes * 13 / 8;
  return round_up(est * boost + 64, 16);
}

int check_launch(dc_ctx *c, const char *what)
{
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
  {
    set_err(c, "%s: %s", what, cudaGetErrorString(e));
    return DC_ERR_LIBRARY_CALL;
  }
  return DC_OK;
}

// Reads `count` uint64 from the device through the pinned mailbox (synchronises the stream).
int read_back(dc_ctx *c, const uint64_t *d_src, uint64_t *dst, int count)
{
  CU_TRY(c, cudaMemcpyAsync(c->h_pinned, d_src, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
  {
    set_err(c, "kernel execution failed: %s", cudaGetErrorString(e));
    return DC_ERR_LIBRARY_CALL;
  }
  for (int i = 0; i < count; i++)
    dst[i] = c->h_pinned[i];
  return DC_OK;
}

// ---- A-XDR fast path: equal-length streams, records that are whole 32-bit words -------------------------------
bool axdr_eligible(const Layout &L, uint32_t vs, uint64_t &len)
{
  if (L.sample_off != nullptr || L.n_streams == 0 || L.channel_len == 0)
    return false;
  if (L.channel_len % L.block_len != 0 && L.blocks_per_channel != 1)
    return false;
  len = L.blocks_per_channel == 1 ? L.channel_len : L.block_len;
  return (len * vs) % 32 == 0;
}

int axdr_fast_encode(dc_ctx *c, const float *d_samples, const Layout &L, const dc_params *p, uint8_t *d_payload, uint64_t payload_cap,
                     uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status, uint64_t *total_bytes)
{
  uint64_t len;
  if (!axdr_eligible(L, p->value_size_bits, len) || (reinterpret_cast<uintptr_t>(d_payload) & 15) != 0)
    return DC_ERR_INVALID_MODE;
  const uint64_t wps = len * p->value_size_bits / 32;
  const uint64_t tot = L.n_streams * wps * 4;
  if (total_bytes)
    *total_bytes = tot;
  if (tot > payload_cap)
  {
    set_err(c, "payload needs %llu bytes, capacity is %llu", (unsigned long long)tot, (unsigned long long)payload_cap);
    return DC_ERR_MEMORY;
  }
  int rc = arena_reserve(c, 4096);
  if (rc != DC_OK)
    return rc;
  c->ev_mask = 0;
  unsigned long long *d_flag = reinterpret_cast<unsigned long long *>(c->arena);
  CU_TRY(c, cudaMemsetAsync(d_flag, 0, sizeof(unsigned long long), c->stream), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaMemsetAsync(d_status, 0, L.n_streams * sizeof(int32_t), c->stream), DC_ERR_LIBRARY_CALL);
  AxdrArgs a;
  memset(&a, 0, sizeof(a));
  a.n_streams = L.n_streams;
  a.len = len;
  a.words_per_stream = wps;
  a.vs = p->value_size_bits;
  a.nf = p->normalization_factor;
  a.samples_in = d_samples;
  a.payload_words = reinterpret_cast<uint32_t *>(d_payload);
  a.bit_len = d_bit_len;
  a.byte_off = d_byte_off;
  a.status = d_status;
  a.err_flag = d_flag;
  {
    Timed t(c, DC_K_FRONT);
    launch_axdr_encode(a, c->stream);
  }
  if ((rc = check_launch(c, "axdr_encode_kernel")) != DC_OK)
    return rc;
  uint64_t flag = 0;
  if ((rc = read_back(c, reinterpret_cast<const uint64_t *>(d_flag), &flag, 1)) != DC_OK)
    return rc;
  return flag ? DC_ERR_INVALID_MODE : DC_OK;   // a failed stream changes the dense layout: redo on the general path
}

int axdr_fast_decode(dc_ctx *c, const uint8_t *d_payload, uint64_t payload_bytes, const Layout &L, const dc_params *p, float *d_samples,
                     uint64_t *d_n_out, int32_t *d_status)
{
  uint64_t len;
  if (!axdr_eligible(L, p->value_size_bits, len) || (reinterpret_cast<uintptr_t>(d_payload) & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(d_samples) & 15) != 0)
    return DC_ERR_INVALID_MODE;
  const uint64_t wps = len * p->value_size_bits / 32;
  if (payload_bytes != L.n_streams * wps * 4)   // not the dense equal-length layout (e.g. failed streams): general path
    return DC_ERR_INVALID_MODE;
  c->ev_mask = 0;
  AxdrArgs a;
  memset(&a, 0, sizeof(a));
  a.n_streams = L.n_streams;
  a.len = len;
  a.words_per_stream = wps;
  a.vs = p->value_size_bits;
  a.nf = p->normalization_factor;
  a.samples_out = d_samples;
  a.payload = d_payload;
  a.n_out = d_n_out;
  a.status = d_status;
  {
    Timed t(c, DC_K_BACK);
    launch_axdr_decode(a, c->stream);
  }
  return check_launch(c, "axdr_decode_kernel");
}

// ---- encode pipeline: front end -> [bac] -> scan -> gather ----------------------------------------------
enum FrontKind { FRONT_DEGA, FRONT_AXDR };

int encode_pipeline(dc_ctx *c, FrontKind kind, const float *d_samples, const Layout &L, uint64_t max_len, const dc_params *p,
                    uint8_t *d_payload, uint64_t payload_cap, uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status,
                    uint64_t *total_bytes)
{
  const uint64_t n = L.n_streams;
  const uint32_t vs = p->value_size_bits;
  c->ev_mask = 0;
  for (unsigned boost = 1;; boost *= 8)
  {
    const uint64_t front_cap = kind == FRONT_DEGA ? seg_cap_bytes(max_len, vs) : fixed_cap_bytes(max_len, vs);
    const uint64_t bac_cap = kind == FRONT_DEGA ? bac_cap_bytes(front_cap, p->adaptive, boost) : 0;
    const uint64_t n_tiles = (n + 4095) / 4096;
    // SEG scratch between front end and coder: word-interleaved per warp (bitio.cuh), so it covers whole warps
    const uint64_t n_front = kind == FRONT_DEGA ? round_up(n, 32) : n;
    const uint64_t need = Bump::size_of(n, 8) + Bump::size_of(n_tiles + 2, 8) + Bump::size_of(2, 8) + Bump::size_of(n_front * front_cap, 1) +
                          Bump::size_of(n * bac_cap, 1) + 4096;
    int rc = arena_reserve(c, need);
    if (rc != DC_OK)
      return rc;
    Bump b(c->arena);
    uint64_t *d_front_bits = b.take<uint64_t>(n);
    uint64_t *d_tiles = b.take<uint64_t>(n_tiles + 2);
    uint64_t *d_flags = b.take<uint64_t>(2);
    uint8_t *d_front = b.take<uint8_t>(n_front * front_cap);
    uint8_t *d_bac = b.take<uint8_t>(n * bac_cap);

    CU_TRY(c, cudaMemsetAsync(d_status, 0, n * sizeof(int32_t), c->stream), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint64_t), c->stream), DC_ERR_LIBRARY_CALL);

    FrontArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.lay = L;
    fa.samples = d_samples;
    fa.in_is_float = 1;
    fa.do_norm = 1;
    fa.do_diff = kind == FRONT_DEGA;
    fa.out_seg = kind == FRONT_DEGA;
    fa.vs = vs;
    fa.nf = p->normalization_factor;
    fa.out.base = d_front;
    fa.out.stride_bytes = front_cap;
    fa.out.nbits = kind == FRONT_DEGA ? d_front_bits : d_bit_len;
    fa.out.interleave = kind == FRONT_DEGA;
    fa.status = d_status;
    fa.work = c->d_work;
    {
      Timed t(c, DC_K_FRONT);
      launch_front(fa, c->stream);
    }
    if ((rc = check_launch(c, "front_kernel")) != DC_OK)
      return rc;

    const uint8_t *d_final = d_front;
    uint64_t final_stride = front_cap;
    if (kind == FRONT_DEGA)
    {
      BacArgs ba;
      memset(&ba, 0, sizeof(ba));
      ba.n_streams = n;
      ba.in.base = d_front;
      ba.in.byte_off = nullptr;
      ba.in.stride_bytes = front_cap;
      ba.in.nbits = d_front_bits;
      ba.in.interleave = 1;
      ba.out.base = d_bac;
      ba.out.stride_bytes = bac_cap;
      ba.out.nbits = d_bit_len;
      ba.adaptive = p->adaptive;
      ba.div_table = c->d_div;
      ba.work = c->d_work;
      ba.status = d_status;
      ba.overflow_flag = reinterpret_cast<unsigned long long *>(d_flags);
      {
        Timed t(c, DC_K_BAC_ENC);
        launch_bac_encode(ba, c->stream);
      }
      if ((rc = check_launch(c, "bac_encode_kernel")) != DC_OK)
        return rc;
      d_final = d_bac;
      final_stride = bac_cap;
    }
    {
      Timed t(c, DC_K_SCAN);
      launch_byte_scan(d_bit_len, n, d_byte_off, d_tiles, c->stream);
    }
    if ((rc = check_launch(c, "scan kernels")) != DC_OK)
      return rc;
    {
      Timed t(c, DC_K_GATHER);
      launch_gather(d_final, final_stride, d_bit_len, d_byte_off, n, d_payload, payload_cap, c->stream);
    }
    if ((rc = check_launch(c, "gather_kernel")) != DC_OK)
      return rc;
    // one synchronisation per call: total size + overflow flag
    CU_TRY(c, cudaMemcpyAsync(d_flags + 1, d_byte_off + n, sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream), DC_ERR_LIBRARY_CALL);
    uint64_t host_flags[2];
    if ((rc = read_back(c, d_flags, host_flags, 2)) != DC_OK)
      return rc;
    if (host_flags[0] != 0 && boost < 512)
      continue;   // a stream outgrew the coder scratch: redo with a larger stride (never seen in practice)
    if (host_flags[0] != 0)
    {
      set_err(c, "arithmetic coder output exceeded the scratch capacity");
      return DC_ERR_MEMORY;
    }
    if (total_bytes)
      *total_bytes = host_flags[1];
    if (host_flags[1] > payload_cap)
    {
      set_err(c, "payload needs %llu bytes, capacity is %llu", (unsigned long long)host_flags[1], (unsigned long long)payload_cap);
      return DC_ERR_MEMORY;
    }
    return DC_OK;
  }
}

// ---- decode pipeline: [bac] -> back end ------------------------------------------------------------------
int decode_pipeline(dc_ctx *c, FrontKind kind, const uint8_t *d_payload, uint64_t payload_bytes, const uint64_t *d_bit_len,
                    const uint64_t *d_byte_off, const Layout &L, uint64_t max_len, const dc_params *p, float *d_samples, uint64_t *d_n_out, int32_t *d_status)
{
  const uint64_t n = L.n_streams;
  const uint32_t vs = p->value_size_bits;
  const uint64_t seg_cap = kind == FRONT_DEGA ? seg_cap_bytes(max_len, vs) : 0;
  c->ev_mask = 0;
  const uint64_t n_seg = round_up(n, 32);   // word-interleaved per warp (bitio.cuh)
  const uint64_t need = Bump::size_of(n, 8) + Bump::size_of(2, 8) + Bump::size_of(n_seg * seg_cap, 1) + 4096;
  int rc = arena_reserve(c, need);
  if (rc != DC_OK)
    return rc;
  Bump b(c->arena);
  uint64_t *d_seg_bits = b.take<uint64_t>(n);
  uint64_t *d_flags = b.take<uint64_t>(2);
  uint8_t *d_seg = b.take<uint8_t>(n_seg * seg_cap);

  CU_TRY(c, cudaMemsetAsync(d_status, 0, n * sizeof(int32_t), c->stream), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint64_t), c->stream), DC_ERR_LIBRARY_CALL);
  launch_validate_index(d_bit_len, d_byte_off, n, payload_bytes, d_status, c->stream);   // the index is the caller's: never trust it

  BackArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.lay = L;
  ka.n_streams = n;
  ka.vs = vs;
  ka.nf = p->normalization_factor;
  ka.samples = d_samples;
  ka.n_out = d_n_out;
  ka.status = d_status;
  ka.work = c->d_work;
  ka.do_denorm = 1;
  if (kind == FRONT_DEGA)
  {
    BacArgs ba;
    memset(&ba, 0, sizeof(ba));
    ba.n_streams = n;
    ba.in.base = d_payload;
    ba.in.byte_off = d_byte_off;
    ba.in.nbits = d_bit_len;
    ba.out.base = d_seg;
    ba.out.stride_bytes = seg_cap;
    ba.out.nbits = d_seg_bits;
    ba.out.interleave = 1;
    ba.adaptive = p->adaptive;
    ba.div_table = c->d_div;
      ba.work = c->d_work;
    ba.status = d_status;
    ba.overflow_flag = reinterpret_cast<unsigned long long *>(d_flags);
    {
      Timed t(c, DC_K_BAC_DEC);
      launch_bac_decode(ba, c->stream);
    }
    if ((rc = check_launch(c, "bac_decode_kernel")) != DC_OK)
      return rc;
    ka.in.base = d_seg;
    ka.in.byte_off = nullptr;
    ka.in.stride_bytes = seg_cap;
    ka.in.nbits = d_seg_bits;
    ka.in.interleave = 1;
    ka.in_seg = 1;
    ka.do_undiff = 1;
  }
  else
  {
    ka.in.base = d_payload;
    ka.in.byte_off = d_byte_off;
    ka.in.nbits = d_bit_len;
    ka.in_seg = 0;
    ka.do_undiff = 0;
  }
  {
    Timed t(c, DC_K_BACK);
    launch_back(ka, c->stream);
  }
  return check_launch(c, "back_kernel");
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
extern "C" {

int dc_abi_version(void) { return DC_ABI_VERSION; }

int dc_create(dc_ctx **out, int device)
{
  if (!out)
    return DC_ERR_INVALID_VALUE;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
  {
    cudaGetLastError();
    return DC_ERR_LIBRARY_INIT;
  }
  dc_ctx *c = new (std::nothrow) dc_ctx;
  if (!c)
    return DC_ERR_MEMORY;
  memset(c, 0, sizeof(*c));
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
  {
    cudaGetLastError();
    delete c;
    return DC_ERR_LIBRARY_INIT;
  }
  std::vector<DivEntry> tab(kDivTableEntries);
  build_div_table(tab.data());
  if (cudaMalloc(&c->d_div, kDivTableEntries * sizeof(DivEntry)) != cudaSuccess ||
      cudaMemcpy(c->d_div, tab.data(), kDivTableEntries * sizeof(DivEntry), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMalloc(&c->d_work, 256) != cudaSuccess ||
      cudaMallocHost(&c->h_pinned, 64 * sizeof(uint64_t)) != cudaSuccess)
  {
    cudaGetLastError();
    dc_destroy(c);
    return DC_ERR_LIBRARY_INIT;
  }
  for (int i = 0; i < 2 * DC_K_COUNT; i++)
    if (cudaEventCreate(&c->ev[i]) != cudaSuccess)
    {
      cudaGetLastError();
      dc_destroy(c);
      return DC_ERR_LIBRARY_INIT;
    }
  bool aux_ok = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < kMaxHostChunks && aux_ok; i++)
    aux_ok = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming) == cudaSuccess;
  if (!aux_ok)
  {
    cudaGetLastError();
    dc_destroy(c);
    return DC_ERR_LIBRARY_INIT;
  }
  c->launches_base = launch_counter();
  *out = c;
  return DC_OK;
}

void dc_destroy(dc_ctx *c)
{
  if (!c)
    return;
  dc_wait(c);
  cudaSetDevice(c->device);
  if (c->stream)
    cudaStreamSynchronize(c->stream);
  if (c->arena)
    cudaFree(c->arena);
  if (c->stage)
    cudaFree(c->stage);
  for (int i = 0; i < 2; i++)
    if (c->chain[i])
      cudaFree(c->chain[i]);
  if (c->d_div)
    cudaFree(c->d_div);
  if (c->d_work)
    cudaFree(c->d_work);
  if (c->h_pinned)
    cudaFreeHost(c->h_pinned);
  for (int i = 0; i < 2 * DC_K_COUNT; i++)
    if (c->ev[i])
      cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < kMaxHostChunks; i++)
  {
    if (c->ev_in[i])
      cudaEventDestroy(c->ev_in[i]);
    if (c->ev_out[i])
      cudaEventDestroy(c->ev_out[i]);
  }
  if (c->s_h2d)
    cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h)
    cudaStreamDestroy(c->s_d2h);
  if (c->stream)
    cudaStreamDestroy(c->stream);
  delete c;
}

const char *dc_last_error(const dc_ctx *c) { return c ? c->err : "no context"; }

uint64_t dc_layout_streams(const dc_layout *in)
{
  if (!in)
    return 0;
  if (in->sample_off)
    return in->n_streams;
  const uint64_t bl = in->block_len == 0 ? std::max<uint64_t>(in->channel_len, 1) : in->block_len;
  const uint64_t bpc = in->channel_len == 0 ? 1 : (in->channel_len + bl - 1) / bl;
  return in->n_channels * bpc;
}

int dc_synchronize(dc_ctx *c)
{
  if (!c)
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  return DC_OK;
}

void *dc_stream(dc_ctx *c) { return c ? (void *)c->stream : nullptr; }

uint64_t dc_launch_count(const dc_ctx *c) { return c ? launch_counter() - c->launches_base : 0; }

uint64_t dc_coder_wave_streams(dc_ctx *c)
{
  if (!c || cudaSetDevice(c->device) != cudaSuccess)
    return 0;
  return (uint64_t)sm_count() * (uint64_t)bac_threads_per_sm();
}

int dc_set_profiling(dc_ctx *c, int enable)
{
  if (!c)
    return DC_ERR_INVALID_VALUE;
  c->profiling = enable != 0;
  c->ev_mask = 0;
  return DC_OK;
}

int dc_get_kernel_ms(dc_ctx *c, float *ms)
{
  if (!c || !ms)
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  for (int i = 0; i < DC_K_COUNT; i++)
  {
    ms[i] = 0.0f;
    if (c->ev_mask & (1u << i))
      CU_TRY(c, cudaEventElapsedTime(&ms[i], c->ev[2 * i], c->ev[2 * i + 1]), DC_ERR_LIBRARY_CALL);
  }
  return DC_OK;
}

// ---- DEGA / A-XDR, device buffers ---------------------------------------------------------------------------
static int encode_device_common(dc_ctx *c, FrontKind kind, const float *d_samples, const dc_layout *layout, const dc_params *params,
                                uint8_t *d_payload, uint64_t payload_cap, uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status,
                                uint64_t *total_bytes)
{
  if (!c || !params_ok(params) || !layout || !d_bit_len || !d_byte_off || !d_status)
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  std::vector<uint64_t> host_off;
  if (layout->sample_off)
  {
    const int rc = fetch_offsets(c, layout, host_off);
    if (rc != DC_OK)
      return rc;
  }
  Layout L;
  uint64_t max_len, total;
  const int rc = make_layout(c, layout, layout->sample_off ? host_off.data() : nullptr, L, max_len, total);
  if (rc != DC_OK)
    return rc;
  if (kind == FRONT_AXDR)
  {
    const int frc = axdr_fast_encode(c, d_samples, L, params, d_payload, payload_cap, d_bit_len, d_byte_off, d_status, total_bytes);
    if (frc != DC_ERR_INVALID_MODE)   // INVALID_MODE = layout not eligible (or a sample failed): take the general path
      return frc;
  }
  return encode_pipeline(c, kind, d_samples, L, max_len, params, d_payload, payload_cap, d_bit_len, d_byte_off, d_status, total_bytes);
}

static int decode_device_common(dc_ctx *c, FrontKind kind, const uint8_t *d_payload, uint64_t payload_bytes, const uint64_t *d_bit_len,
                                const uint64_t *d_byte_off, const dc_layout *layout, const dc_params *params, float *d_samples,
                                uint64_t *d_n_out, int32_t *d_status)
{
  if (!c || !params_ok(params) || !layout || !d_bit_len || !d_byte_off || !d_n_out || !d_status)
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  std::vector<uint64_t> host_off;
  if (layout->sample_off)
  {
    const int rc = fetch_offsets(c, layout, host_off);
    if (rc != DC_OK)
      return rc;
  }
  Layout L;
  uint64_t max_len, total;
  const int rc = make_layout(c, layout, layout->sample_off ? host_off.data() : nullptr, L, max_len, total);
  if (rc != DC_OK)
    return rc;
  if (kind == FRONT_AXDR)
  {
    const int frc = axdr_fast_decode(c, d_payload, payload_bytes, L, params, d_samples, d_n_out, d_status);
    if (frc != DC_ERR_INVALID_MODE)
      return frc;
  }
  return decode_pipeline(c, kind, d_payload, payload_bytes, d_bit_len, d_byte_off, L, max_len, params, d_samples, d_n_out, d_status);
}

int dc_dega_encode_device(dc_ctx *c, const float *d_samples, const dc_layout *layout, const dc_params *params, uint8_t *d_payload,
                          uint64_t payload_cap, uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status, uint64_t *total_bytes)
{
  return encode_device_common(c, FRONT_DEGA, d_samples, layout, params, d_payload, payload_cap, d_bit_len, d_byte_off, d_status,
                              total_bytes);
}

int dc_axdr_encode_device(dc_ctx *c, const float *d_samples, const dc_layout *layout, const dc_params *params, uint8_t *d_payload,
                          uint64_t payload_cap, uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status, uint64_t *total_bytes)
{
  return encode_device_common(c, FRONT_AXDR, d_samples, layout, params, d_payload, payload_cap, d_bit_len, d_byte_off, d_status,
                              total_bytes);
}

int dc_dega_decode_device(dc_ctx *c, const uint8_t *d_payload, uint64_t payload_bytes, const uint64_t *d_bit_len,
                          const uint64_t *d_byte_off, const dc_layout *layout, const dc_params *params, float *d_samples,
                          uint64_t *d_n_out, int32_t *d_status)
{
  return decode_device_common(c, FRONT_DEGA, d_payload, payload_bytes, d_bit_len, d_byte_off, layout, params, d_samples, d_n_out,
                              d_status);
}

int dc_axdr_decode_device(dc_ctx *c, const uint8_t *d_payload, uint64_t payload_bytes, const uint64_t *d_bit_len,
                          const uint64_t *d_byte_off, const dc_layout *layout, const dc_params *params, float *d_samples,
                          uint64_t *d_n_out, int32_t *d_status)
{
  return decode_device_common(c, FRONT_AXDR, d_payload, payload_bytes, d_bit_len, d_byte_off, layout, params, d_samples, d_n_out,
                              d_status);
}

// ---- aggregate -------------------------------------------------------------------------------------------
int dc_aggregate_device(dc_ctx *c, const float *d_samples, uint64_t n_channels, uint64_t channel_len, uint32_t num_values, float *d_out)
{
  if (!c || num_values == 0 || (n_channels * channel_len != 0 && (!d_samples || !d_out)))
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  launch_aggregate(d_samples, n_channels, channel_len, num_values, d_out, c->stream);
  return check_launch(c, "aggregate_kernel");
}

// ---- csv ---------------------------------------------------------------------------------------------------
int dc_csv_decode_host(dc_ctx *c, const uint8_t *text, uint64_t nbytes, uint32_t column, char separator, float *out, uint64_t out_cap,
                       uint64_t *n_out)
{
  if (!c || column == 0 || (nbytes && !text) || !n_out)
    return DC_ERR_INVALID_VALUE;
  *n_out = 0;
  if (nbytes == 0)
    return DC_OK;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  c->ev_mask = 0;
  // phase 1 (staging block): the text and the line ends
  const uint64_t tiles = csv_tiles(nbytes);
  const uint64_t scan_tmp = (tiles + 4095) / 4096 + 2;
  int rc = stage_reserve(c, Bump::size_of(nbytes + 64, 1) + 2 * Bump::size_of(tiles + 1, 8) + Bump::size_of(scan_tmp, 8) + 4096);
  if (rc != DC_OK)
    return rc;
  Bump bs(c->stage);
  uint8_t *d_text = bs.take<uint8_t>(nbytes + 64);
  uint64_t *d_tile_bits = bs.take<uint64_t>(tiles + 1);
  uint64_t *d_tile_off = bs.take<uint64_t>(tiles + 1);
  uint64_t *d_tmp = bs.take<uint64_t>(scan_tmp);
  CU_TRY(c, cudaStreamSynchronize(c->s_d2h), DC_ERR_LIBRARY_CALL);   // the staging block may still be read by an earlier call's copy-back
  CU_TRY(c, cudaMemcpyAsync(d_text, text, nbytes, cudaMemcpyHostToDevice, c->stream), DC_ERR_LIBRARY_CALL);
  launch_csv_count(d_text, nbytes, d_tile_bits, c->stream);
  launch_byte_scan(d_tile_bits, tiles, d_tile_off, d_tmp, c->stream);
  if ((rc = check_launch(c, "csv line count")) != DC_OK)
    return rc;
  uint64_t n_newlines = 0;
  if ((rc = read_back(c, d_tile_off + tiles, &n_newlines, 1)) != DC_OK)
    return rc;
  const uint64_t n_lines = n_newlines + (text[nbytes - 1] != '\n' ? 1 : 0);   // an unterminated last line counts (csv.c:22)
  // phase 2 (arena): per-line results.  The list of fields that need strtof starts small; a text with more of them than
  // the list holds (the parse kernel counts them all) runs this phase a second time with the exact size.
  uint64_t hard_cap = n_lines < (1ull << 16) ? n_lines : (1ull << 16);
  uint64_t *d_flag8, *d_pos, *d_lo, *d_hi, *d_hard, *d_packed;
  float *d_value, *d_out;
  uint64_t total = 0, n_hard = 0;
  for (;;)
  {
    const uint64_t scan_tmp2 = (n_lines + 4095) / 4096 + 2;
    rc = arena_reserve(c, 5 * Bump::size_of(n_lines + 1, 8) + Bump::size_of(n_lines, 4) + Bump::size_of(out_cap, 4) + 4 * Bump::size_of(hard_cap, 8) +
                              Bump::size_of(scan_tmp2, 8) + Bump::size_of(8, 8) + 4096);
    if (rc != DC_OK)
      return rc;
    Bump ba(c->arena);
    uint64_t *d_line_end = ba.take<uint64_t>(n_lines + 1);
    d_flag8 = ba.take<uint64_t>(n_lines + 1);
    d_pos = ba.take<uint64_t>(n_lines + 1);
    d_lo = ba.take<uint64_t>(n_lines + 1);
    d_hi = ba.take<uint64_t>(n_lines + 1);
    d_value = ba.take<float>(n_lines);
    d_out = ba.take<float>(out_cap);
    d_hard = ba.take<uint64_t>(hard_cap);
    d_packed = ba.take<uint64_t>(3 * hard_cap);
    uint64_t *d_tmp2 = ba.take<uint64_t>(scan_tmp2);
    unsigned long long *d_nhard = reinterpret_cast<unsigned long long *>(ba.take<uint64_t>(8));
    CU_TRY(c, cudaMemsetAsync(d_nhard, 0, sizeof(unsigned long long), c->stream), DC_ERR_LIBRARY_CALL);
    launch_csv_lines(d_text, nbytes, d_tile_off, d_line_end, c->stream);
    launch_csv_parse(d_text, nbytes, d_line_end, n_newlines, n_lines, column, (uint8_t)separator, d_value, d_flag8, d_lo, d_hi, d_nhard, d_hard,
                     hard_cap, c->stream);
    launch_byte_scan(d_flag8, n_lines, d_pos, d_tmp2, c->stream);
    if ((rc = check_launch(c, "csv parse")) != DC_OK)
      return rc;
    if ((rc = read_back(c, d_pos + n_lines, &total, 1)) != DC_OK ||
        (rc = read_back(c, reinterpret_cast<const uint64_t *>(d_nhard), &n_hard, 1)) != DC_OK)
      return rc;
    if (n_hard <= hard_cap)
      break;
    hard_cap = n_hard;
  }
  *n_out = total;
  if (total > out_cap)
  {
    set_err(c, "csv column holds %llu values, capacity is %llu", (unsigned long long)total, (unsigned long long)out_cap);
    return DC_ERR_MEMORY;
  }
  launch_csv_compact(d_value, d_flag8, d_pos, n_lines, d_out, out_cap, c->stream);
  if ((rc = check_launch(c, "csv compact")) != DC_OK)
    return rc;
  if (total)
    CU_TRY(c, cudaMemcpyAsync(out, d_out, total * sizeof(float), cudaMemcpyDeviceToHost, c->stream), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  // fields outside the fast grammar (exponents, hex floats, inf/nan, long digit strings): strtof itself, like csv.c:31
  if (n_hard)
  {
    std::vector<uint64_t> packed(3 * n_hard);
    launch_csv_hard(d_hard, n_hard, d_lo, d_hi, d_pos, d_packed, c->stream);
    if ((rc = check_launch(c, "csv hard fields")) != DC_OK)
      return rc;
    CU_TRY(c, cudaMemcpyAsync(packed.data(), d_packed, 3 * n_hard * 8, cudaMemcpyDeviceToHost, c->stream), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
    std::string field;
    for (uint64_t h = 0; h < n_hard; h++)
    {
      field.assign(reinterpret_cast<const char *>(text) + packed[h], packed[n_hard + h] - packed[h]);
      out[packed[2 * n_hard + h]] = strtof(field.c_str(), nullptr);
    }
  }
  return DC_OK;
}

int dc_csv_encode_host(dc_ctx *c, const float *values, uint64_t n, uint32_t column, char separator, uint32_t num_decimal_places, uint8_t *text,
                       uint64_t text_cap, uint64_t *n_out)
{
  if (!c || column == 0 || num_decimal_places > 6 || (n && !values) || !n_out)
    return DC_ERR_INVALID_VALUE;
  *n_out = 0;
  if (n == 0)
    return DC_OK;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  c->ev_mask = 0;
  const uint64_t scan_tmp = (n + 4095) / 4096 + 2;
  int rc = stage_reserve(c, Bump::size_of(n, 4) + 4096);
  if (rc != DC_OK)
    return rc;
  if ((rc = arena_reserve(c, 2 * Bump::size_of(n + 1, 8) + Bump::size_of(scan_tmp, 8) + Bump::size_of(text_cap + 64, 1) + 4096)) != DC_OK)
    return rc;
  Bump bs(c->stage), ba(c->arena);
  float *d_values = bs.take<float>(n);
  uint64_t *d_len = ba.take<uint64_t>(n + 1);
  uint64_t *d_pos = ba.take<uint64_t>(n + 1);
  uint64_t *d_tmp = ba.take<uint64_t>(scan_tmp);
  uint8_t *d_text = ba.take<uint8_t>(text_cap + 64);
  CU_TRY(c, cudaStreamSynchronize(c->s_d2h), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaMemcpyAsync(d_values, values, n * sizeof(float), cudaMemcpyHostToDevice, c->stream), DC_ERR_LIBRARY_CALL);
  launch_csv_measure(d_values, n, num_decimal_places, column - 1, d_len, c->stream);
  launch_byte_scan(d_len, n, d_pos, d_tmp, c->stream);
  if ((rc = check_launch(c, "csv measure")) != DC_OK)
    return rc;
  uint64_t total = 0;
  if ((rc = read_back(c, d_pos + n, &total, 1)) != DC_OK)
    return rc;
  *n_out = total;
  if (total > text_cap)
  {
    set_err(c, "csv text needs %llu bytes, capacity is %llu", (unsigned long long)total, (unsigned long long)text_cap);
    return DC_ERR_MEMORY;
  }
  launch_csv_write(d_values, n, num_decimal_places, column - 1, (uint8_t)separator, d_pos, d_text, c->stream);
  if ((rc = check_launch(c, "csv write")) != DC_OK)
    return rc;
  CU_TRY(c, cudaMemcpyAsync(text, d_text, total, cudaMemcpyDeviceToHost, c->stream), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  return DC_OK;
}

// ---- DEGA, host buffers ----------------------------------------------------------------------------------
// Channel chunks of the host-buffer calls: the copy of chunk k+1 (and the copy-back of chunk k-1) overlaps the kernels
// of chunk k.  These calls are bound by the PCIe copies (1.4 GB in, 0.5 GB out for config 2 against ~15 ms of kernels),
// so chunks are sized for copy overlap rather than for full coder waves: a quarter wave (37 888 streams) still gives
// every SM eight warps through the warp-task scheduler, and eight chunks expose only 1/8 of the first copy in and of
// the last copy out.  Ragged layouts use one chunk.
static int host_chunks(const Layout &L, uint64_t n_channels)
{
  if (L.sample_off != nullptr || n_channels < 2)
    return 1;
  uint64_t per = (uint64_t)sm_count() * 256;   // streams per chunk: a quarter wave of the coder kernels (1024 threads per SM)
  if (const char *e = getenv("DC_HOST_CHUNK_STREAMS"))
    per = strtoull(e, nullptr, 10) > 0 ? strtoull(e, nullptr, 10) : per;
  uint64_t k = L.n_streams / per;
  if (k > (uint64_t)kMaxHostChunks)
    k = kMaxHostChunks;
  if (k > n_channels)
    k = n_channels;
  return k < 1 ? 1 : (int)k;
}

// Channel boundaries of the chunks, bounds[k] .. bounds[k+1]: equal shares.  (Tapering the sizes so that the chunk a
// call cannot overlap -- the last one of an encode, the first one of a decode -- is small was tried and measured
// slower, 35.7 / 37.0 ms against 31.7 / 32.7 ms for config 2: larger chunks in the middle cost more than the tail saves;
// cutting only that edge chunk in two halves, nine chunks in all, measured 65.9 against 65.0 ms per encode + decode.)
static void chunk_bounds(int n_chunks, uint64_t n_channels, uint64_t *bounds)
{
  for (int k = 0; k <= n_chunks; k++)
    bounds[k] = n_channels * (uint64_t)k / (uint64_t)n_chunks;
}

int dc_dega_encode_host(dc_ctx *c, const float *samples, const dc_layout *layout, const dc_params *params, uint8_t *payload,
                        uint64_t payload_cap, uint64_t *bit_len, uint64_t *byte_off, int32_t *status, uint64_t *total_bytes)
{
  if (!c || !params_ok(params) || !layout)
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  Layout L;
  uint64_t max_len, total;
  int rc = make_layout(c, layout, layout->sample_off, L, max_len, total);
  if (rc != DC_OK)
    return rc;
  const uint64_t n = L.n_streams;
  // device payload: every chunk starts 16-byte aligned, so up to 15 bytes of padding per chunk on top of the capacity
  const uint64_t dev_payload_cap = round_up(payload_cap, 16) + 16ull * kMaxHostChunks + 256;
  const uint64_t need = Bump::size_of(total, 4) + Bump::size_of(dev_payload_cap, 1) + Bump::size_of(n, 8) +
                        2 * Bump::size_of(n + 8, 8) + Bump::size_of(n, 4) + 4096;
  if ((rc = stage_reserve(c, need)) != DC_OK)
    return rc;
  Bump b(c->stage);
  float *d_samples = b.take<float>(total);
  uint8_t *d_payload = b.take<uint8_t>(dev_payload_cap);
  uint64_t *d_bits = b.take<uint64_t>(n);
  uint64_t *d_boff = b.take<uint64_t>(n + 8);
  uint64_t *d_off = b.take<uint64_t>(n + 8);
  int32_t *d_status = b.take<int32_t>(n);
  if (layout->sample_off)
  {
    CU_TRY(c, cudaMemcpyAsync(d_off, layout->sample_off, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream), DC_ERR_LIBRARY_CALL);
    L.sample_off = d_off;
  }
  const int n_chunks = host_chunks(L, layout->n_channels);
  uint64_t bounds[kMaxHostChunks + 1];
  chunk_bounds(n_chunks, layout->sample_off ? 0 : layout->n_channels, bounds);
  const uint64_t bpc = L.blocks_per_channel;
  // the staging block may have been read by the previous call's copy-back: order the new copies after it
  CU_TRY(c, cudaStreamSynchronize(c->s_d2h), DC_ERR_LIBRARY_CALL);
  auto chunk_range = [&](int k, uint64_t &c0, uint64_t &c1) {
    if (n_chunks == 1)
    {
      c0 = 0;
      c1 = layout->sample_off ? 0 : layout->n_channels;
      return;
    }
    c0 = bounds[k];
    c1 = bounds[k + 1];
  };
  auto upload = [&](int k) -> int {
    uint64_t c0, c1;
    chunk_range(k, c0, c1);
    const uint64_t first = n_chunks == 1 ? 0 : c0 * L.channel_len;
    const uint64_t count = n_chunks == 1 ? total : (c1 - c0) * L.channel_len;
    if (count)
      CU_TRY(c, cudaMemcpyAsync(d_samples + first, samples + first, count * 4, cudaMemcpyHostToDevice, c->s_h2d), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaEventRecord(c->ev_in[k], c->s_h2d), DC_ERR_LIBRARY_CALL);
    return DC_OK;
  };
  uint64_t base = 0;       // payload bytes of the chunks before this one (host side: packed densely)
  uint64_t dev_base = 0;   // where this chunk's payload starts on the device: 16-byte aligned and past everything the chunks
                           // before it wrote there (their copies to the host may still be in flight on s_d2h)
  uint64_t chunk_base[kMaxHostChunks], chunk_first[kMaxHostChunks], chunk_count[kMaxHostChunks];
  int result = DC_OK;
  if ((rc = upload(0)) != DC_OK)
    return rc;
  for (int k = 0; k < n_chunks; k++)
  {
    if (k + 1 < n_chunks && (rc = upload(k + 1)) != DC_OK)
      return rc;
    uint64_t c0, c1;
    chunk_range(k, c0, c1);
    Layout Lk = L;
    const uint64_t s0 = n_chunks == 1 ? 0 : c0 * bpc;
    if (n_chunks > 1)
      Lk.n_streams = (c1 - c0) * bpc;
    CU_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_in[k], 0), DC_ERR_LIBRARY_CALL);
    uint64_t tot = 0;
    const uint64_t cap_left = payload_cap > base ? payload_cap - base : 0;
    rc = encode_pipeline(c, FRONT_DEGA, d_samples + (n_chunks == 1 ? 0 : c0 * L.channel_len), Lk, max_len, params, d_payload + dev_base,
                         cap_left, d_bits + s0, d_boff + s0, d_status + s0, &tot);
    if (rc != DC_OK && rc != DC_ERR_MEMORY)
      return rc;
    if (rc == DC_ERR_MEMORY)
      result = rc;
    const uint64_t nk = Lk.n_streams;
    // encode_pipeline has synchronised the compute stream: the chunk's results are complete
    if (bit_len && nk)
      CU_TRY(c, cudaMemcpyAsync(bit_len + s0, d_bits + s0, nk * 8, cudaMemcpyDeviceToHost, c->s_d2h), DC_ERR_LIBRARY_CALL);
    if (byte_off && nk)
      CU_TRY(c, cudaMemcpyAsync(byte_off + s0, d_boff + s0, nk * 8, cudaMemcpyDeviceToHost, c->s_d2h), DC_ERR_LIBRARY_CALL);
    if (status && nk)
      CU_TRY(c, cudaMemcpyAsync(status + s0, d_status + s0, nk * 4, cudaMemcpyDeviceToHost, c->s_d2h), DC_ERR_LIBRARY_CALL);
    if (result == DC_OK && payload && tot)
      CU_TRY(c, cudaMemcpyAsync(payload + base, d_payload + dev_base, tot, cudaMemcpyDeviceToHost, c->s_d2h), DC_ERR_LIBRARY_CALL);
    chunk_base[k] = base;
    chunk_first[k] = s0;
    chunk_count[k] = nk;
    base += tot;
    dev_base = round_up(dev_base + tot, 16);
  }
  CU_TRY(c, cudaStreamSynchronize(c->s_d2h), DC_ERR_LIBRARY_CALL);
  if (byte_off)
  {
    // chunk-local offsets become container offsets (chunk 0 already is)
    for (int k = 1; k < n_chunks; k++)
      for (uint64_t i = 0; i < chunk_count[k]; i++)
        byte_off[chunk_first[k] + i] += chunk_base[k];
    byte_off[n] = base;
  }
  if (total_bytes)
    *total_bytes = base;
  return result;
}

int dc_dega_decode_host(dc_ctx *c, const uint8_t *payload, uint64_t payload_bytes, const uint64_t *bit_len, const uint64_t *byte_off,
                        const dc_layout *layout, const dc_params *params, float *samples, uint64_t *n_out, int32_t *status)
{
  if (!c || !params_ok(params) || !layout || !bit_len || !byte_off)
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  Layout L;
  uint64_t max_len, total;
  int rc = make_layout(c, layout, layout->sample_off, L, max_len, total);
  if (rc != DC_OK)
    return rc;
  const uint64_t n = L.n_streams;
  const uint64_t need = Bump::size_of(total, 4) + Bump::size_of(round_up(payload_bytes, 4) + 16, 1) + 2 * Bump::size_of(n, 8) +
                        2 * Bump::size_of(n + 1, 8) + Bump::size_of(n, 4) + 4096;
  if ((rc = stage_reserve(c, need)) != DC_OK)
    return rc;
  Bump b(c->stage);
  float *d_samples = b.take<float>(total);
  uint8_t *d_payload = b.take<uint8_t>(round_up(payload_bytes, 4) + 16);
  uint64_t *d_bits = b.take<uint64_t>(n);
  uint64_t *d_nout = b.take<uint64_t>(n);
  uint64_t *d_boff = b.take<uint64_t>(n + 1);
  uint64_t *d_off = b.take<uint64_t>(n + 1);
  int32_t *d_status = b.take<int32_t>(n);
  CU_TRY(c, cudaStreamSynchronize(c->s_d2h), DC_ERR_LIBRARY_CALL);
  if (layout->sample_off)
  {
    CU_TRY(c, cudaMemcpyAsync(d_off, layout->sample_off, (n + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d), DC_ERR_LIBRARY_CALL);
    L.sample_off = d_off;
  }
  // index first (small), then the payload chunk by chunk at its container offset
  if (n)
    CU_TRY(c, cudaMemcpyAsync(d_bits, bit_len, n * 8, cudaMemcpyHostToDevice, c->s_h2d), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaMemcpyAsync(d_boff, byte_off, (n + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d), DC_ERR_LIBRARY_CALL);
  const int n_chunks = host_chunks(L, layout->n_channels);
  uint64_t bounds[kMaxHostChunks + 1];
  chunk_bounds(n_chunks, layout->sample_off ? 0 : layout->n_channels, bounds);
  const uint64_t bpc = L.blocks_per_channel;
  auto chunk_streams = [&](int k, uint64_t &s0, uint64_t &s1, uint64_t &c0, uint64_t &c1) {
    if (n_chunks == 1)
    {
      s0 = 0;
      s1 = n;
      c0 = 0;
      c1 = layout->n_channels;
      return;
    }
    c0 = bounds[k];
    c1 = bounds[k + 1];
    s0 = c0 * bpc;
    s1 = c1 * bpc;
  };
  auto upload = [&](int k) -> int {
    uint64_t s0, s1, c0, c1;
    chunk_streams(k, s0, s1, c0, c1);
    const uint64_t b0 = byte_off[s0], b1 = byte_off[s1];
    if (b1 < b0 || b1 > payload_bytes)
      return DC_ERR_INVALID_VALUE;
    if (b1 > b0)
      CU_TRY(c, cudaMemcpyAsync(d_payload + b0, payload + b0, b1 - b0, cudaMemcpyHostToDevice, c->s_h2d), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaEventRecord(c->ev_in[k], c->s_h2d), DC_ERR_LIBRARY_CALL);
    return DC_OK;
  };
  if ((rc = upload(0)) != DC_OK)
    return rc;
  for (int k = 0; k < n_chunks; k++)
  {
    if (k + 1 < n_chunks && (rc = upload(k + 1)) != DC_OK)
      return rc;
    uint64_t s0, s1, c0, c1;
    chunk_streams(k, s0, s1, c0, c1);
    Layout Lk = L;
    if (n_chunks > 1)
      Lk.n_streams = s1 - s0;
    const uint64_t first = n_chunks == 1 ? 0 : c0 * L.channel_len;
    const uint64_t count = n_chunks == 1 ? total : (c1 - c0) * L.channel_len;
    CU_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_in[k], 0), DC_ERR_LIBRARY_CALL);
    rc = decode_pipeline(c, FRONT_DEGA, d_payload, payload_bytes, d_bits + s0, d_boff + s0, Lk, max_len, params, d_samples + first, d_nout + s0,
                         d_status + s0);
    if (rc != DC_OK)
      return rc;
    CU_TRY(c, cudaEventRecord(c->ev_out[k], c->stream), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaStreamWaitEvent(c->s_d2h, c->ev_out[k], 0), DC_ERR_LIBRARY_CALL);
    if (samples && count)
      CU_TRY(c, cudaMemcpyAsync(samples + first, d_samples + first, count * 4, cudaMemcpyDeviceToHost, c->s_d2h), DC_ERR_LIBRARY_CALL);
    if (n_out && s1 > s0)
      CU_TRY(c, cudaMemcpyAsync(n_out + s0, d_nout + s0, (s1 - s0) * 8, cudaMemcpyDeviceToHost, c->s_d2h), DC_ERR_LIBRARY_CALL);
    if (status && s1 > s0)
      CU_TRY(c, cudaMemcpyAsync(status + s0, d_status + s0, (s1 - s0) * 4, cudaMemcpyDeviceToHost, c->s_d2h), DC_ERR_LIBRARY_CALL);
  }
  CU_TRY(c, cudaStreamSynchronize(c->s_d2h), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
  {
    set_err(c, "kernel execution failed: %s", cudaGetErrorString(e));
    return DC_ERR_LIBRARY_CALL;
  }
  return DC_OK;
}

// ---- asynchronous forms of the two host-buffer calls --------------------------------------------------------
// The call runs on a worker thread owned by the context and dc_wait() joins it.  A context has its own streams, copy
// engines' queues and staging memory, so an encode on one context (mostly host -> device traffic) and a decode on
// another (mostly device -> host) use both directions of the PCIe link at once; one context alone is half duplex.
static int start_worker(dc_ctx *c, std::thread *t)
{
  if (!t)
    return DC_ERR_MEMORY;
  c->worker = t;
  return DC_OK;
}

int dc_dega_encode_host_async(dc_ctx *c, const float *samples, const dc_layout *layout, const dc_params *params, uint8_t *payload,
                              uint64_t payload_cap, uint64_t *bit_len, uint64_t *byte_off, int32_t *status, uint64_t *total_bytes)
{
  if (!c || !params_ok(params) || !layout)
    return DC_ERR_INVALID_VALUE;
  if (c->worker)
    return DC_ERR_INVALID_MODE;   // one call in flight per context
  const dc_layout lay = *layout;
  const dc_params par = *params;
  c->worker_rc = DC_OK;
  return start_worker(c, new (std::nothrow) std::thread([=]() {
                        c->worker_rc = dc_dega_encode_host(c, samples, &lay, &par, payload, payload_cap, bit_len, byte_off, status, total_bytes);
                      }));
}

int dc_dega_decode_host_async(dc_ctx *c, const uint8_t *payload, uint64_t payload_bytes, const uint64_t *bit_len, const uint64_t *byte_off,
                              const dc_layout *layout, const dc_params *params, float *samples, uint64_t *n_out, int32_t *status)
{
  if (!c || !params_ok(params) || !layout || !bit_len || !byte_off)
    return DC_ERR_INVALID_VALUE;
  if (c->worker)
    return DC_ERR_INVALID_MODE;
  const dc_layout lay = *layout;
  const dc_params par = *params;
  c->worker_rc = DC_OK;
  return start_worker(c, new (std::nothrow) std::thread([=]() {
                        c->worker_rc = dc_dega_decode_host(c, payload, payload_bytes, bit_len, byte_off, &lay, &par, samples, n_out, status);
                      }));
}

int dc_wait(dc_ctx *c)
{
  if (!c)
    return DC_ERR_INVALID_VALUE;
  if (!c->worker)
    return DC_OK;
  c->worker->join();
  delete c->worker;
  c->worker = nullptr;
  return c->worker_rc;
}

// ---- LZMH ---------------------------------------------------------------------------------------------------
static uint64_t lzmh_enc_cap_bytes(uint64_t max_len) { return round_up((max_len * 10 + 7) / 8 + 16, 16); }   // worst case: every byte a 10-bit literal

int dc_lzmh_encode_device(dc_ctx *c, const uint8_t *d_in, const uint64_t *d_in_off, uint64_t n, uint8_t *d_payload, uint64_t payload_cap,
                          uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status, uint64_t *total_bytes)
{
  if (!c || !d_in_off || !d_bit_len || !d_byte_off || !d_status)
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  c->ev_mask = 0;
  // block sizes are needed on the host to size the scratch stride
  std::vector<uint64_t> off(n + 1);
  CU_TRY(c, cudaMemcpyAsync(off.data(), d_in_off, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream), DC_ERR_LIBRARY_CALL);
  CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
  uint64_t max_len = 0;
  for (uint64_t s = 0; s < n; s++)
  {
    if (off[s + 1] < off[s])
      return DC_ERR_INVALID_VALUE;
    max_len = std::max(max_len, off[s + 1] - off[s]);
  }
  const uint64_t cap = lzmh_enc_cap_bytes(max_len);
  const uint64_t n_tiles = (n + 4095) / 4096;
  int rc = arena_reserve(c, Bump::size_of(n_tiles + 2, 8) + Bump::size_of(n * cap, 1) + 4096);
  if (rc != DC_OK)
    return rc;
  Bump b(c->arena);
  uint64_t *d_tiles = b.take<uint64_t>(n_tiles + 2);
  uint8_t *d_scratch = b.take<uint8_t>(n * cap);
  LzmhArgs la;
  memset(&la, 0, sizeof(la));
  la.n_streams = n;
  la.in = d_in;
  la.in_off = d_in_off;
  la.bits_out.base = d_scratch;
  la.bits_out.stride_bytes = cap;
  la.bits_out.nbits = d_bit_len;
  la.status = d_status;
  {
    Timed t(c, DC_K_LZMH_ENC);
    launch_lzmh_encode(la, c->stream);
  }
  {
    Timed t(c, DC_K_SCAN);
    launch_byte_scan(d_bit_len, n, d_byte_off, d_tiles, c->stream);
  }
  {
    Timed t(c, DC_K_GATHER);
    launch_gather(d_scratch, cap, d_bit_len, d_byte_off, n, d_payload, payload_cap, c->stream);
  }
  if ((rc = check_launch(c, "lzmh encode kernels")) != DC_OK)
    return rc;
  uint64_t tot = 0;
  if ((rc = read_back(c, d_byte_off + n, &tot, 1)) != DC_OK)
    return rc;
  if (total_bytes)
    *total_bytes = tot;
  if (tot > payload_cap)
  {
    set_err(c, "payload needs %llu bytes, capacity is %llu", (unsigned long long)tot, (unsigned long long)payload_cap);
    return DC_ERR_MEMORY;
  }
  return DC_OK;
}

int dc_lzmh_decode_device(dc_ctx *c, const uint8_t *d_payload, uint64_t payload_bytes, const uint64_t *d_bit_len,
                          const uint64_t *d_byte_off, uint64_t n, uint8_t *d_out, const uint64_t *d_out_off, uint64_t *d_out_len,
                          int32_t *d_status)
{
  if (!c || !d_bit_len || !d_byte_off || !d_out_off || !d_out_len || !d_status)
    return DC_ERR_INVALID_VALUE;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  c->ev_mask = 0;
  CU_TRY(c, cudaMemsetAsync(d_status, 0, n * sizeof(int32_t), c->stream), DC_ERR_LIBRARY_CALL);
  launch_validate_index(d_bit_len, d_byte_off, n, payload_bytes, d_status, c->stream);
  LzmhArgs la;
  memset(&la, 0, sizeof(la));
  la.n_streams = n;
  la.bits_in.base = d_payload;
  la.bits_in.byte_off = d_byte_off;
  la.bits_in.nbits = d_bit_len;
  la.out = d_out;
  la.out_off = d_out_off;
  la.out_len = d_out_len;
  la.status = d_status;
  {
    Timed t(c, DC_K_LZMH_DEC);
    launch_lzmh_decode(la, c->stream);
  }
  return check_launch(c, "lzmh_decode_kernel");
}

// one LZMH stage over one host stream (plugin face)
static int lzmh_stage_host(dc_ctx *c, int encode, const uint8_t *in, cudaMemcpyKind in_kind, uint64_t in_bits, uint8_t *out, cudaMemcpyKind out_kind,
                           uint64_t out_cap, uint64_t *out_bits)
{
  const uint64_t in_bytes = (in_bits + 7) / 8;
  if (encode && (in_bits & 7))
    return DC_ERR_LIBRARY_CALL;   // 8-bit reads (lzmh.c:165): a fractional last byte is a short read
  for (unsigned boost = 1;; boost *= 8)
  {
    const uint64_t dev_cap = encode ? lzmh_enc_cap_bytes(in_bytes) : round_up((in_bytes * 8 + 4096) * boost, 16);
    int rc = arena_reserve(c, Bump::size_of(in_bytes + 64, 1) + Bump::size_of(dev_cap + 64, 1) + Bump::size_of(16, 8) + 4096);
    if (rc != DC_OK)
      return rc;
    Bump b(c->arena);
    uint8_t *d_in = b.take<uint8_t>(in_bytes + 64);
    uint8_t *d_out = b.take<uint8_t>(dev_cap + 64);
    uint64_t *d_meta = b.take<uint64_t>(16);   // [0..1] in_off / out_off, [2] bits in, [3] bits/bytes out, [4] status
    const uint64_t meta[6] = { 0, encode ? in_bytes : dev_cap, in_bits, 0, 0, 0 };
    CU_TRY(c, cudaMemcpyAsync(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice, c->stream), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaMemsetAsync(d_in + (in_bytes & ~(uint64_t)15), 0, 32, c->stream), DC_ERR_LIBRARY_CALL);
    if (in_bytes)
      CU_TRY(c, cudaMemcpyAsync(d_in, in, in_bytes, in_kind, c->stream), DC_ERR_LIBRARY_CALL);
    LzmhArgs la;
    memset(&la, 0, sizeof(la));
    la.n_streams = 1;
    la.status = reinterpret_cast<int32_t *>(d_meta + 4);
    if (encode)
    {
      la.in = d_in;
      la.in_off = d_meta;
      la.bits_out.base = d_out;
      la.bits_out.stride_bytes = dev_cap;
      la.bits_out.nbits = d_meta + 3;
      launch_lzmh_encode(la, c->stream);
    }
    else
    {
      la.bits_in.base = d_in;
      la.bits_in.byte_off = nullptr;
      la.bits_in.stride_bytes = 0;
      la.bits_in.nbits = d_meta + 2;
      la.out = d_out;
      la.out_off = d_meta;
      la.out_len = d_meta + 3;
      launch_lzmh_decode(la, c->stream);
    }
    if ((rc = check_launch(c, "lzmh stage kernel")) != DC_OK)
      return rc;
    uint64_t back[6];
    if ((rc = read_back(c, d_meta, back, 6)) != DC_OK)
      return rc;
    const int32_t st = (int32_t)(back[4] & 0xFFFFFFFFu);
    if (st == DC_ERR_MEMORY && !encode && boost < 4096)
      continue;   // decoded stream outgrew the guess: retry with a larger window
    if (st != 0)
      return st;
    const uint64_t bits = encode ? back[3] : back[3] * 8;
    *out_bits = bits;
    const uint64_t nbytes = (bits + 7) / 8;
    if (nbytes > out_cap)
      return DC_ERR_MEMORY;
    if (nbytes)
    {
      CU_TRY(c, cudaMemcpyAsync(out, d_out, nbytes, out_kind, c->stream), DC_ERR_LIBRARY_CALL);
      CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
    }
    return DC_OK;
  }
}

// ---- plugin face: one stage over one stream ---------------------------------------------------------------
uint64_t dc_stage_out_bound(int stage, uint64_t in_bits, const dc_params *p)
{
  const uint64_t vs = p ? p->value_size_bits : 32;
  switch (stage)
  {
    case DC_STAGE_NORMALIZE: return (in_bits / 32 * vs + 7) / 8 + 8;
    case DC_STAGE_DENORMALIZE: return in_bits / vs * 4 + 8;
    case DC_STAGE_DIFF_ENC:
    case DC_STAGE_DIFF_DEC: return (in_bits + 7) / 8 + 8;
    case DC_STAGE_SEG_ENC: return (in_bits / vs * (2 * vs + 1) + 7) / 8 + 8;
    case DC_STAGE_SEG_DEC: return (in_bits * vs + 7) / 8 + 8;
    case DC_STAGE_BAC_ENC: return (in_bits * 13 / 8 + 7) / 8 + 64;
    case DC_STAGE_BAC_DEC: return in_bits * 4 + 4096;    /* first guess; dc_stage_run_host reports the exact need on overflow */
    case DC_STAGE_LZMH_ENC: return (in_bits / 8 * 10 + 7) / 8 + 8;
    case DC_STAGE_LZMH_DEC: return in_bits / 10 * 274 / 2 + 4096;
    case DC_STAGE_DEGA_ENC: return (in_bits / 32 * (2 * vs + 1) * 13 / 8 + 7) / 8 + 64;
    case DC_STAGE_DEGA_DEC: return in_bits * 16 + 4096;
    default: return 0;
  }
}

// One stage over one stream.  `in` / `out` are host or device buffers as the copy kinds say (device for the inner stages
// of dc_chain_run_host).
static int stage_run(dc_ctx *c, int stage, const uint8_t *in, cudaMemcpyKind in_kind, uint64_t in_bits, const dc_params *params, uint8_t *out,
                     cudaMemcpyKind out_kind, uint64_t out_cap, uint64_t *out_bits)
{
  if (!c || !params_ok(params) || !out_bits || stage < 0 || stage >= DC_STAGE_COUNT)
    return DC_ERR_INVALID_VALUE;
  *out_bits = 0;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  if (stage == DC_STAGE_LZMH_ENC || stage == DC_STAGE_LZMH_DEC)
    return lzmh_stage_host(c, stage == DC_STAGE_LZMH_ENC, in, in_kind, in_bits, out, out_kind, out_cap, out_bits);

  const uint32_t vs = params->value_size_bits;
  const uint64_t in_bytes = (in_bits + 7) / 8;
  const bool float_in = stage == DC_STAGE_NORMALIZE || stage == DC_STAGE_DEGA_ENC;
  const bool float_out = stage == DC_STAGE_DENORMALIZE || stage == DC_STAGE_DEGA_DEC;
  const uint64_t n_in = float_in ? in_bits / 32 : in_bits / vs;   // values in (front-end stages)

  // single-stream geometry; all scratch strides are per-stream capacities
  Layout L;
  L.n_streams = 1;
  L.channel_len = n_in;
  L.block_len = n_in ? n_in : 1;
  L.blocks_per_channel = 1;
  L.sample_off = nullptr;

  uint64_t mid_cap = 0, out_cap_dev = 0;   // bytes
  for (unsigned boost = 1;; boost *= 8)
  {
    switch (stage)
    {
      case DC_STAGE_NORMALIZE:
      case DC_STAGE_DIFF_ENC:
      case DC_STAGE_DIFF_DEC: out_cap_dev = fixed_cap_bytes(std::max(n_in, in_bits / vs + 1), vs); break;
      case DC_STAGE_SEG_ENC: out_cap_dev = seg_cap_bytes(n_in, vs); break;
      case DC_STAGE_SEG_DEC: out_cap_dev = fixed_cap_bytes(in_bits + 1, vs); break;
      case DC_STAGE_DENORMALIZE: out_cap_dev = round_up((in_bits / vs + 1) * 4, 16); break;
      case DC_STAGE_BAC_ENC: out_cap_dev = bac_cap_bytes(round_up(in_bytes + 8, 16), params->adaptive, boost); break;
      case DC_STAGE_BAC_DEC: out_cap_dev = round_up((in_bytes * 4 + 4096) * boost, 16); break;
      case DC_STAGE_DEGA_ENC:
        mid_cap = seg_cap_bytes(n_in, vs);
        out_cap_dev = bac_cap_bytes(mid_cap, params->adaptive, boost);
        break;
      case DC_STAGE_DEGA_DEC:
        mid_cap = round_up((in_bytes * 4 + 4096) * boost, 16);
        out_cap_dev = 0;   // sized after the coder reports the SEG bit count
        break;
      default: return DC_ERR_INVALID_VALUE;
    }
    // DEGA decode needs the sample capacity: every SEG code word is at least one bit long
    const uint64_t dega_dec_samples = mid_cap * 8;
    if (stage == DC_STAGE_DEGA_DEC)
      out_cap_dev = round_up(dega_dec_samples * 4 + 16, 16);

    const uint64_t need = Bump::size_of(round_up(in_bytes, 16) + 32, 1) + Bump::size_of(mid_cap, 1) + Bump::size_of(out_cap_dev, 1) +
                          Bump::size_of(8, 8) + 4096;
    int rc = arena_reserve(c, need);
    if (rc != DC_OK)
      return rc;
    Bump b(c->arena);
    uint8_t *d_in = b.take<uint8_t>(round_up(in_bytes, 16) + 32);
    uint8_t *d_mid = b.take<uint8_t>(mid_cap);
    uint8_t *d_out = b.take<uint8_t>(out_cap_dev);
    uint64_t *d_meta = b.take<uint64_t>(8);   // [0]=in_bits [1]=mid_bits [2]=out_bits [3]=n_out [4]=status(int32) [5]=overflow flag
    int32_t *d_status = reinterpret_cast<int32_t *>(d_meta + 4);

    CU_TRY(c, cudaMemsetAsync(d_meta, 0, 8 * sizeof(uint64_t), c->stream), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaMemsetAsync(d_in + (in_bytes & ~(uint64_t)15), 0, 32, c->stream), DC_ERR_LIBRARY_CALL);
    if (in_bytes)
      CU_TRY(c, cudaMemcpyAsync(d_in, in, in_bytes, in_kind, c->stream), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaMemcpyAsync(d_meta, &in_bits, sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream), DC_ERR_LIBRARY_CALL);

    BitSrc src_in;
    src_in.base = d_in;
    src_in.byte_off = nullptr;
    src_in.stride_bytes = 0;
    src_in.nbits = d_meta + 0;

    const bool is_front = stage == DC_STAGE_NORMALIZE || stage == DC_STAGE_DIFF_ENC || stage == DC_STAGE_SEG_ENC || stage == DC_STAGE_DEGA_ENC;
    const bool is_back = stage == DC_STAGE_DENORMALIZE || stage == DC_STAGE_DIFF_DEC || stage == DC_STAGE_SEG_DEC || stage == DC_STAGE_DEGA_DEC;

    if (is_front)
    {
      FrontArgs fa;
      memset(&fa, 0, sizeof(fa));
      fa.lay = L;
      fa.samples = reinterpret_cast<const float *>(d_in);
      fa.in = src_in;
      fa.in_is_float = float_in;
      fa.do_norm = float_in;
      fa.do_diff = stage == DC_STAGE_DIFF_ENC || stage == DC_STAGE_DEGA_ENC;
      fa.out_seg = stage == DC_STAGE_SEG_ENC || stage == DC_STAGE_DEGA_ENC;
      fa.vs = vs;
      fa.nf = params->normalization_factor;
      fa.out.base = stage == DC_STAGE_DEGA_ENC ? d_mid : d_out;
      fa.out.stride_bytes = stage == DC_STAGE_DEGA_ENC ? mid_cap : out_cap_dev;
      fa.out.nbits = stage == DC_STAGE_DEGA_ENC ? d_meta + 1 : d_meta + 2;
      fa.status = d_status;
      fa.work = c->d_work;
      launch_front(fa, c->stream);
    }
    if (stage == DC_STAGE_BAC_ENC || stage == DC_STAGE_DEGA_ENC)
    {
      BacArgs ba;
      memset(&ba, 0, sizeof(ba));
      ba.n_streams = 1;
      if (stage == DC_STAGE_BAC_ENC)
        ba.in = src_in;
      else
      {
        ba.in.base = d_mid;
        ba.in.byte_off = nullptr;
        ba.in.stride_bytes = mid_cap;
        ba.in.nbits = d_meta + 1;
      }
      ba.out.base = d_out;
      ba.out.stride_bytes = out_cap_dev;
      ba.out.nbits = d_meta + 2;
      ba.adaptive = params->adaptive;
      ba.div_table = c->d_div;
      ba.work = c->d_work;
      ba.status = d_status;
      ba.overflow_flag = reinterpret_cast<unsigned long long *>(d_meta + 5);
      launch_bac_encode(ba, c->stream);
    }
    if (stage == DC_STAGE_BAC_DEC || stage == DC_STAGE_DEGA_DEC)
    {
      BacArgs ba;
      memset(&ba, 0, sizeof(ba));
      ba.n_streams = 1;
      ba.in = src_in;
      ba.out.base = stage == DC_STAGE_BAC_DEC ? d_out : d_mid;
      ba.out.stride_bytes = stage == DC_STAGE_BAC_DEC ? out_cap_dev : mid_cap;
      ba.out.nbits = stage == DC_STAGE_BAC_DEC ? d_meta + 2 : d_meta + 1;
      ba.adaptive = params->adaptive;
      ba.div_table = c->d_div;
      ba.work = c->d_work;
      ba.status = d_status;
      ba.overflow_flag = reinterpret_cast<unsigned long long *>(d_meta + 5);
      launch_bac_decode(ba, c->stream);
    }
    if (is_back)
    {
      BackArgs ka;
      memset(&ka, 0, sizeof(ka));
      Layout LO = L;
      const uint64_t cap_samples = stage == DC_STAGE_DEGA_DEC ? dega_dec_samples : in_bits / vs + 1;
      LO.channel_len = cap_samples;
      LO.block_len = cap_samples;
      ka.lay = LO;
      ka.n_streams = 1;
      if (stage == DC_STAGE_DEGA_DEC)
      {
        ka.in.base = d_mid;
        ka.in.byte_off = nullptr;
        ka.in.stride_bytes = mid_cap;
        ka.in.nbits = d_meta + 1;
      }
      else
        ka.in = src_in;
      ka.in_seg = stage == DC_STAGE_SEG_DEC || stage == DC_STAGE_DEGA_DEC;
      ka.do_undiff = stage == DC_STAGE_DIFF_DEC || stage == DC_STAGE_DEGA_DEC;
      ka.do_denorm = float_out;
      ka.vs = vs;
      ka.nf = params->normalization_factor;
      ka.samples = reinterpret_cast<float *>(d_out);
      ka.out.base = d_out;
      ka.out.stride_bytes = out_cap_dev;
      ka.out.nbits = d_meta + 2;
      ka.n_out = d_meta + 3;
      ka.status = d_status;
      ka.work = c->d_work;
      launch_back(ka, c->stream);
    }
    if ((rc = check_launch(c, "stage kernels")) != DC_OK)
      return rc;
    uint64_t meta[8];
    if ((rc = read_back(c, d_meta, meta, 8)) != DC_OK)
      return rc;
    const int32_t st = (int32_t)(meta[4] & 0xFFFFFFFFu);
    if (meta[5] != 0 && boost < 4096)
      continue;   // coder scratch too small: retry with a larger stride
    if (st != 0)
      return st;
    uint64_t bits = float_out ? meta[3] * 32 : meta[2];
    if (float_in && (in_bits % 32) != 0)
      return DC_ERR_LIBRARY_CALL;   // short read of the last float (io_macros.h:13-27)
    *out_bits = bits;
    const uint64_t nbytes = (bits + 7) / 8;
    if (nbytes > out_cap)
    {
      set_err(c, "stage output needs %llu bytes, capacity is %llu", (unsigned long long)nbytes, (unsigned long long)out_cap);
      return DC_ERR_MEMORY;
    }
    if (nbytes)
    {
      CU_TRY(c, cudaMemcpyAsync(out, d_out, nbytes, out_kind, c->stream), DC_ERR_LIBRARY_CALL);
      CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
    }
    return DC_OK;
  }
}

int dc_stage_run_host(dc_ctx *c, int stage, const uint8_t *in, uint64_t in_bits, const dc_params *params, uint8_t *out, uint64_t out_cap,
                      uint64_t *out_bits)
{
  return stage_run(c, stage, in, cudaMemcpyHostToDevice, in_bits, params, out, cudaMemcpyDeviceToHost, out_cap, out_bits);
}

// ---- chains: `encode X # encode Y # ...` with the intermediate results kept on the device ---------------------
static int chain_reserve(dc_ctx *c, int which, uint64_t bytes)
{
  if (bytes <= c->chain_size[which])
    return DC_OK;
  if (c->chain[which])
    cudaFree(c->chain[which]);
  c->chain[which] = nullptr;
  c->chain_size[which] = 0;
  const uint64_t want = round_up(bytes + bytes / 4 + 4096, 4096);
  CU_TRY(c, cudaMalloc(&c->chain[which], want), DC_ERR_MEMORY);
  c->chain_size[which] = want;
  return DC_OK;
}

int dc_chain_run_host(dc_ctx *c, uint32_t n_stages, const int *stages, const dc_params *params, const uint8_t *in, uint64_t in_bits, uint8_t *out,
                      uint64_t out_cap, uint64_t *out_bits, uint32_t *failed_stage)
{
  if (!c || n_stages == 0 || n_stages > 16 || !stages || !params || !out_bits)   // 16 = MAX_NUM_ENCODERS, DCCLI/inc/params.h:10
    return DC_ERR_INVALID_VALUE;
  *out_bits = 0;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  const uint8_t *cur = in;
  uint64_t bits = in_bits;
  for (uint32_t i = 0; i < n_stages; i++)
  {
    if (failed_stage)
      *failed_stage = i;
    const bool first = i == 0, last = i + 1 == n_stages;
    if (last)
      return stage_run(c, stages[i], cur, first ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, bits, &params[i], out, cudaMemcpyDeviceToHost,
                       out_cap, out_bits);
    // an inner stage writes into the ping-pong buffer the next one reads; the decoders' sizes are only known afterwards,
    // so a result that did not fit reports its size and the stage runs once more
    const int which = (int)(i & 1);
    uint64_t need = dc_stage_out_bound(stages[i], bits, &params[i]), got = 0;
    int rc = DC_OK;
    for (int attempt = 0; attempt < 2; attempt++)
    {
      if ((rc = chain_reserve(c, which, need)) != DC_OK)
        return rc;
      rc = stage_run(c, stages[i], cur, first ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, bits, &params[i], c->chain[which],
                     cudaMemcpyDeviceToDevice, c->chain_size[which], &got);
      if (rc != DC_ERR_MEMORY || (got + 7) / 8 <= c->chain_size[which])
        break;
      need = (got + 7) / 8;
    }
    if (rc != DC_OK)
      return rc;
    cur = c->chain[which];
    bits = got;
  }
  return DC_OK;
}

}  // extern "C"
