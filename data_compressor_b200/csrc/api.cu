// api.cu -- host side of the C ABI declared in include/dc_cuda.h.
//
// No CPU fallback anywhere in this file: every codec result is produced by the kernels in dega_kernels.cu /
// lzmh_kernels.cu.  Without a usable GPU dc_create() fails with DC_ERR_LIBRARY_INIT.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
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
  cudaStream_t s_h2d, s_d2h;          // the copy streams of the host-buffer calls: every copy in / every copy out of every job
  uint64_t *h_pinned;   // small pinned mailbox for device->host scalars
  uint64_t launches_base;
  int profiling;
  cudaEvent_t ev[2 * DC_K_COUNT];
  uint32_t ev_mask;   // kernels timed during the last call
  struct JobEngine *eng;                    // worker thread + queue of the host-buffer calls (see "the job engine")
  std::vector<struct HostJob *> *async_jobs;   // *_host_async jobs not yet collected by dc_wait
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

// DC_TRACE=1: host-side timeline of the host-buffer calls on stderr (which call, which chunk, when), for finding out what
// two contexts working side by side wait for
bool trace_on()
{
  static const bool on = getenv("DC_TRACE") != nullptr;
  return on;
}
void trace(const dc_ctx *c, const char *what, int k = -1)
{
  if (!trace_on())
    return;
  static const auto t0 = std::chrono::steady_clock::now();
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  fprintf(stderr, "[dc %p %10.3f ms] %s %d\n", (const void *)c, ms, what, k);
}

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

// pub == nullptr: blocking; runs again with a larger coder scratch if a stream outgrew it, returns the total.
// pub != nullptr: one pass with the scratch factor `boost0`, nothing is waited for; the overflow flag and the total
// appear in pub[0..1] (mapped host memory) once the stream has got there (the job engine reads them after an event).
int encode_pipeline(dc_ctx *c, FrontKind kind, const float *d_samples, const Layout &L, uint64_t max_len, const dc_params *p,
                    uint8_t *d_payload, uint64_t payload_cap, uint64_t *d_bit_len, uint64_t *d_byte_off, int32_t *d_status,
                    uint64_t *total_bytes, uint64_t *pub = nullptr, unsigned boost0 = 1)
{
  const uint64_t n = L.n_streams;
  const uint32_t vs = p->value_size_bits;
  c->ev_mask = 0;
  for (unsigned boost = boost0;; boost *= 8)
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
    if (pub)
    {
      launch_publish(d_flags, d_byte_off + n, pub, c->stream);
      return check_launch(c, "publish_kernel");
    }
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
static int engine_create(dc_ctx *c);
static void engine_destroy(dc_ctx *c);

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
  const bool aux_ok = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
                      cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) == cudaSuccess && engine_create(c) == DC_OK;
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
  cudaSetDevice(c->device);
  engine_destroy(c);   // finishes what is queued, stops the worker
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

// ---- DEGA, host buffers: the job engine ---------------------------------------------------------------------
// A host-buffer call is a JOB: its channels are cut into chunks, and every chunk goes through copy in -> kernels -> copy
// out.  One worker thread per context runs the jobs; at most one encode and one decode are in flight at a time, each on
// its own compute stream with its own scratch (a "lane"), and ALL copies of both go through one host->device stream and
// one device->host stream, in the order the worker issues them.  That order is the point:
//   * an encode is mostly host->device traffic, a decode mostly device->host, so the two together use the PCIe link full
//     duplex (measured on this pool: 1.4 GB one way 25.5 ms, 1.4 GB each way at once 28.5 ms);
//   * left to the hardware -- two contexts with their own copy streams -- the two calls do NOT overlap: a copy stream
//     that always has its next copy queued keeps the engine (the other call's first 62 MB copy landed after the first
//     call's last one, 30 ms later), and a copy that waits for an event blocks ready copies queued behind it, including
//     the 16-byte read-back of a chunk's size (profiles/r2_e2e_trace.txt).  So the worker only ever issues copies that
//     are READY (a chunk's results are issued after its kernels have finished, never behind an event), cuts every copy
//     into pieces of 32 MB and keeps only three of them in flight per direction, so that it decides every ~0.6 ms
//     whose piece goes next -- host -> device: the decode's first (its payload is a third of the bytes and unlocks
//     the long copy back), device -> host: in turns -- and reads a chunk's payload size from mapped host memory
//     (publish_kernel) instead of through a copy.
// Chunk size: the coder kernels run one stream per thread, so a chunk needs about half a wave of streams
// (dc_coder_wave_streams) to keep the schedulers busy; smaller chunks hide more of the first copy in and the last copy
// out but their kernels are latency bound (a quarter-wave chunk takes 3 ms instead of 1), and the encode and decode
// kernels cannot share an SM (each takes its whole register file).
static const int kJobKinds = 2;   // 0: encode jobs, 1: decode jobs; one lane (compute stream + scratch) each
static const int kJobSlots = 4;   // two jobs per kind can be in flight (slot = 2 * kind + buffer): the second one is admitted
                                  // when every upload of the first has been issued, so that its uploads fill the time the
                                  // first one spends on its last kernels and its last copy back (6 ms per job otherwise)

struct HostJob
{
  int kind;   // 0 encode, 1 decode
  const float *samples_in;
  float *samples_out;
  const uint8_t *payload_in;
  uint8_t *payload_out;
  uint64_t payload_size;   // capacity (encode) / bytes (decode)
  const uint64_t *bit_len_in, *byte_off_in;
  uint64_t *bit_len_out, *byte_off_out;
  int32_t *status;
  uint64_t *n_out, *total_bytes;
  dc_layout layout;
  dc_params params;
  int rc;
  bool done;
};

struct Lane   // what the kernel pipelines take from the context: swapped in while a job's chunk is launched
{
  cudaStream_t stream;
  uint8_t *arena;
  uint64_t arena_size;
  unsigned int *d_work;
};

struct Piece   // one copy of at most kPieceBytes; `after` (if set) is recorded behind it
{
  void *dst;
  const void *src;
  uint64_t bytes;
  cudaEvent_t after;
};
static const uint64_t kPieceBytes = 32ull << 20;
static const int kPiecesInFlight = 3;   // per direction: ~2 ms of copies queued, so a piece that becomes ready never waits longer than that

struct CopyDir   // one direction of the link: its stream and the markers of the pieces in flight
{
  cudaStream_t stream;
  cudaMemcpyKind kind;
  cudaEvent_t ring[8];
  int head, count;
  int turn;   // device -> host: whose piece goes next when both jobs have one
};

struct JobRun   // a job admitted to its slot
{
  HostJob *job;
  int slot;
  uint64_t age;   // admission number: within a kind the older job goes first everywhere
  std::deque<Piece> h2d, d2h;   // copies decided on, not yet handed to the copy streams
  int up_issued, out_issued;    // chunks whose last upload piece / last result piece has been issued (its event recorded)
  Layout L;
  uint64_t max_len, total, n, bpc;
  int n_chunks;
  uint64_t bounds[kMaxHostChunks + 1];
  float *d_samples;
  uint8_t *d_payload;
  uint64_t *d_bits, *d_boff, *d_off, *d_nout;
  int32_t *d_status;
  uint64_t region;   // encode: device payload bytes set aside per chunk
  int up, kn, fin, done;   // chunks with: upload queued / kernels queued / results queued / results landed
  unsigned boost[kMaxHostChunks];
  uint64_t base;     // encode: payload bytes of the chunks finished so far (host side: packed densely)
  uint64_t chunk_base[kMaxHostChunks], chunk_first[kMaxHostChunks], chunk_count[kMaxHostChunks];
  int result;
  bool finished;
};

struct JobEngine
{
  std::mutex mu;
  std::condition_variable cv, cv_done;
  std::deque<HostJob *> queue;
  std::thread worker;
  bool started = false, quit = false;
  JobRun *run[kJobSlots] = { nullptr, nullptr, nullptr, nullptr };
  uint64_t submitted = 0, completed = 0, admitted = 0;
  Lane lane[kJobKinds];
  uint8_t *stage[kJobSlots] = { nullptr, nullptr, nullptr, nullptr };
  uint64_t stage_size[kJobSlots] = { 0, 0, 0, 0 };
  cudaEvent_t ev_in[kJobSlots][kMaxHostChunks], ev_k[kJobSlots][kMaxHostChunks], ev_out[kJobSlots][kMaxHostChunks];
  CopyDir dir[2];   // 0: host -> device, 1: device -> host
  uint64_t *pub = nullptr;   // mapped pinned: [slot][chunk][2] = {coder overflow flag, payload bytes}
};

// installs a lane's stream / scratch / work counter in the context for the duration of a launch sequence
struct UseLane
{
  dc_ctx *c;
  Lane *l;
  Lane saved;
  UseLane(dc_ctx *ctx, Lane *lane) : c(ctx), l(lane)
  {
    saved.stream = c->stream;
    saved.arena = c->arena;
    saved.arena_size = c->arena_size;
    saved.d_work = c->d_work;
    c->stream = l->stream;
    c->arena = l->arena;
    c->arena_size = l->arena_size;
    c->d_work = l->d_work;
  }
  ~UseLane()
  {
    l->arena = c->arena;   // may have grown
    l->arena_size = c->arena_size;
    c->stream = saved.stream;
    c->arena = saved.arena;
    c->arena_size = saved.arena_size;
    c->d_work = saved.d_work;
  }
};

}   // extern "C"

static int engine_create(dc_ctx *c)
{
  JobEngine *e = new (std::nothrow) JobEngine;
  c->async_jobs = new (std::nothrow) std::vector<HostJob *>;
  c->eng = e;
  if (!e || !c->async_jobs)
    return DC_ERR_MEMORY;
  memset(e->lane, 0, sizeof(e->lane));
  memset(e->ev_in, 0, sizeof(e->ev_in));
  memset(e->ev_k, 0, sizeof(e->ev_k));
  memset(e->ev_out, 0, sizeof(e->ev_out));
  bool ok = cudaHostAlloc(&e->pub, (size_t)kJobSlots * kMaxHostChunks * 2 * sizeof(uint64_t), cudaHostAllocMapped) == cudaSuccess;
  memset(e->dir, 0, sizeof(e->dir));
  e->dir[0].stream = c->s_h2d;
  e->dir[0].kind = cudaMemcpyHostToDevice;
  e->dir[1].stream = c->s_d2h;
  e->dir[1].kind = cudaMemcpyDeviceToHost;
  for (int d = 0; d < 2 && ok; d++)
    for (int i = 0; i < 8 && ok; i++)
      ok = cudaEventCreateWithFlags(&e->dir[d].ring[i], cudaEventDisableTiming) == cudaSuccess;
  for (int s = 0; s < kJobKinds && ok; s++)
    ok = cudaStreamCreateWithFlags(&e->lane[s].stream, cudaStreamNonBlocking) == cudaSuccess && cudaMalloc(&e->lane[s].d_work, 256) == cudaSuccess;
  for (int s = 0; s < kJobSlots && ok; s++)
  {
    for (int k = 0; k < kMaxHostChunks && ok; k++)
      ok = cudaEventCreateWithFlags(&e->ev_in[s][k], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&e->ev_k[s][k], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&e->ev_out[s][k], cudaEventDisableTiming) == cudaSuccess;
  }
  return ok ? DC_OK : DC_ERR_LIBRARY_INIT;
}

static void engine_destroy(dc_ctx *c)
{
  JobEngine *e = c->eng;
  if (e)
  {
    if (e->started)
    {
      {
        std::lock_guard<std::mutex> lk(e->mu);
        e->quit = true;
      }
      e->cv.notify_all();
      e->worker.join();   // the worker finishes every queued job first
    }
    for (int s = 0; s < kJobKinds; s++)
    {
      if (e->lane[s].stream)
      {
        cudaStreamSynchronize(e->lane[s].stream);
        cudaStreamDestroy(e->lane[s].stream);
      }
      if (e->lane[s].arena)
        cudaFree(e->lane[s].arena);
      if (e->lane[s].d_work)
        cudaFree(e->lane[s].d_work);
    }
    for (int s = 0; s < kJobSlots; s++)
    {
      if (e->stage[s])
        cudaFree(e->stage[s]);
      for (int k = 0; k < kMaxHostChunks; k++)
      {
        if (e->ev_in[s][k])
          cudaEventDestroy(e->ev_in[s][k]);
        if (e->ev_k[s][k])
          cudaEventDestroy(e->ev_k[s][k]);
        if (e->ev_out[s][k])
          cudaEventDestroy(e->ev_out[s][k]);
      }
    }
    for (int d = 0; d < 2; d++)
      for (int i = 0; i < 8; i++)
        if (e->dir[d].ring[i])
          cudaEventDestroy(e->dir[d].ring[i]);
    if (e->pub)
      cudaFreeHost(e->pub);
    delete e;
    c->eng = nullptr;
  }
  if (c->async_jobs)
  {
    for (HostJob *j : *c->async_jobs)
      delete j;
    delete c->async_jobs;
    c->async_jobs = nullptr;
  }
}

static int host_chunks(const Layout &L, uint64_t n_channels)
{
  if (L.sample_off != nullptr || n_channels < 2)
    return 1;
  uint64_t per = (uint64_t)sm_count() * (uint64_t)bac_threads_per_sm() / 2;   // streams per chunk: half a wave of the coder kernels
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
// slower in round 1: larger chunks in the middle cost more than the tail saves.)
static void chunk_bounds(int n_chunks, uint64_t n_channels, uint64_t *bounds)
{
  for (int k = 0; k <= n_chunks; k++)
    bounds[k] = n_channels * (uint64_t)k / (uint64_t)n_chunks;
}

static int job_stage_reserve(dc_ctx *c, int slot, uint64_t bytes)
{
  JobEngine *e = c->eng;
  if (bytes <= e->stage_size[slot])
    return DC_OK;
  if (e->stage[slot])
    cudaFree(e->stage[slot]);   // the slot is idle: nothing of it is in flight
  e->stage[slot] = nullptr;
  e->stage_size[slot] = 0;
  const uint64_t want = round_up(bytes, 1 << 20);
  if (cudaMalloc(&e->stage[slot], want) != cudaSuccess)
  {
    cudaGetLastError();
    set_err(c, "device staging allocation of %llu bytes failed", (unsigned long long)bytes);
    return DC_ERR_MEMORY;
  }
  e->stage_size[slot] = want;
  return DC_OK;
}

static void chunk_channels(const JobRun *r, int k, uint64_t &c0, uint64_t &c1)
{
  if (r->n_chunks == 1)
  {
    c0 = 0;
    c1 = r->job->layout.sample_off ? 0 : r->job->layout.n_channels;
    return;
  }
  c0 = r->bounds[k];
  c1 = r->bounds[k + 1];
}

// sets up a job in its (idle) slot: geometry, device staging, the small copies that precede every chunk
static int job_admit(dc_ctx *c, HostJob *j, JobRun *r, int slot)
{
  JobEngine *e = c->eng;
  CU_TRY(c, cudaSetDevice(c->device), DC_ERR_LIBRARY_INIT);
  r->job = j;
  r->slot = slot;
  r->age = e->admitted++;
  r->up = r->kn = r->fin = r->done = 0;
  r->up_issued = r->out_issued = 0;
  r->base = 0;
  r->result = DC_OK;
  r->finished = false;
  int rc = make_layout(c, &j->layout, j->layout.sample_off, r->L, r->max_len, r->total);
  if (rc != DC_OK)
    return rc;
  r->n = r->L.n_streams;
  r->bpc = r->L.blocks_per_channel;
  r->n_chunks = host_chunks(r->L, j->layout.n_channels);
  chunk_bounds(r->n_chunks, j->layout.sample_off ? 0 : j->layout.n_channels, r->bounds);
  for (int k = 0; k < r->n_chunks; k++)
    r->boost[k] = 1;
  const uint64_t n = r->n;
  uint64_t payload_dev = 0;
  if (j->kind == 0)
  {
    // every chunk gets its own region of the device payload (its start cannot depend on the sizes of the chunks before
    // it: those are only known when their kernels have finished): the caller's capacity, or the most the chunk's
    // samples can code to (65 SEG bits per sample at value size 32, 9/8 of that out of the coder), whichever is smaller
    uint64_t most_channels = 0;
    for (int k = 0; k < r->n_chunks; k++)
    {
      uint64_t c0, c1;
      chunk_channels(r, k, c0, c1);
      most_channels = std::max(most_channels, c1 - c0);
    }
    const uint64_t chunk_samples = r->n_chunks == 1 ? r->total : most_channels * r->L.channel_len;
    const uint64_t chunk_streams = r->n_chunks == 1 ? n : most_channels * r->bpc;
    const uint64_t worst = (chunk_samples * (2ull * j->params.value_size_bits + 1) * 13 / 8 + 7) / 8 + chunk_streams * 80 + 4096;
    r->region = round_up(std::min(j->payload_size, worst) + 64, 256);
    payload_dev = r->region * (uint64_t)r->n_chunks;
  }
  else
    payload_dev = round_up(j->payload_size, 4) + 16;
  const uint64_t need = Bump::size_of(r->total, 4) + Bump::size_of(payload_dev, 1) + 2 * Bump::size_of(n, 8) + 2 * Bump::size_of(n + 8, 8) +
                        Bump::size_of(n, 4) + 4096;
  if ((rc = job_stage_reserve(c, r->slot, need)) != DC_OK)
    return rc;
  Bump b(e->stage[r->slot]);
  r->d_samples = b.take<float>(r->total);
  r->d_payload = b.take<uint8_t>(payload_dev);
  r->d_bits = b.take<uint64_t>(n);
  r->d_nout = b.take<uint64_t>(n);
  r->d_boff = b.take<uint64_t>(n + 8);
  r->d_off = b.take<uint64_t>(n + 8);
  r->d_status = b.take<int32_t>(n);
  if (j->layout.sample_off)
  {
    CU_TRY(c, cudaMemcpyAsync(r->d_off, j->layout.sample_off, (n + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d), DC_ERR_LIBRARY_CALL);
    r->L.sample_off = r->d_off;
  }
  if (j->kind == 1)
  {
    // index first (small), the payload follows chunk by chunk at its container offset
    if (n)
      CU_TRY(c, cudaMemcpyAsync(r->d_bits, j->bit_len_in, n * 8, cudaMemcpyHostToDevice, c->s_h2d), DC_ERR_LIBRARY_CALL);
    CU_TRY(c, cudaMemcpyAsync(r->d_boff, j->byte_off_in, (n + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d), DC_ERR_LIBRARY_CALL);
  }
  trace(c, j->kind == 0 ? "encode job admitted, chunks" : "decode job admitted, chunks", r->n_chunks);
  return DC_OK;
}

static void chunk_streams(const JobRun *r, int k, uint64_t &s0, uint64_t &s1, uint64_t &first, uint64_t &count)
{
  uint64_t c0, c1;
  chunk_channels(r, k, c0, c1);
  if (r->n_chunks == 1)
  {
    s0 = 0;
    s1 = r->n;
    first = 0;
    count = r->total;
    return;
  }
  s0 = c0 * r->bpc;
  s1 = c1 * r->bpc;
  first = c0 * r->L.channel_len;
  count = (c1 - c0) * r->L.channel_len;
}

// cuts one copy into pieces; the last piece carries `after` (a copy of 0 bytes still carries it)
static void push_pieces(std::deque<Piece> &q, void *dst, const void *src, uint64_t bytes, cudaEvent_t after)
{
  uint8_t *d = static_cast<uint8_t *>(dst);
  const uint8_t *s = static_cast<const uint8_t *>(src);
  while (bytes > kPieceBytes)
  {
    q.push_back(Piece{ d, s, kPieceBytes, nullptr });
    d += kPieceBytes;
    s += kPieceBytes;
    bytes -= kPieceBytes;
  }
  q.push_back(Piece{ d, s, bytes, after });
}

// queues the kernels of chunk k of an encode job on its lane (after the chunk's upload)
static int encode_chunk_launch(dc_ctx *c, JobRun *r, int k)
{
  JobEngine *e = c->eng;
  HostJob *j = r->job;
  uint64_t s0, s1, first, count;
  chunk_streams(r, k, s0, s1, first, count);
  Layout Lk = r->L;
  Lk.n_streams = s1 - s0;
  UseLane use(c, &e->lane[r->slot / 2]);
  CU_TRY(c, cudaStreamWaitEvent(c->stream, e->ev_in[r->slot][k], 0), DC_ERR_LIBRARY_CALL);
  uint64_t *pub = e->pub + ((uint64_t)r->slot * kMaxHostChunks + k) * 2;
  const int rc = encode_pipeline(c, FRONT_DEGA, r->d_samples + first, Lk, r->max_len, &j->params, r->d_payload + (uint64_t)k * r->region, r->region,
                                 r->d_bits + s0, r->d_boff + s0, r->d_status + s0, nullptr, pub, r->boost[k]);
  if (rc != DC_OK)
    return rc;
  CU_TRY(c, cudaEventRecord(e->ev_k[r->slot][k], c->stream), DC_ERR_LIBRARY_CALL);
  return DC_OK;
}

// One pass over a job: queues whatever has become possible.  Returns true when something moved.
static bool job_step(dc_ctx *c, JobRun *r)
{
  JobEngine *e = c->eng;
  HostJob *j = r->job;
  const int slot = r->slot;
  bool moved = false;
  auto fail = [&](int rc) {
    // a CUDA call failed: let everything queued for this job drain, then report
    cudaStreamSynchronize(e->lane[slot / 2].stream);
    cudaStreamSynchronize(c->s_h2d);
    cudaStreamSynchronize(c->s_d2h);
    cudaGetLastError();
    r->result = rc;
    r->finished = true;
    return true;
  };
  // 1. uploads: at most two chunks between "upload decided" and "kernels finished"
  while (r->up < r->n_chunks && r->up < r->fin + 2)
  {
    const int k = r->up;
    uint64_t s0, s1, first, count;
    chunk_streams(r, k, s0, s1, first, count);
    if (j->kind == 0)
      push_pieces(r->h2d, r->d_samples + first, j->samples_in + first, count * 4, e->ev_in[slot][k]);
    else
    {
      const uint64_t b0 = j->byte_off_in[s0], b1 = j->byte_off_in[s1];
      if (b1 < b0 || b1 > j->payload_size)
        return fail(DC_ERR_INVALID_VALUE);
      push_pieces(r->h2d, r->d_payload + b0, j->payload_in + b0, b1 - b0, e->ev_in[slot][k]);
    }
    trace(c, j->kind == 0 ? "encode: upload decided, chunk" : "decode: upload decided, chunk", k);
    r->up++;
    moved = true;
  }
  // 2. kernels of every chunk whose upload has been issued (its event is recorded; stream order keeps the lane's scratch
  //    safe from chunk to chunk)
  while (r->kn < r->up_issued)
  {
    const int k = r->kn;
    int rc;
    if (j->kind == 0)
      rc = encode_chunk_launch(c, r, k);
    else
    {
      uint64_t s0, s1, first, count;
      chunk_streams(r, k, s0, s1, first, count);
      Layout Lk = r->L;
      Lk.n_streams = s1 - s0;
      UseLane use(c, &e->lane[slot / 2]);
      rc = cudaStreamWaitEvent(c->stream, e->ev_in[slot][k], 0) == cudaSuccess ? DC_OK : DC_ERR_LIBRARY_CALL;
      if (rc == DC_OK)
        rc = decode_pipeline(c, FRONT_DEGA, r->d_payload, j->payload_size, r->d_bits + s0, r->d_boff + s0, Lk, r->max_len, &j->params,
                             r->d_samples + first, r->d_nout + s0, r->d_status + s0);
      if (rc == DC_OK && cudaEventRecord(e->ev_k[slot][k], c->stream) != cudaSuccess)
        rc = DC_ERR_LIBRARY_CALL;
    }
    if (rc != DC_OK)
      return fail(rc);
    r->kn++;
    moved = true;
  }
  // 3. results of the chunks whose kernels have finished, in order: only now are their copies queued (a copy that had to
  //    wait for an event would hold up every copy queued behind it, the other job's included)
  while (r->fin < r->kn)
  {
    const int k = r->fin;
    const cudaError_t q = cudaEventQuery(e->ev_k[slot][k]);
    if (q == cudaErrorNotReady)
      break;
    if (q != cudaSuccess)
      return fail(DC_ERR_LIBRARY_CALL);
    uint64_t s0, s1, first, count;
    chunk_streams(r, k, s0, s1, first, count);
    const uint64_t nk = s1 - s0;
    if (j->kind == 0)
    {
      const volatile uint64_t *pub = e->pub + ((uint64_t)slot * kMaxHostChunks + k) * 2;
      const uint64_t overflow = pub[0], tot = pub[1];
      if (overflow != 0 && r->boost[k] < 512)
      {
        // a stream outgrew the coder scratch (never seen on load profiles): this chunk once more with a larger stride
        r->boost[k] *= 8;
        const int rc = encode_chunk_launch(c, r, k);
        if (rc != DC_OK)
          return fail(rc);
        moved = true;
        break;
      }
      if (overflow != 0)
      {
        set_err(c, "arithmetic coder output exceeded the scratch capacity");
        r->result = DC_ERR_MEMORY;
      }
      const uint64_t cap_left = j->payload_size > r->base ? j->payload_size - r->base : 0;
      if (tot > cap_left || tot > r->region)
      {
        set_err(c, "payload needs more than the capacity of %llu bytes", (unsigned long long)j->payload_size);
        r->result = DC_ERR_MEMORY;
      }
      if (j->bit_len_out && nk)
        push_pieces(r->d2h, j->bit_len_out + s0, r->d_bits + s0, nk * 8, nullptr);
      if (j->byte_off_out && nk)
        push_pieces(r->d2h, j->byte_off_out + s0, r->d_boff + s0, nk * 8, nullptr);
      if (j->status && nk)
        push_pieces(r->d2h, j->status + s0, r->d_status + s0, nk * 4, nullptr);
      const bool copy_payload = r->result == DC_OK && j->payload_out && tot;
      push_pieces(r->d2h, copy_payload ? j->payload_out + r->base : nullptr, r->d_payload + (uint64_t)k * r->region, copy_payload ? tot : 0,
                  e->ev_out[slot][k]);
      r->chunk_base[k] = r->base;
      r->chunk_first[k] = s0;
      r->chunk_count[k] = nk;
      r->base += tot;
    }
    else
    {
      if (j->n_out && nk)
        push_pieces(r->d2h, j->n_out + s0, r->d_nout + s0, nk * 8, nullptr);
      if (j->status && nk)
        push_pieces(r->d2h, j->status + s0, r->d_status + s0, nk * 4, nullptr);
      const bool copy_samples = j->samples_out && count;
      push_pieces(r->d2h, copy_samples ? j->samples_out + first : nullptr, r->d_samples + first, copy_samples ? count * 4 : 0, e->ev_out[slot][k]);
    }
    trace(c, j->kind == 0 ? "encode: kernels finished, results queued, chunk" : "decode: kernels finished, results queued, chunk", k);
    r->fin++;
    moved = true;
  }
  // 4. results landed
  while (r->done < r->out_issued)
  {
    const cudaError_t q = cudaEventQuery(e->ev_out[slot][r->done]);
    if (q == cudaErrorNotReady)
      break;
    if (q != cudaSuccess)
      return fail(DC_ERR_LIBRARY_CALL);
    r->done++;
    moved = true;
  }
  if (r->done == r->n_chunks && !r->finished)
  {
    if (j->kind == 0)
    {
      if (j->byte_off_out)
      {
        // chunk-local offsets become container offsets (chunk 0 already is)
        for (int k = 1; k < r->n_chunks; k++)
          for (uint64_t i = 0; i < r->chunk_count[k]; i++)
            j->byte_off_out[r->chunk_first[k] + i] += r->chunk_base[k];
        j->byte_off_out[r->n] = r->base;
      }
      if (j->total_bytes)
        *j->total_bytes = r->base;
    }
    trace(c, j->kind == 0 ? "encode job finished" : "decode job finished");
    r->finished = true;
    moved = true;
  }
  return moved;
}

// Hands pieces to one copy stream while fewer than kPiecesInFlight are in flight there.  Returns true when something moved.
static bool pump_copies(dc_ctx *c, int d)
{
  JobEngine *e = c->eng;
  CopyDir &dir = e->dir[d];
  bool moved = false;
  while (dir.count > 0 && cudaEventQuery(dir.ring[dir.head]) == cudaSuccess)
  {
    dir.head = (dir.head + 1) % 8;
    dir.count--;
    moved = true;
  }
  while (dir.count < kPiecesInFlight)
  {
    // whose piece: host -> device the decode's first, device -> host the kinds in turns; within a kind the older job's
    JobRun *r = nullptr;
    for (int t = 0; t < kJobKinds && !r; t++)
    {
      const int kind = d == 0 ? kJobKinds - 1 - t : (dir.turn + t) % kJobKinds;
      for (int b = 0; b < 2; b++)
      {
        JobRun *cand = e->run[2 * kind + b];
        if (cand && !cand->finished && !(d == 0 ? cand->h2d : cand->d2h).empty() && (!r || cand->age < r->age))
          r = cand;
      }
    }
    if (!r)
      break;
    std::deque<Piece> &q = d == 0 ? r->h2d : r->d2h;
    const Piece p = q.front();
    q.pop_front();
    bool ok = p.bytes == 0 || cudaMemcpyAsync(p.dst, p.src, p.bytes, dir.kind, dir.stream) == cudaSuccess;
    if (ok && p.after)
    {
      ok = cudaEventRecord(p.after, dir.stream) == cudaSuccess;
      (d == 0 ? r->up_issued : r->out_issued)++;
    }
    ok = ok && cudaEventRecord(dir.ring[(dir.head + dir.count) % 8], dir.stream) == cudaSuccess;
    dir.count++;
    if (d == 1)
      dir.turn = (r->slot / 2 + 1) % kJobKinds;
    if (!ok)
    {
      cudaGetLastError();
      r->result = DC_ERR_LIBRARY_CALL;
      r->h2d.clear();
      r->d2h.clear();
      cudaStreamSynchronize(e->lane[r->slot / 2].stream);
      cudaStreamSynchronize(c->s_h2d);
      cudaStreamSynchronize(c->s_d2h);
      r->finished = true;
    }
    moved = true;
  }
  return moved;
}

static void job_worker(dc_ctx *c)
{
  JobEngine *e = c->eng;
  cudaSetDevice(c->device);
  auto any_run = [&] {
    for (int s = 0; s < kJobSlots; s++)
      if (e->run[s])
        return true;
    return false;
  };
  // the slot a job of `kind` may take now, or -1: an idle kind takes its first slot; a second job joins when the one in
  // flight has issued all its uploads
  auto free_slot = [&](int kind) {
    JobRun *a = e->run[2 * kind], *b = e->run[2 * kind + 1];
    if (!a && !b)
      return 2 * kind;
    if (a && b)
      return -1;
    JobRun *busy = a ? a : b;
    if (busy->up_issued < busy->n_chunks)
      return -1;
    return a ? 2 * kind + 1 : 2 * kind;
  };
  std::unique_lock<std::mutex> lk(e->mu);
  for (;;)
  {
    e->cv.wait(lk, [&] { return e->quit || !e->queue.empty() || any_run(); });
    if (e->quit && e->queue.empty() && !any_run())
      return;
    // admit queued jobs, in submission order
    while (!e->queue.empty())
    {
      HostJob *j = e->queue.front();
      const int slot = free_slot(j->kind);
      if (slot < 0)
        break;
      e->queue.pop_front();
      lk.unlock();
      JobRun *r = new (std::nothrow) JobRun;
      const int rc = r ? job_admit(c, j, r, slot) : DC_ERR_MEMORY;
      lk.lock();
      if (rc == DC_OK)
        e->run[slot] = r;
      else
      {
        delete r;
        j->rc = rc;
        j->done = true;
        e->completed++;
        e->cv_done.notify_all();
      }
    }
    lk.unlock();
    bool moved = false;
    for (int kind = kJobKinds - 1; kind >= 0; kind--)   // the decode first: its uploads are small and what its long copy back waits for
    {
      JobRun *a = e->run[2 * kind], *b = e->run[2 * kind + 1];
      if (a && b && b->age < a->age)
        std::swap(a, b);   // the older job first: its kernels go on the lane ahead of the younger one's
      if (a && !a->finished)
        moved |= job_step(c, a);
      if (b && !b->finished)
        moved |= job_step(c, b);
    }
    moved |= pump_copies(c, 0);
    moved |= pump_copies(c, 1);
    lk.lock();
    for (int s = 0; s < kJobSlots; s++)
      if (e->run[s] && e->run[s]->finished)
      {
        HostJob *j = e->run[s]->job;
        j->rc = e->run[s]->result;
        j->done = true;
        delete e->run[s];
        e->run[s] = nullptr;
        e->completed++;
        e->cv_done.notify_all();
      }
    if (!moved && any_run())
    {
      lk.unlock();
      std::this_thread::sleep_for(std::chrono::microseconds(20));   // everything queued is still running
      lk.lock();
    }
  }
}

// hands a job to the worker; blocking: waits for it and returns its code
static int job_submit(dc_ctx *c, const HostJob &proto, bool blocking)
{
  JobEngine *e = c->eng;
  HostJob *j = new (std::nothrow) HostJob(proto);
  if (!j)
    return DC_ERR_MEMORY;
  j->rc = DC_OK;
  j->done = false;
  std::unique_lock<std::mutex> lk(e->mu);
  if (!e->started)
  {
    e->worker = std::thread(job_worker, c);
    e->started = true;
  }
  e->queue.push_back(j);
  e->submitted++;
  e->cv.notify_all();
  if (!blocking)
  {
    c->async_jobs->push_back(j);
    return DC_OK;
  }
  e->cv_done.wait(lk, [&] { return j->done; });
  const int rc = j->rc;
  delete j;
  return rc;
}

static int host_job(dc_ctx *c, bool blocking, int kind, const float *samples_in, float *samples_out, const uint8_t *payload_in, uint8_t *payload_out,
                    uint64_t payload_size, const uint64_t *bit_len_in, const uint64_t *byte_off_in, uint64_t *bit_len_out, uint64_t *byte_off_out,
                    const dc_layout *layout, const dc_params *params, int32_t *status, uint64_t *n_out, uint64_t *total_bytes)
{
  if (!c || !params_ok(params) || !layout || (kind == 1 && (!bit_len_in || !byte_off_in)))
    return DC_ERR_INVALID_VALUE;
  HostJob j;
  memset(&j, 0, sizeof(j));
  j.kind = kind;
  j.samples_in = samples_in;
  j.samples_out = samples_out;
  j.payload_in = payload_in;
  j.payload_out = payload_out;
  j.payload_size = payload_size;
  j.bit_len_in = bit_len_in;
  j.byte_off_in = byte_off_in;
  j.bit_len_out = bit_len_out;
  j.byte_off_out = byte_off_out;
  j.status = status;
  j.n_out = n_out;
  j.total_bytes = total_bytes;
  j.layout = *layout;
  j.params = *params;
  return job_submit(c, j, blocking);
}

extern "C" {

int dc_dega_encode_host(dc_ctx *c, const float *samples, const dc_layout *layout, const dc_params *params, uint8_t *payload,
                        uint64_t payload_cap, uint64_t *bit_len, uint64_t *byte_off, int32_t *status, uint64_t *total_bytes)
{
  return host_job(c, true, 0, samples, nullptr, nullptr, payload, payload_cap, nullptr, nullptr, bit_len, byte_off, layout, params, status, nullptr,
                  total_bytes);
}

int dc_dega_decode_host(dc_ctx *c, const uint8_t *payload, uint64_t payload_bytes, const uint64_t *bit_len, const uint64_t *byte_off,
                        const dc_layout *layout, const dc_params *params, float *samples, uint64_t *n_out, int32_t *status)
{
  return host_job(c, true, 1, nullptr, samples, payload, nullptr, payload_bytes, bit_len, byte_off, nullptr, nullptr, layout, params, status, n_out,
                  nullptr);
}

// Asynchronous forms: the job is queued and the call returns; dc_wait() blocks until every queued job has finished.
int dc_dega_encode_host_async(dc_ctx *c, const float *samples, const dc_layout *layout, const dc_params *params, uint8_t *payload,
                              uint64_t payload_cap, uint64_t *bit_len, uint64_t *byte_off, int32_t *status, uint64_t *total_bytes)
{
  return host_job(c, false, 0, samples, nullptr, nullptr, payload, payload_cap, nullptr, nullptr, bit_len, byte_off, layout, params, status, nullptr,
                  total_bytes);
}

int dc_dega_decode_host_async(dc_ctx *c, const uint8_t *payload, uint64_t payload_bytes, const uint64_t *bit_len, const uint64_t *byte_off,
                              const dc_layout *layout, const dc_params *params, float *samples, uint64_t *n_out, int32_t *status)
{
  return host_job(c, false, 1, nullptr, samples, payload, nullptr, payload_bytes, bit_len, byte_off, nullptr, nullptr, layout, params, status, n_out,
                  nullptr);
}

int dc_wait_next(dc_ctx *c)
{
  if (!c)
    return DC_ERR_INVALID_VALUE;
  if (c->async_jobs->empty())
    return DC_OK;
  JobEngine *e = c->eng;
  HostJob *j = c->async_jobs->front();
  c->async_jobs->erase(c->async_jobs->begin());
  std::unique_lock<std::mutex> lk(e->mu);
  e->cv_done.wait(lk, [&] { return j->done; });
  const int rc = j->rc;
  delete j;
  return rc;
}

int dc_wait(dc_ctx *c)
{
  if (!c)
    return DC_ERR_INVALID_VALUE;
  int rc = DC_OK;
  while (!c->async_jobs->empty())
  {
    const int r = dc_wait_next(c);
    if (rc == DC_OK)
      rc = r;   // the first failure, in submission order
  }
  return rc;
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
    case DC_STAGE_COPY: return (in_bits + 7) / 8 + 8;
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
  if (stage == DC_STAGE_COPY)
  {
    // Copy (copy.c:11-29) moves block_size_bits at a time and fails with ERROR_LIBRARY_CALL on a short last block
    // (READ_BITS_CHECKED, io_macros.h:13-27); a stream of whole blocks comes out unchanged.  Inside a chain that is one
    // device-to-device copy: the stage exists so that a command line that names it stays on the device.
    const uint64_t block = params->block_size_bits ? params->block_size_bits : 8;
    if (in_bits % block != 0)
      return DC_ERR_LIBRARY_CALL;
    const uint64_t nbytes = (in_bits + 7) / 8;
    *out_bits = in_bits;
    if (nbytes > out_cap)
    {
      set_err(c, "stage output needs %llu bytes, capacity is %llu", (unsigned long long)nbytes, (unsigned long long)out_cap);
      return DC_ERR_MEMORY;
    }
    if (nbytes)
    {
      CU_TRY(c, cudaMemcpyAsync(out, in, nbytes, cudaMemcpyDefault, c->stream), DC_ERR_LIBRARY_CALL);
      CU_TRY(c, cudaStreamSynchronize(c->stream), DC_ERR_LIBRARY_CALL);
    }
    return DC_OK;
  }

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
