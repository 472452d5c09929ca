// container.cpp -- on-disk container of a batch encode (host code only; no CUDA calls).
//
// The reference has no container: one DCCLI run writes one stream to one file (DCIOLib/src/bit_file_buffer.c:310-333,
// DCCLI/src/cli.c:129-148).  A batch encode produces hundreds of thousands of streams, so they go into one file:
//
//   [0, 128)                      dc_container_header (little endian)
//   [index_offset, +8 * n)        bit_len[n_streams], uint64: exact bit count of every stream
//   [payload_offset, +payload)    the streams back to back; stream s occupies ceil(bit_len[s] / 8) bytes starting at the
//                                 exclusive prefix sum of those byte counts
//
// Every stream's byte range is byte-identical to the file the reference's DCCLI writes for the same samples and chain,
// so `dd`-extracting one range and running the reference's `decode ...` chain on it restores that stream
// (tests/test_container.py does exactly that with oracle/_ref/DCCLI).
// One exception: a stream of 0 bits occupies 0 bytes here, while DCCLI's file for it is the single byte 0x00 (its
// uninit always writes the byte register, bit_file_buffer.c:329-333,312-318).  Only A-XDR of an empty meter and LZMH
// of an empty or 403-byte block produce 0 bits (lzmh.c:169-173); DEGA never does (an empty input codes to 3 bits).
// The bit count in the index is what tells such a stream from a one-byte one, which the reference's file cannot.  Writers are position based (pwrite): the ranks
// of a multi-GPU job write their index and payload slices into the same file at the offsets the size exchange
// (shard.exchange_sizes) gave them, in any order.
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/dc_cuda.h"

static const char kMagic[8] = { 'D', 'C', 'B', '2', '0', '0', 0, 1 };

static int write_all(int fd, const void *buf, uint64_t n, uint64_t off)
{
  const uint8_t *p = static_cast<const uint8_t *>(buf);
  while (n)
  {
    const ssize_t w = pwrite(fd, p, n > (1u << 30) ? (1u << 30) : (size_t)n, (off_t)off);
    if (w < 0 && errno == EINTR)
      continue;
    if (w <= 0)
      return DC_ERR_FILE_IO;
    p += w;
    off += (uint64_t)w;
    n -= (uint64_t)w;
  }
  return DC_OK;
}

static int read_all(int fd, void *buf, uint64_t n, uint64_t off)
{
  uint8_t *p = static_cast<uint8_t *>(buf);
  while (n)
  {
    const ssize_t r = pread(fd, p, n > (1u << 30) ? (1u << 30) : (size_t)n, (off_t)off);
    if (r < 0 && errno == EINTR)
      continue;
    if (r <= 0)
      return DC_ERR_FILE_IO;
    p += r;
    off += (uint64_t)r;
    n -= (uint64_t)r;
  }
  return DC_OK;
}

// all comparisons are written so that no sum or product of header fields can wrap (the header may come from a corrupt file)
static bool header_ok(const dc_container_header *h)
{
  if (memcmp(h->magic, kMagic, 8) != 0 || h->version != DC_CONTAINER_VERSION || h->chain < DC_CHAIN_DEGA || h->chain > DC_CHAIN_LZMH)
    return false;
  if (h->index_offset < sizeof(dc_container_header) || h->payload_offset < h->index_offset)
    return false;
  if (h->n_streams > (h->payload_offset - h->index_offset) / 8)
    return false;
  return h->payload_bytes <= UINT64_MAX - h->payload_offset;
}

static bool range_ok(uint64_t first, uint64_t n, uint64_t limit) { return first <= limit && n <= limit - first; }

extern "C" {

int dc_container_init_header(dc_container_header *h, uint32_t chain, const dc_params *params, const dc_layout *layout, uint64_t n_streams,
                             uint64_t payload_bytes)
{
  if (!h || chain < DC_CHAIN_DEGA || chain > DC_CHAIN_LZMH)
    return DC_ERR_INVALID_VALUE;
  memset(h, 0, sizeof(*h));
  memcpy(h->magic, kMagic, 8);
  h->version = DC_CONTAINER_VERSION;
  h->chain = chain;
  if (params)
  {
    h->value_size_bits = params->value_size_bits;
    h->normalization_factor = params->normalization_factor;
    h->adaptive = params->adaptive;
  }
  if (layout)
  {
    h->n_channels = layout->n_channels;
    h->channel_len = layout->channel_len;
    h->block_len = layout->block_len;
  }
  h->n_streams = n_streams;
  h->payload_bytes = payload_bytes;
  h->index_offset = sizeof(dc_container_header);
  h->payload_offset = (h->index_offset + 8 * n_streams + 63) / 64 * 64;
  return DC_OK;
}

int dc_container_create(const char *path, const dc_container_header *h)
{
  if (!path || !h || !header_ok(h))
    return DC_ERR_INVALID_VALUE;
  const int fd = open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
  if (fd < 0)
    return DC_ERR_FILE_IO;
  int rc = write_all(fd, h, sizeof(*h), 0);
  if (rc == DC_OK && ftruncate(fd, (off_t)(h->payload_offset + h->payload_bytes)) != 0)
    rc = DC_ERR_FILE_IO;
  if (close(fd) != 0 && rc == DC_OK)
    rc = DC_ERR_FILE_IO;
  return rc;
}

int dc_container_write_shard(const char *path, const dc_container_header *h, uint64_t first_stream, uint64_t n, const uint64_t *bit_len,
                             uint64_t byte_base, const uint8_t *payload, uint64_t payload_bytes)
{
  if (!path || !h || !header_ok(h) || !range_ok(first_stream, n, h->n_streams) || !range_ok(byte_base, payload_bytes, h->payload_bytes) ||
      (n && !bit_len) || (payload_bytes && !payload))
    return DC_ERR_INVALID_VALUE;
  const int fd = open(path, O_WRONLY);
  if (fd < 0)
    return DC_ERR_FILE_IO;
  int rc = write_all(fd, bit_len, 8 * n, h->index_offset + 8 * first_stream);
  if (rc == DC_OK)
    rc = write_all(fd, payload, payload_bytes, h->payload_offset + byte_base);
  if (close(fd) != 0 && rc == DC_OK)
    rc = DC_ERR_FILE_IO;
  return rc;
}

int dc_container_read_header(const char *path, dc_container_header *h)
{
  if (!path || !h)
    return DC_ERR_INVALID_VALUE;
  const int fd = open(path, O_RDONLY);
  if (fd < 0)
    return DC_ERR_FILE_IO;
  int rc = read_all(fd, h, sizeof(*h), 0);
  struct stat st;
  if (rc == DC_OK && (!header_ok(h) || fstat(fd, &st) != 0 || (uint64_t)st.st_size < h->payload_offset + h->payload_bytes))
    rc = DC_ERR_INVALID_FORMAT;
  close(fd);
  return rc;
}

int dc_container_read_index(const char *path, const dc_container_header *h, uint64_t first_stream, uint64_t n, uint64_t *bit_len)
{
  if (!path || !h || !header_ok(h) || !range_ok(first_stream, n, h->n_streams) || (n && !bit_len))
    return DC_ERR_INVALID_VALUE;
  const int fd = open(path, O_RDONLY);
  if (fd < 0)
    return DC_ERR_FILE_IO;
  const int rc = read_all(fd, bit_len, 8 * n, h->index_offset + 8 * first_stream);
  close(fd);
  return rc;
}

int dc_container_read_payload(const char *path, const dc_container_header *h, uint64_t byte_off, uint64_t nbytes, uint8_t *dst)
{
  if (!path || !h || !header_ok(h) || !range_ok(byte_off, nbytes, h->payload_bytes) || (nbytes && !dst))
    return DC_ERR_INVALID_VALUE;
  const int fd = open(path, O_RDONLY);
  if (fd < 0)
    return DC_ERR_FILE_IO;
  const int rc = read_all(fd, dst, nbytes, h->payload_offset + byte_off);
  close(fd);
  return rc;
}

}   // extern "C"
