// Host-side chunk decoder for zarr stores written by scikit-allel's vcf_to_zarr (the reference's
// --zarr / --windows input, locator/locator.py:187-194, scripts/vcf_to_zarr.py:12): zarr's default
// compressor is Blosc (LZ4, byte shuffle).  Blosc 1.x frame format and the LZ4 block format are
// restated from their published specifications; no third-party code.  Pure host code (no CUDA).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// LZ4 block: sequences of [token][literal length ext][literals][offset:2][match length ext]
bool lz4_block(const uint8_t* src, int64_t slen, uint8_t* dst, int64_t dlen) {
  const uint8_t* ip = src;
  const uint8_t* iend = src + slen;
  uint8_t* op = dst;
  uint8_t* oend = dst + dlen;
  while (ip < iend) {
    const unsigned token = *ip++;
    int64_t lit = token >> 4;
    if (lit == 15) {
      unsigned b;
      do {
        if (ip >= iend) return false;
        b = *ip++;
        lit += b;
      } while (b == 255);
    }
    if (lit > iend - ip || lit > oend - op) return false;
    memcpy(op, ip, (size_t)lit);
    ip += lit;
    op += lit;
    if (ip >= iend) break;  // the last sequence has literals only
    if (iend - ip < 2) return false;
    const int64_t off = (int64_t)ip[0] | ((int64_t)ip[1] << 8);
    ip += 2;
    if (off == 0 || off > op - dst) return false;
    int64_t ml = (token & 15) + 4;
    if ((token & 15) == 15) {
      unsigned b;
      do {
        if (ip >= iend) return false;
        b = *ip++;
        ml += b;
      } while (b == 255);
    }
    if (ml > oend - op) return false;
    const uint8_t* m = op - off;
    for (int64_t i = 0; i < ml; ++i) op[i] = m[i];  // overlapping copies replicate the pattern
    op += ml;
  }
  return op == oend;
}

}  // namespace

extern "C" {

// Decompress one Blosc-1 frame (LZ4 / LZ4HC codec or stored, byte shuffle or none) into dst.
// Returns the number of bytes written (the frame's nbytes) or a negative value on error.
int64_t loc_blosc_decompress(const uint8_t* src, int64_t src_len, uint8_t* dst, int64_t dst_len) {
  if (src == nullptr || dst == nullptr || src_len < 16) {
    loc::fail("loc_blosc_decompress: truncated frame", __FILE__, __LINE__);
    return -1;
  }
  const unsigned flags = src[2];
  const int64_t typesize = src[3] ? src[3] : 1;
  const int64_t nbytes = rd32(src + 4), blocksize = rd32(src + 8), cbytes = rd32(src + 12);
  if (nbytes > dst_len || cbytes > src_len || (nbytes > 0 && blocksize <= 0)) {
    loc::fail("loc_blosc_decompress: frame sizes do not match the buffers", __FILE__, __LINE__);
    return -1;
  }
  if (nbytes == 0) return 0;
  if (flags & 0x2) {  // stored (memcpy) frame
    if (16 + nbytes > src_len) return -1;
    memcpy(dst, src + 16, (size_t)nbytes);
    return nbytes;
  }
  if (flags & 0x4) {
    loc::fail("loc_blosc_decompress: bit-shuffled frames are not supported", __FILE__, __LINE__);
    return -2;
  }
  const unsigned codec = flags >> 5;
  if (codec != 1) {  // 0 blosclz, 1 lz4 / lz4hc, 2 snappy, 3 zlib, 4 zstd
    loc::fail("loc_blosc_decompress: only the LZ4 codec of Blosc is supported", __FILE__, __LINE__);
    return -3;
  }
  const bool shuffle = (flags & 0x1) != 0, dont_split = (flags & 0x10) != 0;
  const int64_t nblocks = (nbytes + blocksize - 1) / blocksize;
  if (16 + 4 * nblocks > src_len) return -1;
  std::vector<uint8_t> tmp((size_t)blocksize);
  for (int64_t b = 0; b < nblocks; ++b) {
    const int64_t bsize = (b == nblocks - 1 && nbytes % blocksize) ? nbytes % blocksize : blocksize;
    const bool leftover = bsize != blocksize;
    const int64_t nsplits =
        (!dont_split && !leftover && typesize <= 16 && blocksize / typesize >= 128) ? typesize : 1;
    const int64_t neblock = bsize / nsplits;
    int64_t pos = rd32(src + 16 + 4 * b);
    uint8_t* out = shuffle ? tmp.data() : dst + b * blocksize;
    for (int64_t sp = 0; sp < nsplits; ++sp) {
      if (pos + 4 > src_len) return -1;
      const int64_t cb = rd32(src + pos);
      pos += 4;
      if (cb < 0 || pos + cb > src_len) return -1;
      if (cb == neblock) {
        memcpy(out + sp * neblock, src + pos, (size_t)neblock);
      } else if (!lz4_block(src + pos, cb, out + sp * neblock, neblock)) {
        loc::fail("loc_blosc_decompress: corrupt LZ4 stream", __FILE__, __LINE__);
        return -4;
      }
      pos += cb;
    }
    if (shuffle) {  // undo the byte transpose: shuffled[j * n + i] = original[i * typesize + j]
      uint8_t* d = dst + b * blocksize;
      const int64_t n = bsize / typesize;
      for (int64_t j = 0; j < typesize; ++j)
        for (int64_t i = 0; i < n; ++i) d[i * typesize + j] = tmp[(size_t)(j * n + i)];
      memcpy(d + n * typesize, tmp.data() + n * typesize, (size_t)(bsize - n * typesize));
    }
  }
  return nbytes;
}

}  // extern "C"
