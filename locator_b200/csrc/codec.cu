// Host-side chunk decoder for zarr stores written by scikit-allel's vcf_to_zarr (the reference's
// --zarr / --windows input, locator/locator.py:187-194, scripts/vcf_to_zarr.py:12): zarr's default
// compressor is Blosc (LZ4, byte shuffle).  Blosc 1.x frame format and the LZ4 block format are
// restated from their published specifications; no third-party code.  Pure host code (no CUDA).
#include <dlfcn.h>
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
    // short literal runs dominate genotype chunks: one unconditional 16-byte move when both buffers have the
    // room (the surplus bytes are overwritten by what follows), a real memcpy otherwise
    if (lit <= 16 && iend - ip >= 16 && oend - op >= 16) {
      memcpy(op, ip, 16);
    } else {
      memcpy(op, ip, (size_t)lit);
    }
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
    if (off >= 8 && oend - op >= ml + 8) {
      // 8-byte steps: every read lies at least 8 bytes behind its write, so it only sees finished bytes; the
      // last step may write up to 7 bytes past the match, inside the buffer, overwritten by what follows
      for (int64_t i = 0; i < ml; i += 8) memcpy(op + i, m + i, 8);
    } else if (off >= ml) {
      memcpy(op, m, (size_t)ml);
    } else {
      // overlapping match = the last `off` bytes repeated (genotype chunks are mostly runs: off = 1, ml in the
      // hundreds).  Copy from the start of the pattern in pieces that double: the bytes written so far stay a
      // whole number of periods, so source [m, m + c) never reaches the destination and has the right phase.
      int64_t filled = 0;
      while (filled < ml) {
        const int64_t c = (off + filled) < (ml - filled) ? (off + filled) : (ml - filled);
        memcpy(op + filled, m, (size_t)c);
        filled += c;
      }
    }
    op += ml;
  }
  return op == oend;
}

// Blosc frames whose codec is Zstandard (stores written with zarr.Blosc(cname="zstd")): the streams inside
// are ordinary zstd frames.  The system's libzstd is bound at first use (its two entry points have had this
// ABI since zstd 1.0); nothing is needed at build time and LZ4 stores never touch it.
typedef size_t (*zstd_decompress_fn)(void*, size_t, const void*, size_t);
typedef unsigned (*zstd_iserror_fn)(size_t);
struct Zstd {
  zstd_decompress_fn decompress = nullptr;
  zstd_iserror_fn is_error = nullptr;
  Zstd() {
    for (const char* name : {"libzstd.so.1", "libzstd.so"}) {
      void* h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h == nullptr) continue;
      decompress = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
      is_error = (zstd_iserror_fn)dlsym(h, "ZSTD_isError");
      if (decompress != nullptr && is_error != nullptr) return;
      decompress = nullptr;
    }
  }
};
const Zstd& zstd() {
  static const Zstd z;  // thread-safe one-time initialisation
  return z;
}

}  // namespace

extern "C" {

// One plain Zstandard frame (zarr compressor {"id": "zstd"}) into dst; returns the decompressed size or < 0.
int64_t loc_zstd_decompress(const uint8_t* src, int64_t src_len, uint8_t* dst, int64_t dst_len) {
  if (src == nullptr || dst == nullptr || src_len <= 0 || dst_len < 0) {
    loc::fail("loc_zstd_decompress: bad arguments", __FILE__, __LINE__);
    return -1;
  }
  if (zstd().decompress == nullptr) {
    loc::fail("loc_zstd_decompress: libzstd.so.1 could not be loaded", __FILE__, __LINE__);
    return -5;
  }
  const size_t got = zstd().decompress(dst, (size_t)dst_len, src, (size_t)src_len);
  if (zstd().is_error(got)) {
    loc::fail("loc_zstd_decompress: corrupt Zstandard stream", __FILE__, __LINE__);
    return -4;
  }
  return (int64_t)got;
}

// Decompress one Blosc-1 frame (LZ4 / LZ4HC / Zstandard codec or stored, byte shuffle or none) into dst.
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
  if (codec != 1 && codec != 4) {  // 0 blosclz, 1 lz4 / lz4hc, 2 snappy, 3 zlib, 4 zstd
    loc::fail("loc_blosc_decompress: only the LZ4 and Zstandard codecs of Blosc are supported", __FILE__, __LINE__);
    return -3;
  }
  if (codec == 4 && zstd().decompress == nullptr) {
    loc::fail("loc_blosc_decompress: Zstandard frame, but libzstd.so.1 could not be loaded", __FILE__, __LINE__);
    return -5;
  }
  // the byte transpose of 1-byte items (int8 calldata/GT, by far the largest array) is the identity
  const bool shuffle = (flags & 0x1) != 0 && typesize > 1, dont_split = (flags & 0x10) != 0;
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
      } else if (codec == 4) {
        const size_t got = zstd().decompress(out + sp * neblock, (size_t)neblock, src + pos, (size_t)cb);
        if (zstd().is_error(got) || (int64_t)got != neblock) {
          loc::fail("loc_blosc_decompress: corrupt Zstandard stream", __FILE__, __LINE__);
          return -4;
        }
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

// ---------------------------------------------------------------------------------------------
// VCF text -> GT cube (the allel.read_vcf(...)['calldata/GT'] of load_genotypes, locator.py:195-199):
// int8 [n_variants][n_samples][2], missing allele = -1, haploid call -> second allele -1, plus POS.
// Data lines are indexed once, then parsed by n_threads host threads over disjoint line ranges.
// ---------------------------------------------------------------------------------------------
#include <thread>

namespace {

struct VcfLines {
  std::vector<int64_t> beg, end;  // [beg, end) of every data line, '\r' stripped
};

void vcf_index(const char* buf, int64_t len, VcfLines& L) {
  int64_t p = 0;
  while (p < len) {
    const char* nl = (const char*)memchr(buf + p, '\n', (size_t)(len - p));
    int64_t e = nl ? (nl - buf) : len;
    int64_t ee = e;
    if (ee > p && buf[ee - 1] == '\r') --ee;
    if (ee > p && buf[p] != '#') {
      L.beg.push_back(p);
      L.end.push_back(ee);
    }
    p = e + 1;
  }
}

// one allele token [s, e): "." / "" -> -1, digits -> value; anything else -> error
inline bool parse_allele(const char* s, const char* e, int8_t& out) {
  if (s == e || (e - s == 1 && *s == '.')) {
    out = -1;
    return true;
  }
  int v = 0;
  for (const char* c = s; c < e; ++c) {
    if (*c < '0' || *c > '9') return false;
    v = v * 10 + (*c - '0');
    if (v > 127) return false;
  }
  out = (int8_t)v;
  return true;
}

bool vcf_parse_line(const char* s, const char* e, int64_t n_samples, int8_t* gt, int64_t* pos) {
  // fields: CHROM POS ID REF ALT QUAL FILTER INFO FORMAT sample...
  const char* f = s;
  const char* fld[10];
  int nf = 0;
  fld[nf++] = f;
  while (nf < 10) {
    const char* t = (const char*)memchr(f, '\t', (size_t)(e - f));
    if (!t) break;
    f = t + 1;
    fld[nf++] = f;
  }
  if (nf < 9) return false;
  {  // POS
    int64_t v = 0;
    const char* c = fld[1];
    if (c >= fld[2] - 1) return false;
    for (; c < fld[2] - 1; ++c) {
      if (*c < '0' || *c > '9') return false;
      v = v * 10 + (*c - '0');
    }
    *pos = v;
  }
  // index of the GT key in FORMAT
  const char* fmt_e = nf >= 10 ? fld[9] - 1 : e;
  int gi = -1, k = 0;
  for (const char* c = fld[8]; c <= fmt_e;) {
    const char* t = (const char*)memchr(c, ':', (size_t)(fmt_e - c));
    const char* ke = t ? t : fmt_e;
    if (ke - c == 2 && c[0] == 'G' && c[1] == 'T') {
      gi = k;
      break;
    }
    if (!t) break;
    c = t + 1;
    ++k;
  }
  if (gi < 0) return false;
  for (int64_t i = 0; i < 2 * n_samples; ++i) gt[i] = -1;
  if (nf < 10) return true;
  const char* c = fld[9];
  for (int64_t smp = 0; smp < n_samples && c <= e; ++smp) {
    const char* t = (const char*)memchr(c, '\t', (size_t)(e - c));
    const char* se = t ? t : e;
    // gi-th colon-separated subfield
    const char* a = c;
    for (int q = 0; q < gi && a; ++q) {
      const char* col = (const char*)memchr(a, ':', (size_t)(se - a));
      a = col ? col + 1 : nullptr;
    }
    if (a) {
      const char* col = (const char*)memchr(a, ':', (size_t)(se - a));
      const char* ae = col ? col : se;
      // alleles separated by '/' or '|' (the first two are kept)
      const char* sep = a;
      while (sep < ae && *sep != '/' && *sep != '|') ++sep;
      if (!parse_allele(a, sep, gt[2 * smp])) return false;
      if (sep < ae) {
        const char* b = sep + 1;
        const char* sep2 = b;
        while (sep2 < ae && *sep2 != '/' && *sep2 != '|') ++sep2;
        if (!parse_allele(b, sep2, gt[2 * smp + 1])) return false;
      }
    } else {
      return false;  // the sample field has fewer subfields than FORMAT promises
    }
    if (!t) break;
    c = t + 1;
  }
  return true;
}

}  // namespace

extern "C" {

// Number of data lines (variants) in an uncompressed VCF text buffer.
int64_t loc_vcf_count(const char* h_buf, int64_t len) {
  if (h_buf == nullptr || len < 0) return -1;
  VcfLines L;
  vcf_index(h_buf, len, L);
  return (int64_t)L.beg.size();
}

// GT int8 [n_variants][n_samples][2] and POS int64 [n_variants] of every data line.  Returns 0, or 1 with
// loc_last_error set when a line cannot be parsed (the caller may fall back to a slower, more lenient reader).
int loc_vcf_parse_gt(const char* h_buf, int64_t len, int64_t n_samples, int64_t n_variants, int8_t* h_gt, int64_t* h_pos,
                     int32_t n_threads) {
  LOC_CHECK(h_buf != nullptr && h_gt != nullptr && h_pos != nullptr && n_samples >= 0 && n_variants >= 0,
            "loc_vcf_parse_gt: bad arguments");
  VcfLines L;
  vcf_index(h_buf, len, L);
  LOC_CHECK((int64_t)L.beg.size() == n_variants, "loc_vcf_parse_gt: n_variants does not match the buffer (loc_vcf_count)");
  int nt = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
  if (n_variants < 4 * nt) nt = 1;
  std::vector<int> bad((size_t)nt, 0);
  auto work = [&](int t) {
    const int64_t v0 = n_variants * t / nt, v1 = n_variants * (t + 1) / nt;
    for (int64_t v = v0; v < v1; ++v)
      if (!vcf_parse_line(h_buf + L.beg[(size_t)v], h_buf + L.end[(size_t)v], n_samples, h_gt + v * n_samples * 2, h_pos + v)) {
        bad[(size_t)t] = 1;
        return;
      }
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  for (int b : bad) LOC_CHECK(b == 0, "loc_vcf_parse_gt: a data line could not be parsed (no GT key, bad POS or allele)");
  return 0;
}

}  // extern "C"
