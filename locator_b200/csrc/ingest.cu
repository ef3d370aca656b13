// K1/K2: genotype filter, 2-bit pack, row/column gathers, jacknife column replacement.
// Integer, HBM-bound, bit-exact against oracle/ingest_ref.py.
//
// Reference semantics (locator/locator.py): filter_snps :265-281 (allel count_alleles /
// is_biallelic / to_allele_counts()[:, :, 1]), split_train_test :303-307, bootstrap gather
// :651-653, jacknife replace :726-727, replace_md :258-261.
#include "common.cuh"

namespace loc {
thread_local std::string g_err;
std::atomic<long long> g_launches{0};

// ---------------------------------------------------------------------------------------------
// Site statistics: one warp per site, lanes stride over the N calls (char2 loads, coalesced).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_site_stats(const int8_t* __restrict__ gt, int64_t nvar, int64_t nsamp,
                                                    int min_mac, int32_t* __restrict__ n_alleles,
                                                    int32_t* __restrict__ alt_count, int32_t* __restrict__ n_missing,
                                                    uint8_t* __restrict__ keep) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t v = warp; v < nvar; v += nwarps) {
    const char2* row = reinterpret_cast<const char2*>(gt + v * nsamp * 2);
    unsigned seen[4] = {0u, 0u, 0u, 0u};  // allele indices 0..127
    int alt = 0, miss = 0;
    for (int64_t s = lane; s < nsamp; s += 32) {
      char2 c = row[s];
      int a0 = c.x, a1 = c.y;
      if (a0 >= 0) seen[a0 >> 5] |= 1u << (a0 & 31);
      if (a1 >= 0) seen[a1 >> 5] |= 1u << (a1 & 31);
      alt += (a0 == 1) + (a1 == 1);
      miss += (a0 < 0) | (a1 < 0);
    }
    int na = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) na += __popc(__reduce_or_sync(0xffffffffu, seen[q]));
    alt = __reduce_add_sync(0xffffffffu, alt);
    miss = __reduce_add_sync(0xffffffffu, miss);
    if (lane == 0) {
      if (n_alleles) n_alleles[v] = na;
      if (alt_count) alt_count[v] = alt;
      if (n_missing) n_missing[v] = miss;
      if (keep) keep[v] = (na == 2 && (min_mac == 1 || alt >= min_mac)) ? 1 : 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Pack: tile of 32 samples x 32 words (512 SNPs). Reads are coalesced over samples (char2 per
// call), the smem transpose makes the uint32 writes coalesced over words.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_sites(const int8_t* __restrict__ gt, int64_t nsamp,
                                                    const int64_t* __restrict__ site_idx, int64_t K,
                                                    uint32_t* __restrict__ packed, int64_t row_words) {
  __shared__ uint32_t tile[32][33];
  const int tx = threadIdx.x;  // 0..31
  const int ty = threadIdx.y;  // 0..7
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int64_t w0 = (int64_t)blockIdx.y * 32;
  const int64_t s = s0 + tx;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int wl = ty + 8 * r;
    const int64_t w = w0 + wl;
    uint32_t word = 0u;
    if (s < nsamp && w < row_words) {
#pragma unroll 4
      for (int q = 0; q < 16; ++q) {
        const int64_t k = w * 16 + q;
        if (k < K) {
          const int64_t v = site_idx[k];
          char2 c = reinterpret_cast<const char2*>(gt)[v * nsamp + s];
          uint32_t g = (uint32_t)(c.x == 1) + (uint32_t)(c.y == 1);
          word |= g << (2 * q);
        }
      }
    }
    tile[wl][tx] = word;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int sl = ty + 8 * r;
    const int64_t ss = s0 + sl;
    const int64_t w = w0 + tx;
    if (ss < nsamp && w < row_words) packed[ss * row_words + w] = tile[tx][sl];
  }
}

__global__ void k_patch_calls(uint32_t* __restrict__ packed, int64_t row_words, const int64_t* __restrict__ ks,
                              const int64_t* __restrict__ samp, const uint8_t* __restrict__ val, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t k = ks[i];
  uint32_t* wp = packed + samp[i] * row_words + (k >> 4);
  const int sh = 2 * (int)(k & 15);
  atomicAnd(wp, ~(3u << sh));
  atomicOr(wp, ((uint32_t)val[i] & 3u) << sh);
}

// uint8 [n][K] -> packed. One thread per (row, word).
__global__ void k_pack_counts(const uint8_t* __restrict__ counts, int64_t n, int64_t K,
                              uint32_t* __restrict__ packed, int64_t row_words) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (w >= row_words || r >= n) return;
  const uint8_t* src = counts + r * K + w * 16;
  uint32_t word = 0u;
  const int64_t rem = K - w * 16;
#pragma unroll
  for (int q = 0; q < 16; ++q)
    if (q < rem) word |= ((uint32_t)src[q] & 3u) << (2 * q);
  packed[r * row_words + w] = word;
}

__global__ void k_unpack_counts(const uint32_t* __restrict__ packed, int64_t n, int64_t K, int64_t row_words,
                                uint8_t* __restrict__ counts) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (k >= K || r >= n) return;
  counts[r * K + k] = (uint8_t)((packed[r * row_words + (k >> 4)] >> (2 * (k & 15))) & 3u);
}

__global__ void k_gather_rows(const uint4* __restrict__ in, int64_t row_vec, const int64_t* __restrict__ rows,
                              uint4* __restrict__ out) {
  const int64_t r = blockIdx.y;
  const uint4* src = in + rows[r] * row_vec;
  uint4* dst = out + r * row_vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row_vec; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// out[r][k] = in[r][cols[k]]: one thread per output word (16 random 2-bit reads from one row,
// which stays L1/L2 resident: a row is K/4 bytes).
__global__ void __launch_bounds__(256) k_gather_cols(const uint32_t* __restrict__ in, int64_t row_words_in,
                                                     const int64_t* __restrict__ cols, int64_t K_out,
                                                     uint32_t* __restrict__ out, int64_t row_words_out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (w >= row_words_out) return;
  const uint32_t* src = in + r * row_words_in;
  uint32_t word = 0u;
#pragma unroll 4
  for (int q = 0; q < 16; ++q) {
    const int64_t k = w * 16 + q;
    if (k < K_out) {
      const int64_t c = cols[k];
      word |= ((src[c >> 4] >> (2 * (c & 15))) & 3u) << (2 * q);
    }
  }
  out[r * row_words_out + w] = word;
}

__global__ void k_replace_cols(uint32_t* __restrict__ packed, int64_t n, int64_t row_words,
                               const int64_t* __restrict__ sites, int64_t nsites, const uint8_t* __restrict__ vals) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // sample row (fast -> coalesced vals)
  const int64_t i = blockIdx.y;
  if (r >= n || i >= nsites) return;
  const int64_t k = sites[i];
  uint32_t* wp = packed + r * row_words + (k >> 4);
  const int sh = 2 * (int)(k & 15);
  atomicAnd(wp, ~(3u << sh));
  atomicOr(wp, ((uint32_t)vals[i * n + r] & 3u) << sh);
}

}  // namespace loc

using namespace loc;

extern "C" {

int loc_abi_version(void) { return LOC_ABI_VERSION; }
const char* loc_last_error(void) { return loc::g_err.c_str(); }
int64_t loc_launch_count(void) { return (int64_t)loc::g_launches.load(); }

int loc_site_stats(const int8_t* d_gt, int64_t nvar, int64_t nsamp, int32_t min_mac, int32_t* d_n_alleles,
                   int32_t* d_alt_count, int32_t* d_n_missing, uint8_t* d_keep, void* stream) {
  LOC_CHECK(nvar >= 0 && nsamp > 0, "loc_site_stats: bad shape");
  if (nvar == 0) return 0;
  LOC_CHECK(d_gt != nullptr, "loc_site_stats: null genotype pointer");
  const int warps_per_block = 8;
  int64_t blocks = cdiv(nvar, warps_per_block);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_site_stats<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_gt, nvar, nsamp, min_mac, d_n_alleles,
                                                                 d_alt_count, d_n_missing, d_keep);
  LOC_LAUNCHED();
  return 0;
}

int loc_pack_sites(const int8_t* d_gt, int64_t nvar, int64_t nsamp, const int64_t* d_site_idx, int64_t K,
                   uint32_t* d_packed, int64_t row_words, void* stream) {
  (void)nvar;
  LOC_CHECK(nsamp > 0 && K >= 0, "loc_pack_sites: bad shape");
  LOC_CHECK(row_words >= cdiv(K, 16) && row_words % 4 == 0, "loc_pack_sites: row_words must be >= ceil(K/16) and a multiple of 4");
  if (row_words == 0) return 0;
  dim3 grid((unsigned)cdiv(nsamp, 32), (unsigned)cdiv(row_words, 32));
  LOC_CHECK(grid.y <= 65535, "loc_pack_sites: K too large for one launch");
  k_pack_sites<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(d_gt, nsamp, d_site_idx, K, d_packed, row_words);
  LOC_LAUNCHED();
  return 0;
}

int loc_patch_calls(uint32_t* d_packed, int64_t row_words, const int64_t* d_k, const int64_t* d_samp,
                    const uint8_t* d_val, int64_t n, void* stream) {
  if (n <= 0) return 0;
  k_patch_calls<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(d_packed, row_words, d_k, d_samp, d_val, n);
  LOC_LAUNCHED();
  return 0;
}

int loc_pack_counts(const uint8_t* d_counts, int64_t n, int64_t K, uint32_t* d_packed, int64_t row_words,
                    void* stream) {
  LOC_CHECK(row_words >= cdiv(K, 16) && row_words % 4 == 0, "loc_pack_counts: bad row_words");
  if (n <= 0 || row_words == 0) return 0;
  LOC_CHECK(n <= 65535, "loc_pack_counts: more than 65535 rows per call");
  dim3 grid((unsigned)cdiv(row_words, 256), (unsigned)n);
  k_pack_counts<<<grid, 256, 0, (cudaStream_t)stream>>>(d_counts, n, K, d_packed, row_words);
  LOC_LAUNCHED();
  return 0;
}

int loc_unpack_counts(const uint32_t* d_packed, int64_t n, int64_t K, int64_t row_words, uint8_t* d_counts,
                      void* stream) {
  if (n <= 0 || K <= 0) return 0;
  LOC_CHECK(n <= 65535, "loc_unpack_counts: more than 65535 rows per call");
  dim3 grid((unsigned)cdiv(K, 256), (unsigned)n);
  k_unpack_counts<<<grid, 256, 0, (cudaStream_t)stream>>>(d_packed, n, K, row_words, d_counts);
  LOC_LAUNCHED();
  return 0;
}

int loc_gather_rows(const uint32_t* d_in, int64_t row_words, const int64_t* d_rows, int64_t n_out,
                    uint32_t* d_out, void* stream) {
  LOC_CHECK(row_words % 4 == 0, "loc_gather_rows: row_words must be a multiple of 4");
  if (n_out <= 0 || row_words == 0) return 0;
  LOC_CHECK(n_out <= 65535, "loc_gather_rows: more than 65535 rows per call");
  const int64_t row_vec = row_words / 4;
  dim3 grid((unsigned)(cdiv(row_vec, 256) > 64 ? 64 : cdiv(row_vec, 256)), (unsigned)n_out);
  k_gather_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(d_in), row_vec, d_rows,
                                                        reinterpret_cast<uint4*>(d_out));
  LOC_LAUNCHED();
  return 0;
}

int loc_gather_cols(const uint32_t* d_in, int64_t n, int64_t row_words_in, const int64_t* d_cols, int64_t K_out,
                    uint32_t* d_out, int64_t row_words_out, void* stream) {
  LOC_CHECK(row_words_out >= cdiv(K_out, 16) && row_words_out % 4 == 0, "loc_gather_cols: bad row_words_out");
  if (n <= 0 || row_words_out == 0) return 0;
  LOC_CHECK(n <= 65535, "loc_gather_cols: more than 65535 rows per call");
  dim3 grid((unsigned)cdiv(row_words_out, 256), (unsigned)n);
  k_gather_cols<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, row_words_in, d_cols, K_out, d_out, row_words_out);
  LOC_LAUNCHED();
  return 0;
}

int loc_replace_cols(uint32_t* d_packed, int64_t n, int64_t row_words, const int64_t* d_sites, int64_t nsites,
                     const uint8_t* d_vals, void* stream) {
  if (n <= 0 || nsites <= 0) return 0;
  LOC_CHECK(nsites <= 65535 * 32, "loc_replace_cols: too many sites per call");
  // grid.y is limited to 65535: loop in slabs.
  for (int64_t off = 0; off < nsites; off += 65535) {
    const int64_t cnt = (nsites - off) < 65535 ? (nsites - off) : 65535;
    dim3 grid((unsigned)cdiv(n, 128), (unsigned)cnt);
    k_replace_cols<<<grid, 128, 0, (cudaStream_t)stream>>>(d_packed, n, row_words, d_sites + off, cnt,
                                                           d_vals + off * n);
    LOC_LAUNCHED();
  }
  return 0;
}

}  // extern "C"
