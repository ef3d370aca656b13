// Pre-arranged copies of the hidden kernels W_i[k][j] that the two hidden-stack kernels stream into
// shared memory (kept in sync with the canonical Keras-order buffer by k_hidden_update / reslice).
#pragma once
#include <stdint.h>

namespace loc {

// hidden.cu (CUDA cores, C = cluster size, Hc = H / C), reduction-index pairs interleaved for FFMA2:
//   fs[i-1][r][k/2][jl][k&1] = W_i[k][r*Hc + jl]      (forward slice of CTA r)
//   bs[i-1][r][j/2][il][j&1] = W_i[r*Hc + il][j]      (backward slice of CTA r, transposed)
__device__ __forceinline__ void store_sliced(float* fs, float* bs, int H, int Hc, int layer, int k, int j, float w) {
  const int64_t base = (int64_t)(layer - 1) * H * H;
  fs[base + (int64_t)(j / Hc) * H * Hc + (int64_t)(k / 2) * (2 * Hc) + 2 * (j % Hc) + (k & 1)] = w;
  bs[base + (int64_t)(k / Hc) * H * Hc + (int64_t)(j / 2) * (2 * Hc) + 2 * (k % Hc) + (j & 1)] = w;
}

// hidden_tc.cu (tcgen05, H = 256, 64-column groups): ready-made UMMA operand images
//   fs[i-1][cj = j/64] : MN-major, [2 chunks of 32 j][256 k rows][128 B], 32-byte atoms XOR (k & 3)
//   bs[i-1][ci = k/64] : K-major,  [8 chunks of 32 j][8 groups of 8 rows (k % 64)][128 B], 16-byte chunks XOR row
__device__ __forceinline__ void store_images(float* fs, float* bs, int layer, int k, int j, float w) {
  const int64_t base = (int64_t)(layer - 1) * 256 * 256;
  {
    const int cjj = j >> 6, jl = j & 63, c2 = jl >> 5, jj = jl & 31;
    const int off = c2 * (256 * 32) + k * 32 + ((((jj >> 3) ^ (k & 3)) << 3)) + (jj & 7);  // floats
    fs[base + (int64_t)cjj * 256 * 64 + off] = w;
  }
  {
    const int ci = k >> 6, il = k & 63, kc = j >> 5, jj = j & 31, g = il >> 3, rr = il & 7;
    const int off = kc * 2048 + g * 256 + rr * 32 + ((((jj >> 2) ^ rr) << 2)) + (jj & 3);  // floats
    bs[base + (int64_t)ci * 256 * 64 + off] = w;
  }
}

}  // namespace loc
