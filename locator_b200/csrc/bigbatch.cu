// Optimizer steps of 33..256 rows (--batch_size > 32: locator/locator.py:69,371 pass any batch size to model.fit).
//
// The fused tensor-core kernels of the default path hold ONE 32-row batch tile (l1_tc.cu, hidden_tc.cu).  Only the
// BatchNormalization in front of the first Dense couples the rows of a batch; everything behind it is row-wise, and
// every gradient is a sum over rows.  A larger step is therefore run as
//
//   k_bb_stats         batch mean / variance of every SNP over ALL rows of the step (+ moving statistics)
//   first-layer forward with those statistics (the inference kernels read them in place of the moving ones:
//                      one wide pass over W1 for up to 256 rows on the tcgen05 path), then the 32-row chunks
//                      through the unchanged hidden stack, one launch each: loss scaled by the step's row count,
//                      dropout stream indexed by the row's position in the step, activations / dz kept per chunk
//   k_bb_l1_bwd        S = (x - mean)^T dZ1 over all rows, dW1 = inv S + beta c0, Adam on W1 | m | v in one pass
//                      over the weights, BatchNorm gamma / beta Adam -- fp32 on the CUDA cores, either W1 layout
//   k_bb_hidden_update dW + Adam of the small layers summed over the chunks
//
// so W1 | m | v are still streamed once per optimizer step.  Keras semantics as restated in oracle/model_ref.py
// (RefLocator.gradients / train_step); algebra of the first layer as in l1_simt.cu.
#include "model.cuh"
#include "hidden_slices.cuh"

namespace loc {

constexpr int kBbT = 16;  // SNPs per chunk of the backward = one packed word per row

__device__ __forceinline__ void bb_moments(int n1, int n2, int nb, float& mean, float& var) {
  // tf.nn.moments: mean, then mean of squared differences (as l1_simt.cu: batch_moments)
  const float fn = (float)nb;
  const int n0 = nb - n1 - n2;
  mean = (float)(n1 + 2 * n2) / fn;
  const float d0 = 0.f - mean, d1 = 1.f - mean, d2 = 2.f - mean;
  var = ((float)n0 * d0 * d0 + (float)n1 * d1 * d1 + (float)n2 * d2 * d2) / fn;
}

// One thread per packed word (16 SNPs): genotype counts over the step's rows -> batch statistics.
__global__ void __launch_bounds__(128) k_bb_stats(BigArgs a) {
  if (a.gated && a.st->stopped) return;
  __shared__ int64_t s_rows[LOC_MAX_BATCH_SIZE];
  const int nb = a.nb;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) s_rows[b] = row_of(a.src, a.st, b);
  __syncthreads();
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 16 >= a.K) return;
  unsigned n1[16], n2[16];  // rows with one / two alternate alleles, per SNP of the word
#pragma unroll
  for (int t = 0; t < 16; ++t) n1[t] = n2[t] = 0u;
#pragma unroll 4
  for (int b = 0; b < nb; ++b) {
    const uint32_t x = __ldg(a.packed + s_rows[b] * a.row_words + w);
    const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
    const uint32_t is1 = lo & ~hi, is2 = hi & ~lo;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      n1[t] += (is1 >> (2 * t)) & 1u;
      n2[t] += (is2 >> (2 * t)) & 1u;
    }
  }
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int64_t k = w * 16 + t;
    if (k < a.K) {
      float mean, var;
      bb_moments((int)n1[t], (int)n2[t], nb, mean, var);
      a.bmean[k] = mean;
      a.bvar[k] = var;
      a.mmean[k] = a.mmean[k] * kBnMom + mean * kBnOneMinusMom;
      a.mvar[k] = a.mvar[k] * kBnMom + var * kBnOneMinusMom;
    }
  }
}

__device__ __forceinline__ const float* bb_dz1(const BigArgs& a, int b) {  // dZ1 row b of the step
  return a.dzs + ((int64_t)(b >> 5) * a.L * kMaxB + (b & 31)) * a.H;
}

// blockDim = max(H, 64); thread j <-> output column j.  A block walks a contiguous range of 16-SNP words.
__global__ void __launch_bounds__(1024) k_bb_l1_bwd(BigArgs a) {
  if (a.gated && a.st->stopped) return;
  extern __shared__ __align__(16) float bb_smem[];
  __shared__ int64_t s_rows[LOC_MAX_BATCH_SIZE];
  __shared__ float sc[kBbT][4];  // mean, inv, beta, rs
  const int H = a.H, nb = a.nb, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int Bp = (nb + 3) & ~3;
  float* xs = bb_smem;           // [kBbT][Bp] centred genotypes (0 beyond nb)
  float* pq = xs + kBbT * Bp;    // [nwarps][kBbT][2] per-warp partial P, Q
  for (int b = tid; b < nb; b += blockDim.x) s_rows[b] = row_of(a.src, a.st, b);
  const float alpha = a.st->alpha;
  const bool col = tid < H;
  float c0 = 0.f;
  if (col)
    for (int b = 0; b < nb; ++b) c0 += bb_dz1(a, b)[tid];
  __syncthreads();

  const int64_t nwords = (a.K + kBbT - 1) / kBbT;
  const int64_t c_begin = nwords * blockIdx.x / gridDim.x, c_end = nwords * (blockIdx.x + 1) / gridDim.x;
  for (int64_t c = c_begin; c < c_end; ++c) {
    const int64_t k0 = c * kBbT;
    const int tmax = (int)((a.K - k0) < kBbT ? (a.K - k0) : kBbT);
    if (tid < kBbT) {
      float mean = 0.f, inv = 0.f, beta = 0.f, rs = 0.f;
      if (tid < tmax) {
        mean = a.bmean[k0 + tid];
        rs = rsqrtf(a.bvar[k0 + tid] + kBnEps);
        inv = rs * a.gamma[k0 + tid];
        beta = a.beta[k0 + tid];
      }
      sc[tid][0] = mean;
      sc[tid][1] = inv;
      sc[tid][2] = beta;
      sc[tid][3] = rs;
    }
    __syncthreads();
    for (int b = tid; b < Bp; b += blockDim.x) {
      const uint32_t x = b < nb ? __ldg(a.packed + s_rows[b] * a.row_words + c) : 0u;
#pragma unroll
      for (int t = 0; t < kBbT; ++t) xs[t * Bp + b] = (b < nb && t < tmax) ? (float)((x >> (2 * t)) & 3u) - sc[t][0] : 0.f;
    }
    __syncthreads();
    float acc[kBbT];
#pragma unroll
    for (int t = 0; t < kBbT; ++t) acc[t] = 0.f;
    if (col) {
      for (int b0 = 0; b0 < Bp; b0 += 4) {
        float d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] = (b0 + e) < nb ? __ldg(bb_dz1(a, b0 + e) + tid) : 0.f;
#pragma unroll
        for (int t = 0; t < kBbT; ++t) {
          const float4 x4 = *reinterpret_cast<const float4*>(xs + t * Bp + b0);
          acc[t] = fmaf(x4.x, d[0], acc[t]);
          acc[t] = fmaf(x4.y, d[1], acc[t]);
          acc[t] = fmaf(x4.z, d[2], acc[t]);
          acc[t] = fmaf(x4.w, d[3], acc[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < kBbT; ++t) {
      float p = 0.f, q = 0.f;
      if (col && t < tmax) {
        const int64_t idx = a.tiled ? w1_tiled_index(k0 + t, tid) : (k0 + t) * H + tid;
        float w = a.W1[idx], m = a.mW1[idx], v = a.vW1[idx];
        const float S = acc[t];
        const float g = sc[t][1] * S + sc[t][2] * c0;
        p = w * S;
        q = w * c0;
        adam_update(w, m, v, g, alpha);
        a.W1[idx] = w;
        a.mW1[idx] = m;
        a.vW1[idx] = v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        p += __shfl_xor_sync(0xffffffffu, p, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0) {
        pq[(warp * kBbT + t) * 2] = p;
        pq[(warp * kBbT + t) * 2 + 1] = q;
      }
    }
    __syncthreads();
    if (tid < tmax) {
      float P = 0.f, Q = 0.f;
      for (int w = 0; w < nwarps; ++w) {
        P += pq[(w * kBbT + tid) * 2];
        Q += pq[(w * kBbT + tid) * 2 + 1];
      }
      const int64_t k = k0 + tid;
      const float dgamma = sc[tid][3] * P;  // exactly 0 for a SNP that is constant in the batch (centred genotypes)
      const float dbeta = Q;
      float gm = a.gamma[k], m = a.m_gamma[k], v = a.v_gamma[k];
      adam_update(gm, m, v, dgamma, alpha);
      a.gamma[k] = gm;
      a.m_gamma[k] = m;
      a.v_gamma[k] = v;
      float bt = a.beta[k];
      m = a.m_beta[k];
      v = a.v_beta[k];
      adam_update(bt, m, v, dbeta, alpha);
      a.beta[k] = bt;
      a.m_beta[k] = m;
      a.v_beta[k] = v;
    }
    __syncthreads();
  }
}

// dW + Adam of the small layers over the chunks of the step.  Blocks [0, (L-1)*H/16): 16 input rows x H outputs of
// hidden layer i; last block: b1, Dense(2), Dense(2).  blockDim = H, thread <-> output column j (as k_hidden_update).
constexpr int kBbRows = 16;

__global__ void __launch_bounds__(1024) k_bb_hidden_update(BigArgs a) {
  if (a.gated && a.st->stopped) return;
  __shared__ float as[kBbRows][kMaxB + 1];
  const int H = a.H, L = a.L, j = threadIdx.x;
  const SmallLayout sl{H, L};
  const int rb_n = H / kBbRows;
  const int nblk_hidden = (L - 1) * rb_n;
  const int nchunks = (a.nb + kMaxB - 1) / kMaxB;
  const int64_t chunk_stride = (int64_t)L * kMaxB * H;
  const float alpha = a.st->alpha;
  auto adam_at = [&](int64_t idx, float g) {
    float w = a.small[idx], m = a.m_small[idx], v = a.v_small[idx];
    adam_update(w, m, v, g, alpha);
    a.small[idx] = w;
    a.m_small[idx] = m;
    a.v_small[idx] = v;
  };
  if ((int)blockIdx.x < nblk_hidden) {
    const int i = 1 + blockIdx.x / rb_n, rb = blockIdx.x % rb_n;
    float g16[kBbRows];
#pragma unroll
    for (int r = 0; r < kBbRows; ++r) g16[r] = 0.f;
    float bsum = 0.f;
    for (int ci = 0; ci < nchunks; ++ci) {
      const float* dzs = a.dzs + ci * chunk_stride;
      const float* acts = a.acts + ci * chunk_stride;
      float dz[kMaxB];
#pragma unroll
      for (int b = 0; b < kMaxB; ++b) dz[b] = dzs[((int64_t)i * kMaxB + b) * H + j];
      __syncthreads();  // the previous chunk's activations have been consumed
      for (int idx = threadIdx.x; idx < kBbRows * kMaxB; idx += blockDim.x) {
        const int kk = idx % kBbRows, b = idx / kBbRows;
        as[kk][b] = acts[((int64_t)(i - 1) * kMaxB + b) * H + rb * kBbRows + kk];
      }
      __syncthreads();
#pragma unroll
      for (int b = 0; b < kMaxB; ++b) bsum += dz[b];
#pragma unroll
      for (int r = 0; r < kBbRows; ++r) {
        float g = 0.f;
#pragma unroll
        for (int b = 0; b < kMaxB; ++b) g = fmaf(as[r][b], dz[b], g);
        g16[r] += g;
      }
    }
#pragma unroll
    for (int r = 0; r < kBbRows; ++r) {
      const int k = rb * kBbRows + r;
      const int64_t idx = sl.Wh(i) + (int64_t)k * H + j;
      float w = a.small[idx], m = a.m_small[idx], v = a.v_small[idx];
      adam_update(w, m, v, g16[r], alpha);
      a.small[idx] = w;
      a.m_small[idx] = m;
      a.v_small[idx] = v;
      if (a.slice_mode == 1)
        store_images(a.w_fs, a.w_bs, i, k, j, w);
      else
        store_sliced(a.w_fs, a.w_bs, H, a.Hc, i, k, j, w);
    }
    if (rb == 0) adam_at(sl.bh(i) + j, bsum);
  } else {
    float gb1 = 0.f, g0 = 0.f, g1 = 0.f, sb = 0.f;
    for (int ci = 0; ci < nchunks; ++ci) {
      const float* dzs = a.dzs + ci * chunk_stride;
      const float* acts = a.acts + ci * chunk_stride;
      const float* y1 = a.outs + ci * 256;
      const float* dy1 = y1 + 64;
      const float* dy2 = y1 + 128;
      for (int b = 0; b < kMaxB; ++b) {
        gb1 += dzs[(int64_t)b * H + j];  // b1: column sum of dZ1
        const float av = acts[((int64_t)(L - 1) * kMaxB + b) * H + j];  // Dense(2): Wo1[k][c] = sum_b a_{L-1}[b][k] dy1[b][c]
        g0 = fmaf(av, dy1[2 * b], g0);
        g1 = fmaf(av, dy1[2 * b + 1], g1);
      }
      if (j < 2) {  // bo1
        for (int b = 0; b < kMaxB; ++b) sb += dy1[2 * b + j];
      } else if (j < 6) {  // Wo2[i][c] = sum_b y1[b][i] dy2[b][c]
        const int q = j - 2, i2 = q / 2, c = q % 2;
        for (int b = 0; b < kMaxB; ++b) sb = fmaf(y1[2 * b + i2], dy2[2 * b + c], sb);
      } else if (j < 8) {  // bo2
        for (int b = 0; b < kMaxB; ++b) sb += dy2[2 * b + (j - 6)];
      }
    }
    adam_at(sl.b1() + j, gb1);
    adam_at(sl.Wo1() + 2 * j, g0);
    adam_at(sl.Wo1() + 2 * j + 1, g1);
    if (j < 2)
      adam_at(sl.bo1() + j, sb);
    else if (j < 6)
      adam_at(sl.Wo2() + (j - 2), sb);
    else if (j < 8)
      adam_at(sl.bo2() + (j - 6), sb);
  }
}

static int bb_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int bb_stats_launch(const BigArgs& a, cudaStream_t s) {
  LOC_CHECK(a.nb >= 1 && a.nb <= LOC_MAX_BATCH_SIZE, "batch statistics: a step holds 1..256 rows");
  const int64_t nwords = cdiv(a.K, 16);
  k_bb_stats<<<(unsigned)cdiv(nwords, 128), 128, 0, s>>>(a);
  LOC_LAUNCHED();
  return 0;
}

int bb_l1_backward_launch(const BigArgs& a, cudaStream_t s) {
  LOC_CHECK(a.H % 32 == 0 && a.H >= 32 && a.H <= 1024, "first layer (large batch): width must be a multiple of 32 in [32, 1024]");
  const int threads = a.H < 64 ? 64 : a.H;
  const int Bp = (a.nb + 3) & ~3;
  const size_t smem = ((size_t)kBbT * Bp + (size_t)(threads / 32) * kBbT * 2) * sizeof(float);
  const int64_t nwords = cdiv(a.K, kBbT);
  const int per_sm = threads <= 256 ? 4 : (threads <= 512 ? 2 : 1);
  int64_t blocks = (int64_t)bb_sms() * per_sm;
  if (blocks > nwords) blocks = nwords;
  k_bb_l1_bwd<<<(unsigned)blocks, threads, smem, s>>>(a);
  LOC_LAUNCHED();
  return 0;
}

int bb_hidden_update_launch(const BigArgs& a, cudaStream_t s) {
  LOC_CHECK(a.H % kBbRows == 0 && a.H >= 32 && a.H <= 1024, "hidden update (large batch): bad width");
  const int nblk = (a.L - 1) * (a.H / kBbRows) + 1;
  k_bb_hidden_update<<<nblk, a.H, 0, s>>>(a);
  LOC_LAUNCHED();
  return 0;
}

}  // namespace loc
