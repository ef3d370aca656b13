// The width-H stack behind the first layer: Dense(H,elu) x (L-1) (+Dropout), Dense(2), Dense(2),
// Euclidean-distance loss, and the backward chain down to dZ1 -- one thread-block cluster per step.
//
// Reference: load_network locator/locator.py:311-327 (layers, loss :314-315), model.fit :367-376;
// Keras semantics restated in oracle/model_ref.py (RefLocator.forward / gradients).
//
// Work split: CTA r of the C-CTA cluster owns columns [r*Hc, (r+1)*Hc) of every layer's output.
// A layer is: all-gather the previous activations (through L2; barrier.cluster release/acquire),
// each CTA computes its [32 x Hc] slice with the 8 warps splitting the reduction dimension,
// fixed-order cross-warp sum (deterministic), elu / dropout, publish the slice.  The backward
// pass mirrors it with row slices of W.  dW + Adam of these layers is NOT done here: the
// activations and dz of every layer are left in L2 for k_hidden_update, which runs off the
// critical path next to the first-layer backward.
#include "model.cuh"
#include "philox.cuh"

namespace loc {

__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_size() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// Release/acquire at cluster scope: global writes made before it by any CTA of the cluster are
// visible to every CTA after it.
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int kHidThreads = 256;
constexpr int kHidWarps = kHidThreads / 32;
constexpr int kPad = 4;

// [32][H] tile from L2 into padded shared memory.
__device__ __forceinline__ void load_full(float* full, const float* __restrict__ src, int H) {
  const int HP = H + kPad;
  const int nvec = kMaxB * H / 4;
  for (int i = threadIdx.x; i < nvec; i += kHidThreads) {
    const int b = (i * 4) / H, k = (i * 4) % H;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + i);
    *reinterpret_cast<float4*>(full + b * HP + k) = v;
  }
}

// red[w][b][jl] = sum over warp w's share of the reduction dimension.
//   fwd (transpose == 0): out[b][jl] = sum_k full[b][k] * W[k*H + j0 + jl]
//   bwd (transpose == 1): out[b][il] = sum_j full[b][j] * W[(j0 + il)*H + j]
template <int TRANSPOSE>
__device__ __forceinline__ void slice_matmul(const float* full, const float* __restrict__ W, float* red, int H, int Hc,
                                             int j0) {
  const int HP = H + kPad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kper = H / kHidWarps;
  const int kbeg = warp * kper, kend = kbeg + kper;
  const int jq_n = Hc / 4;
  const int ntile = (kMaxB / 4) * jq_n;
  for (int tile = lane; tile < ntile; tile += 32) {
    const int bq = tile / jq_n, jq = tile % jq_n;
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
    for (int k = kbeg; k < kend; k += 4) {
      float4 a4[4], w4[4];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) a4[bb] = *reinterpret_cast<const float4*>(full + (4 * bq + bb) * HP + k);
      if (TRANSPOSE == 0) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          w4[kk] = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(k + kk) * H + j0 + 4 * jq));
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const float av[4] = {a4[bb].x, a4[bb].y, a4[bb].z, a4[bb].w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            acc[bb][0] = fmaf(av[kk], w4[kk].x, acc[bb][0]);
            acc[bb][1] = fmaf(av[kk], w4[kk].y, acc[bb][1]);
            acc[bb][2] = fmaf(av[kk], w4[kk].z, acc[bb][2]);
            acc[bb][3] = fmaf(av[kk], w4[kk].w, acc[bb][3]);
          }
        }
      } else {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
          w4[ii] = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(j0 + 4 * jq + ii) * H + k));
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            float s = acc[bb][ii];
            s = fmaf(a4[bb].x, w4[ii].x, s);
            s = fmaf(a4[bb].y, w4[ii].y, s);
            s = fmaf(a4[bb].z, w4[ii].z, s);
            s = fmaf(a4[bb].w, w4[ii].w, s);
            acc[bb][ii] = s;
          }
      }
    }
#pragma unroll
    for (int bb = 0; bb < 4; ++bb)
      *reinterpret_cast<float4*>(red + ((int64_t)warp * kMaxB + 4 * bq + bb) * Hc + 4 * jq) =
          make_float4(acc[bb][0], acc[bb][1], acc[bb][2], acc[bb][3]);
  }
}

__global__ void __launch_bounds__(kHidThreads) k_hidden(HidArgs a) {
  if (a.gated && a.st->stopped) return;
  extern __shared__ __align__(16) float smem[];
  __shared__ int64_t s_rows[kMaxB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = a.H, L = a.L;
  const int C = (int)cluster_size(), r = (int)cluster_rank();
  const int Hc = H / C, j0 = r * Hc;
  const int HP = H + kPad;
  const int nb = a.src.nb;
  const SmallLayout sl{H, L};

  float* full = smem;                               // [32][HP]
  float* red = full + kMaxB * HP;                   // [8][32][Hc] (>= 1024 floats)
  float* own_a = red + kHidWarps * kMaxB * Hc;      // [L][32][Hc] elu outputs (pre-dropout) of the own slice
  float* keep = own_a + (int64_t)L * kMaxB * Hc;    // [32][Hc]   dropout multiplier of the own slice
  float* ysm = keep + kMaxB * Hc;                   // y1, y2, dy1, dy2 [32][2] each, dist[32]
  float* y1s = ysm;
  float* y2s = ysm + 64;
  float* dy1s = ysm + 128;
  float* dy2s = ysm + 192;
  float* dist = ysm + 256;

  if (tid < kMaxB) s_rows[tid] = tid < nb ? row_of(a.src, a.st, tid) : 0;
  const int step_id = a.st->step_id;
  const float keep_scale = 1.0f / (1.0f - a.p_drop);
  const bool drop_on = a.training && a.p_drop > 0.f;

  // elu + (dropout) + publish of the own slice of layer i, from the cross-warp partial sums in red.
  auto finish_fwd = [&](int i, int nparts, const float* bias) {
    for (int idx = tid; idx < kMaxB * Hc; idx += kHidThreads) {
      const int b = idx / Hc, jl = idx % Hc;
      float z = 0.f;
      for (int w = 0; w < nparts; ++w) z += red[((int64_t)w * kMaxB + b) * Hc + jl];
      z += bias[j0 + jl];
      float act = elu_f(z);
      own_a[((int64_t)i * kMaxB + b) * Hc + jl] = act;
      if (i == a.n_before - 1) {
        float mult = 1.f;
        if (drop_on) {
          bool kp;
          if (a.masks != nullptr) {
            const int64_t s = step_id < a.n_masks ? step_id : a.n_masks - 1;
            kp = a.masks[(s * kMaxB + b) * H + j0 + jl] != 0;
          } else {
            kp = philox_uniform((uint64_t)b * H + j0 + jl, kDropoutStreamBase + (uint32_t)step_id, a.seed) >= a.p_drop;
          }
          mult = kp ? keep_scale : 0.f;
        }
        keep[b * Hc + jl] = mult;
        act *= mult;
      }
      a.acts[((int64_t)i * kMaxB + b) * H + j0 + jl] = b < nb ? act : 0.f;
    }
  };
  // dz of layer i for the own slice from d loss / d (post-dropout activation).
  auto finish_bwd = [&](int i, int b, int jl, float da) {
    if (i == a.n_before - 1) da *= keep[b * Hc + jl];
    const float act = own_a[((int64_t)i * kMaxB + b) * Hc + jl];
    const float dz = da * elu_grad_from_out(act);
    a.dzs[((int64_t)i * kMaxB + b) * H + j0 + jl] = b < nb ? dz : 0.f;
  };

  // ---- layer 0: reduce the split-K partial tiles of Z1 (fixed order), bias, elu ----
  {
    const int nvec = kMaxB * Hc / 4;                       // float4 outputs of the slice
    const int G = nvec >= kHidThreads ? 1 : kHidThreads / nvec;  // groups splitting the partial range
    const int hv = H / 4;
    for (int idx = tid; idx < nvec * G; idx += kHidThreads) {
      const int o = idx % nvec, g = idx / nvec;
      const int b = (o * 4) / Hc, jl = (o * 4) % Hc;
      const int pbeg = a.n_partials * g / G, pend = a.n_partials * (g + 1) / G;
      const float4* src = reinterpret_cast<const float4*>(a.partials) + ((int64_t)b * H + j0 + jl) / 4;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int p = pbeg; p < pend; ++p) {
        const float4 v = __ldcg(src + (int64_t)p * kMaxB * hv);
        s.x += v.x;
        s.y += v.y;
        s.z += v.z;
        s.w += v.w;
      }
      *reinterpret_cast<float4*>(red + ((int64_t)g * kMaxB + b) * Hc + jl) = s;
    }
    __syncthreads();
    finish_fwd(0, G, a.small + sl.b1());
  }
  cluster_barrier();

  // ---- layers 1..L-1 forward ----
  for (int i = 1; i < L; ++i) {
    load_full(full, a.acts + (int64_t)(i - 1) * kMaxB * H, H);
    __syncthreads();
    slice_matmul<0>(full, a.small + sl.Wh(i), red, H, Hc, j0);
    __syncthreads();
    finish_fwd(i, kHidWarps, a.small + sl.bh(i));
    cluster_barrier();
  }

  // ---- Dense(2), Dense(2), loss (every CTA, redundantly) ----
  load_full(full, a.acts + (int64_t)(L - 1) * kMaxB * H, H);
  __syncthreads();
  {
    const float* Wo1 = a.small + sl.Wo1();
    for (int b = warp; b < kMaxB; b += kHidWarps) {
      float s0 = 0.f, s1 = 0.f;
      for (int k = lane; k < H; k += 32) {
        const float av = full[b * HP + k];
        s0 = fmaf(av, __ldg(Wo1 + 2 * k), s0);
        s1 = fmaf(av, __ldg(Wo1 + 2 * k + 1), s1);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (lane == 0) {
        y1s[2 * b] = s0;
        y1s[2 * b + 1] = s1;
      }
    }
  }
  __syncthreads();
  if (tid < kMaxB) {
    const int b = tid;
    const float* bo1 = a.small + sl.bo1();
    const float* Wo2 = a.small + sl.Wo2();
    const float* bo2 = a.small + sl.bo2();
    const float u0 = y1s[2 * b] + bo1[0], u1 = y1s[2 * b + 1] + bo1[1];
    y1s[2 * b] = u0;
    y1s[2 * b + 1] = u1;
    const float v0 = u0 * Wo2[0] + u1 * Wo2[2] + bo2[0];
    const float v1 = u0 * Wo2[1] + u1 * Wo2[3] + bo2[1];
    y2s[2 * b] = v0;
    y2s[2 * b + 1] = v1;
    float d = 0.f, g0 = 0.f, g1 = 0.f;
    if (b < nb && (a.training || a.has_targets)) {
      const float t0 = a.locs[2 * s_rows[b]], t1 = a.locs[2 * s_rows[b] + 1];
      const float e0 = v0 - t0, e1 = v1 - t1;
      d = sqrtf(e0 * e0 + e1 * e1);
      const float den = d * (float)nb;  // no epsilon: NaN when the prediction hits the target, as in the reference
      g0 = e0 / den;
      g1 = e1 / den;
    }
    dist[b] = d;
    dy2s[2 * b] = g0;
    dy2s[2 * b + 1] = g1;
    dy1s[2 * b] = g0 * Wo2[0] + g1 * Wo2[1];
    dy1s[2 * b + 1] = g0 * Wo2[2] + g1 * Wo2[3];
    if (r == 0 && a.write_pred && b < nb) {
      a.pred_out[2 * s_rows[b]] = v0;
      a.pred_out[2 * s_rows[b] + 1] = v1;
    }
  }
  __syncthreads();
  if (r == 0) {
    if (tid < 64) {
      a.outs[tid] = y1s[tid];
      a.outs[64 + tid] = dy1s[tid];
      a.outs[128 + tid] = dy2s[tid];
      a.outs[192 + tid] = y2s[tid];
    }
    if (tid == 0 && (a.training || a.has_targets)) {
      float s = 0.f;
      for (int b = 0; b < nb; ++b) s += dist[b];
      const float mean = s / (float)nb;
      DevState* st = a.st;
      if (a.training) {
        st->loss_total += mean * (float)nb;
        st->loss_count += (float)nb;
        st->last_loss = mean;
        if (!isfinite(mean)) st->nonfinite = 1;
      } else {
        st->val_total += mean * (float)nb;
        st->val_count += (float)nb;
      }
    }
  }
  if (!a.training) return;

  // ---- backward: d loss / d a_{L-1} through Dense(2) ----
  {
    const float* Wo1 = a.small + sl.Wo1();
    for (int idx = tid; idx < kMaxB * Hc; idx += kHidThreads) {
      const int b = idx / Hc, jl = idx % Hc;
      const int i = j0 + jl;
      const float da = dy1s[2 * b] * Wo1[2 * i] + dy1s[2 * b + 1] * Wo1[2 * i + 1];
      finish_bwd(L - 1, b, jl, da);
    }
  }
  cluster_barrier();
  for (int i = L - 1; i >= 1; --i) {
    load_full(full, a.dzs + (int64_t)i * kMaxB * H, H);
    __syncthreads();
    slice_matmul<1>(full, a.small + sl.Wh(i), red, H, Hc, j0);
    __syncthreads();
    for (int idx = tid; idx < kMaxB * Hc; idx += kHidThreads) {
      const int b = idx / Hc, jl = idx % Hc;
      float da = 0.f;
      for (int w = 0; w < kHidWarps; ++w) da += red[((int64_t)w * kMaxB + b) * Hc + jl];
      finish_bwd(i - 1, b, jl, da);
    }
    cluster_barrier();
  }
  // Optimizer bookkeeping for the update kernels of this step (Keras Adam: alpha from t >= 1).
  if (r == 0 && tid == 0) {
    DevState* st = a.st;
    const int t = st->t + 1;
    st->t = t;
    st->step_id = step_id + 1;
    const float b1p = powf(kAdamB1, (float)t), b2p = powf(kAdamB2, (float)t);
    st->alpha = st->lr * sqrtf(1.f - b2p) / (1.f - b1p);
  }
}

// ---------------------------------------------------------------------------------------------
// dW + Adam of the small layers.  Blocks [0, (L-1)*H/16): 16 input rows x H outputs of hidden
// layer i; last block: b1, Dense(2), Dense(2).  blockDim = H, thread <-> output column j.
// ---------------------------------------------------------------------------------------------
constexpr int kUpdRows = 16;

__global__ void __launch_bounds__(1024) k_hidden_update(UpdArgs a) {
  if (a.gated && a.st->stopped) return;
  __shared__ float as[kUpdRows][kMaxB + 1];
  const int H = a.H, L = a.L, j = threadIdx.x;
  const SmallLayout sl{H, L};
  const float alpha = a.st->alpha;
  const int rb_n = H / kUpdRows;
  const int nblk_hidden = (L - 1) * rb_n;
  auto adam_at = [&](int64_t idx, float g) {
    float w = a.small[idx], m = a.m_small[idx], v = a.v_small[idx];
    adam_update(w, m, v, g, alpha);
    a.small[idx] = w;
    a.m_small[idx] = m;
    a.v_small[idx] = v;
  };
  if ((int)blockIdx.x < nblk_hidden) {
    const int i = 1 + blockIdx.x / rb_n, rb = blockIdx.x % rb_n;
    float dz[kMaxB];
    float bsum = 0.f;
#pragma unroll
    for (int b = 0; b < kMaxB; ++b) {
      dz[b] = a.dzs[((int64_t)i * kMaxB + b) * H + j];
      bsum += dz[b];
    }
    for (int idx = threadIdx.x; idx < kUpdRows * kMaxB; idx += blockDim.x) {
      const int kk = idx % kUpdRows, b = idx / kUpdRows;
      as[kk][b] = a.acts[((int64_t)(i - 1) * kMaxB + b) * H + rb * kUpdRows + kk];
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < kUpdRows; ++kk) {
      float g = 0.f;
#pragma unroll
      for (int b = 0; b < kMaxB; ++b) g = fmaf(as[kk][b], dz[b], g);
      adam_at(sl.Wh(i) + (int64_t)(rb * kUpdRows + kk) * H + j, g);
    }
    if (rb == 0) adam_at(sl.bh(i) + j, bsum);
  } else {
    // b1: gradient is the column sum of dZ1
    float g = 0.f;
    for (int b = 0; b < kMaxB; ++b) g += a.dzs[(int64_t)b * H + j];
    adam_at(sl.b1() + j, g);
    // Dense(2): Wo1[k][c] = sum_b a_{L-1}[b][k] dy1[b][c]
    const float* y1 = a.outs;
    const float* dy1 = a.outs + 64;
    const float* dy2 = a.outs + 128;
    float g0 = 0.f, g1 = 0.f;
    for (int b = 0; b < kMaxB; ++b) {
      const float av = a.acts[((int64_t)(L - 1) * kMaxB + b) * H + j];
      g0 = fmaf(av, dy1[2 * b], g0);
      g1 = fmaf(av, dy1[2 * b + 1], g1);
    }
    adam_at(sl.Wo1() + 2 * j, g0);
    adam_at(sl.Wo1() + 2 * j + 1, g1);
    if (j < 2) {  // bo1
      float s = 0.f;
      for (int b = 0; b < kMaxB; ++b) s += dy1[2 * b + j];
      adam_at(sl.bo1() + j, s);
    } else if (j < 6) {  // Wo2[i][c] = sum_b y1[b][i] dy2[b][c]
      const int q = j - 2, i2 = q / 2, c = q % 2;
      float s = 0.f;
      for (int b = 0; b < kMaxB; ++b) s = fmaf(y1[2 * b + i2], dy2[2 * b + c], s);
      adam_at(sl.Wo2() + q, s);
    } else if (j < 8) {  // bo2
      const int c = j - 6;
      float s = 0.f;
      for (int b = 0; b < kMaxB; ++b) s += dy2[2 * b + c];
      adam_at(sl.bo2() + c, s);
    }
  }
}

size_t hidden_smem_bytes(int H, int L, int cluster) {
  const int Hc = H / cluster;
  size_t red = (size_t)kHidWarps * kMaxB * Hc;
  if (red < 1024) red = 1024;
  const size_t floats = (size_t)kMaxB * (H + kPad) + red + (size_t)L * kMaxB * Hc + (size_t)kMaxB * Hc + 320;
  return floats * sizeof(float);
}

static int launch_hidden_cluster(const HidArgs& a, int cluster, cudaStream_t s, bool dry) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cluster);
  cfg.blockDim = dim3(kHidThreads);
  cfg.dynamicSmemBytes = hidden_smem_bytes(a.H, a.L, cluster);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (dry) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_hidden, &cfg);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return n;
  }
  LOC_CUDA(cudaLaunchKernelEx(&cfg, k_hidden, a));
  loc::g_launches.fetch_add(1);
  return 0;
}

int hidden_max_cluster(int H) {
  static int cached_H = -1, cached = 0;
  if (cached_H == H) return cached;
  cudaFuncSetAttribute(k_hidden, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaGetLastError();
  int best = 0;
  const int cands[5] = {16, 8, 4, 2, 1};
  for (int ci = 0; ci < 5 && !best; ++ci) {
    const int c = cands[ci];
    if (H % (4 * c) != 0) continue;
    const size_t smem = hidden_smem_bytes(H, 16, c);
    if (smem > 227 * 1024) continue;
    if (cudaFuncSetAttribute(k_hidden, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    HidArgs dummy = {};
    dummy.H = H;
    dummy.L = 16;
    if (launch_hidden_cluster(dummy, c, 0, true) > 0) best = c;
  }
  cached_H = H;
  cached = best;
  return best;
}

int hidden_launch(const HidArgs& a, int cluster, cudaStream_t s) {
  LOC_CHECK(cluster > 0 && a.H % (4 * cluster) == 0, "hidden stack: width must be a multiple of 4 x cluster size");
  LOC_CHECK(a.H % 32 == 0, "hidden stack: width must be a multiple of 32");
  const size_t smem = hidden_smem_bytes(a.H, a.L, cluster);
  LOC_CHECK(smem <= 227 * 1024, "hidden stack: shared memory budget exceeded for this width / nlayers");
  LOC_CUDA(cudaFuncSetAttribute(k_hidden, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return launch_hidden_cluster(a, cluster, s, false);
}

int hidden_update_launch(const UpdArgs& a, cudaStream_t s) {
  LOC_CHECK(a.H % kUpdRows == 0 && a.H >= 32 && a.H <= 1024, "hidden update: bad width");
  const int nblk = (a.L - 1) * (a.H / kUpdRows) + 1;
  k_hidden_update<<<nblk, a.H, 0, s>>>(a);
  LOC_LAUNCHED();
  return 0;
}

}  // namespace loc
