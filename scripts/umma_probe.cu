// Bring-up probe (GPU box): single-CTA tcgen05.mma kind::tf32 against a CPU product, for the
// operand layouts the first-layer kernels use.  nvcc -gencode arch=compute_100a,code=sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

enum Mode { KM_NONE = 0, KM_SW128 = 1, MN_SW128 = 2, MN_SW128_32B = 3 };

struct Cfg {
  int a_mode, b_mode, N, Ktot;
  uint32_t idesc;
  int a_lbo, a_sbo, b_lbo, b_sbo;  // bytes
  int a_kadv, b_kadv;              // bytes per MMA along K
  int layout_a, layout_b;          // descriptor layout type
  int version;
  int M;
};

__host__ __device__ inline uint32_t place(int mode, int mn, int k, int MN, int Ktot) {
  if (mode == KM_NONE) return (uint32_t)((mn / 8) * (128 * (Ktot / 4)) + (k / 4) * 128 + (mn % 8) * 16 + (k % 4) * 4);
  if (mode == KM_SW128) return (uint32_t)(mn * 128 + ((((k / 4) ^ (mn & 7))) * 16) + (k % 4) * 4);
  if (mode == MN_SW128_32B)  // [mn/32 chunks][k rows][128 B], 32-byte sub-chunks XOR (row & 3)
    return (uint32_t)((mn / 32) * (Ktot * 128) + k * 128 + ((((mn % 32) / 8) ^ (k & 3)) * 32) + (mn % 8) * 4);
  // MN_SW128: [mn/32 chunks][k rows][128 B]
  return (uint32_t)((mn / 32) * (Ktot * 128) + k * 128 + ((((mn % 32) / 4) ^ (k & 7)) * 16) + (mn % 4) * 4);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo, int sbo, int layout, int version) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)(((uint32_t)lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(((uint32_t)sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)version << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(Cfg c, const float* A, const float* B, float* D, float* D2, int do_st) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sm;
  uint8_t* sB = sm + 65536;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < c.M * c.Ktot; i += 128) {
    const int m = i / c.Ktot, k = i % c.Ktot;
    *(float*)(sA + place(c.a_mode, m, k, c.M, c.Ktot)) = A[i];
  }
  for (int i = tid; i < c.N * c.Ktot; i += 128) {
    const int n = i / c.Ktot, k = i % c.Ktot;
    *(float*)(sB + place(c.b_mode, n, k, c.N, c.Ktot)) = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) D2[0] = __uint_as_float(tmem);
  // ---- T0: tcgen05.st / ld round trip at columns 128.. ----
  if (do_st) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint((float)(tid * 8 + i));
    const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16) + 128u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(ta)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) D2[1 + tid * 8 + i] = __uint_as_float(r[i]);
  }
  if (tid == 0) {
    const int nk = c.Ktot / 8;
    for (int ks = 0; ks < nk; ++ks) {
      const uint64_t ad = make_desc(smem_u32(sA) + ks * c.a_kadv, c.a_lbo, c.a_sbo, c.layout_a, c.version);
      const uint64_t bd = make_desc(smem_u32(sB) + ks * c.b_kadv, c.b_lbo, c.b_sbo, c.layout_b, c.version);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(ad), "l"(bd), "r"(c.idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
  }
  // wait
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(ta)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) D[(32 * warp + lane) * c.N + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static uint32_t idesc(int M, int N, int amn, int bmn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

static void run(const char* name, Cfg c, int do_st) {
  if (c.M == 0) c.M = 128;
  const int M = c.M;
  std::vector<float> A(M * c.Ktot), B(c.N * c.Ktot), D(128 * c.N, -1.f), D2(1 + 128 * 8, -1.f);
  for (int i = 0; i < M * c.Ktot; ++i) A[i] = (float)((i * 7 + 3) % 11 - 5);
  for (int i = 0; i < c.N * c.Ktot; ++i) B[i] = (float)((i * 5 + 1) % 7 - 3);
  float *dA, *dB, *dD, *dD2;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMalloc(&dD2, D2.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD2, D2.data(), D2.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  probe<<<1, 128, 140 * 1024>>>(c, dA, dB, dD, dD2, do_st);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-28s CUDA error: %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  int nz = 0;
  if (M == 64) {  // which TMEM lanes hold data?
    printf("  lanes with data:");
    for (int l = 0; l < 128; ++l) {
      bool any = false;
      for (int n = 0; n < c.N; ++n) any |= (D[l * c.N + n] != 0.f && D[l * c.N + n] != -1.f);
      if (any) printf(" %d", l);
    }
    printf("\n");
  }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < c.N; ++n) {
      double r = 0;
      for (int k = 0; k < c.Ktot; ++k) r += (double)A[m * c.Ktot + k] * B[n * c.Ktot + k];
      const int lane = (M == 64) ? 32 * (m / 16) + (m % 16) : m;  // hypothesis for M = 64
      double d = fabs(r - D[lane * c.N + n]);
      if (d > maxerr) maxerr = d;
      if (fabs(r) > maxref) maxref = fabs(r);
      if (D[lane * c.N + n] != 0.f) ++nz;
    }
  printf("%-28s maxerr %.3g (max ref %.3g) nonzero %d/%d  D[0][0..3]= %g %g %g %g  D[1][0]=%g\n", name, maxerr, maxref, nz,
         M * c.N, D[0], D[1], D[2], D[3], D[c.N]);
  if (do_st) {
    int bad = 0;
    for (int i = 0; i < 128 * 8; ++i) bad += (D2[1 + i] != (float)i);
    uint32_t t;
    memcpy(&t, &D2[0], 4);
    printf("  tmem base 0x%08x; st/ld round trip mismatches: %d\n", t, bad);
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  cudaFree(dD2);
}

int main() {
  for (int version = 1; version >= 0; --version) {
    printf("== descriptor version %d ==\n", version);
    if (version == 1) {  // hidden-stack configs: M = 64, N = 8
      Cfg c = {MN_SW128_32B, KM_SW128, 8, 32, idesc(64, 8, 1, 0), 32 * 128, 512, 16, 1024, 1024, 32, 1, 2, version, 64};
      run("M64 MN_32B x KM_SW128 N8", c, 0);
      Cfg c2 = {KM_SW128, KM_SW128, 8, 32, idesc(64, 8, 0, 0), 16, 1024, 16, 1024, 32, 32, 2, 2, version, 64};
      run("M64 KM_SW128 x KM_SW128 N8", c2, 0);
    }
    {  // K-major / K-major, no swizzle
      Cfg c = {KM_NONE, KM_NONE, 32, 32, idesc(128, 32, 0, 0), 128, 128 * 8, 128, 128 * 8, 256, 256, 0, 0, version};
      run("KM_NONE x KM_NONE N32", c, version == 1);
    }
    {  // K-major / K-major, SW128
      Cfg c = {KM_SW128, KM_SW128, 32, 32, idesc(128, 32, 0, 0), 16, 1024, 16, 1024, 32, 32, 2, 2, version};
      run("KM_SW128 x KM_SW128 N32", c, 0);
    }
    {  // forward config: A MN-major SW128, B MN-major SW128
      Cfg c = {MN_SW128, MN_SW128, 32, 32, idesc(128, 32, 1, 1), 32 * 128, 1024, 32 * 128, 1024, 1024, 1024, 2, 2, version};
      run("MN_SW128 x MN_SW128 N32", c, 0);
    }
    {  // forward config with the 32B-atom swizzle (tf32 MN-major): LBO = chunk stride, SBO = 4-row group stride
      Cfg c = {MN_SW128_32B, MN_SW128_32B, 32, 32, idesc(128, 32, 1, 1), 32 * 128, 512, 32 * 128, 512, 1024, 1024, 1, 1, version};
      run("MN_32B x MN_32B N32", c, 0);
    }
    {  // backward config: A MN-major 32B-atom, B K-major SW128, N = 64
      Cfg c = {MN_SW128_32B, KM_SW128, 64, 32, idesc(128, 64, 1, 0), 32 * 128, 512, 16, 1024, 1024, 32, 1, 2, version};
      run("MN_32B x KM_SW128 N64", c, 0);
    }
    {  // swapped LBO/SBO roles, in case the canonical reading is the other way round
      Cfg c = {MN_SW128_32B, KM_SW128, 64, 32, idesc(128, 64, 1, 0), 512, 32 * 128, 16, 1024, 1024, 32, 1, 2, version};
      run("MN_32B(swapped) x KM N64", c, 0);
    }
    {  // backward config: A MN-major SW128, B K-major SW128, N = 64
      Cfg c = {MN_SW128, KM_SW128, 64, 32, idesc(128, 64, 1, 0), 32 * 128, 1024, 16, 1024, 1024, 32, 2, 2, version};
      run("MN_SW128 x KM_SW128 N64", c, 0);
    }
  }
  return 0;
}
