// Peer-memory exchange for a model sharded over SNP columns (tensor parallelism).
//
// The one exchange of the path is the sum over shards of the [32][256] first-layer tile.  Instead of a
// library all-reduce, every shard reduces its own split-K partial tiles and PUSHES the result into slot
// [rank] of every peer's buffer over NVLink (plain st.global to cudaIpc-mapped peer memory), then raises
// one flag per (sender, block) in the peer's memory with a system-scope release.  The receiver's hidden-stack kernel polls its
// LOCAL flags (acquire.sys) in its prologue and sums the `world` slots in rank order -- the same
// order on every shard, so the replicated hidden stacks stay bitwise identical.  Buffers are double
// buffered by exchange parity: a shard can be at most one exchange ahead of a peer (it needs the peer's
// tile of exchange s before it can produce exchange s + 1).
#include "model.cuh"

namespace loc {

constexpr int kTpBlocks = 64;   // blocks of the reduce-and-push kernel = flags per sender
constexpr int kTpThreads = 128;
constexpr int kTpMaxWorld = 8;

struct TpPeers {
  float* slots[kTpMaxWorld];      // peer p's buffer: [2 parities][world][tile]
  uint32_t* flags[kTpMaxWorld];   // peer p's flags:  [2 parities][world][kTpBlocks]
};

}  // namespace loc

struct loc_tp {
  int rank, world, tile;          // tile = kMaxB * width floats
  uint8_t* local;                 // cudaMalloc'd: slots then flags then error word
  size_t slot_bytes, flag_bytes;
  void* peer_base[loc::kTpMaxWorld];
  loc::TpPeers peers;
  uint32_t seq;                   // exchanges issued so far (identical on every shard)
  int connected;
};

namespace loc {

__global__ void __launch_bounds__(512) k_reduce_push(const float* __restrict__ partials, int n_partials, int n, TpPeers peers,
                                                     int rank, int world, uint32_t seq) {
  __shared__ float sred[4][128];
  const float v = reduce_partial_tiles(partials, n_partials, n, sred);
  const int i = blockIdx.x * kTpThreads + threadIdx.x;
  if (threadIdx.x < kTpThreads && i < n) {
    const int64_t off = ((int64_t)(seq & 1u) * world + rank) * n + i;
    for (int d = 0; d < world; ++d) peers.slots[d][off] = v;
  }
  __syncthreads();
  if (threadIdx.x < world) {  // one thread per destination: fence (cumulative over the block's stores), then flag
    __threadfence_system();
    uint32_t* f = peers.flags[threadIdx.x] + ((int64_t)(seq & 1u) * world + rank) * kTpBlocks + blockIdx.x;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
  }
}

int tp_exchange(loc_tp* tp, const float* partials, int n_partials, cudaStream_t s) {
  LOC_CHECK(tp != nullptr && tp->connected, "sharded model: peer buffers are not connected (loc_tp_connect)");
  const uint32_t seq = ++tp->seq;
  const int n = tp->tile;
  LOC_CHECK(cdiv(n, kTpThreads) == kTpBlocks, "sharded model: unexpected tile size");
  k_reduce_push<<<kTpBlocks, 512, 0, s>>>(partials, n_partials, n, tp->peers, tp->rank, tp->world, seq);
  LOC_LAUNCHED();
  return 0;  // the consumer (hidden-stack kernel) polls the local flags in its prologue: tp_wait_info
}

void tp_wait_info(const loc_tp* tp, const uint32_t** flags, int* count, uint32_t* seq, int** err) {
  *flags = reinterpret_cast<const uint32_t*>(tp->local + tp->slot_bytes) + (size_t)(tp->seq & 1u) * tp->world * kTpBlocks;
  *count = tp->world * kTpBlocks;
  *seq = tp->seq;
  *err = reinterpret_cast<int*>(tp->local + tp->slot_bytes + tp->flag_bytes);
}

int tp_world(const loc_tp* tp) { return tp->world; }

const float* tp_tiles(const loc_tp* tp) {
  return reinterpret_cast<const float*>(tp->local) + (size_t)(tp->seq & 1u) * tp->world * tp->tile;
}

}  // namespace loc

extern "C" {

int loc_tp_create(loc_tp** out, int32_t rank, int32_t world, int32_t width) {
  LOC_CHECK(out != nullptr && world >= 1 && world <= loc::kTpMaxWorld && rank >= 0 && rank < world && width == 256,
            "loc_tp_create: bad arguments (1 <= world <= 8, width 256)");
  loc_tp* tp = new (std::nothrow) loc_tp();
  LOC_CHECK(tp != nullptr, "loc_tp_create: out of host memory");
  memset(tp, 0, sizeof(*tp));
  tp->rank = rank;
  tp->world = world;
  tp->tile = loc::kMaxB * width;
  tp->slot_bytes = (size_t)2 * world * tp->tile * sizeof(float);
  tp->flag_bytes = (size_t)2 * world * loc::kTpBlocks * sizeof(uint32_t);
  const size_t total = tp->slot_bytes + tp->flag_bytes + 256;
  if (cudaMalloc(&tp->local, total) != cudaSuccess || cudaMemset(tp->local, 0, total) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    delete tp;
    return loc::fail("loc_tp_create: device allocation failed", __FILE__, __LINE__);
  }
  *out = tp;
  return 0;
}

int loc_tp_handle(loc_tp* tp, uint8_t* h_handle) {
  LOC_CHECK(tp != nullptr && h_handle != nullptr, "loc_tp_handle: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  LOC_CUDA(cudaIpcGetMemHandle(&h, tp->local));
  memcpy(h_handle, &h, sizeof(h));
  return 0;
}

int loc_tp_connect(loc_tp* tp, const uint8_t* h_handles) {
  LOC_CHECK(tp != nullptr && h_handles != nullptr && !tp->connected, "loc_tp_connect: bad arguments");
  for (int p = 0; p < tp->world; ++p) {
    void* base = tp->local;
    if (p != tp->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, h_handles + (size_t)p * sizeof(h), sizeof(h));
      LOC_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
      tp->peer_base[p] = base;
    }
    tp->peers.slots[p] = reinterpret_cast<float*>(base);
    tp->peers.flags[p] = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(base) + tp->slot_bytes);
  }
  tp->connected = 1;
  return 0;
}

int loc_tp_error(loc_tp* tp) {
  if (tp == nullptr) return 1;
  int e = 0;
  if (cudaMemcpy(&e, tp->local + tp->slot_bytes + tp->flag_bytes, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return 1;
  return e;
}

int loc_tp_destroy(loc_tp* tp) {
  if (tp == nullptr) return 0;
  cudaDeviceSynchronize();
  for (int p = 0; p < tp->world; ++p)
    if (tp->peer_base[p] != nullptr) cudaIpcCloseMemHandle(tp->peer_base[p]);
  if (tp->local != nullptr) cudaFree(tp->local);
  delete tp;
  return 0;
}

}  // extern "C"
