// tcgen05 first layer -- placeholder until the kernels land.
#include "model.cuh"
namespace loc {
bool l1_tc_supported(int64_t, int) { return false; }
int l1_tc_partials(int64_t) { return 0; }
int l1_forward_tc(const L1Args&, int, cudaStream_t) { return fail("tcgen05 first layer not built", __FILE__, __LINE__); }
int l1_backward_tc(const L1Args&, int, cudaStream_t) { return fail("tcgen05 first layer not built", __FILE__, __LINE__); }
}  // namespace loc
