// Per-SNP batch statistics shared by the first-layer kernels (SIMT and tcgen05).
#pragma once
#include "model.cuh"

namespace loc {

// Batch statistics the way tf.nn.moments computes them (mean, then mean of squared differences),
// from the genotype counts of the nb rows.  True divisions: a column that is constant within the
// batch gives mean == x exactly, hence an exactly-zero centred column and dgamma == 0 as in Keras.
__device__ __forceinline__ void moments_from_counts(int n1, int n2, int nb, float& mean, float& var) {
  const float fn = (float)nb;
  const int n0 = nb - n1 - n2;
  mean = (float)(n1 + 2 * n2) / fn;
  const float d0 = 0.f - mean, d1 = 1.f - mean, d2 = 2.f - mean;
  var = ((float)n0 * d0 * d0 + (float)n1 * d1 * d1 + (float)n2 * d2 * d2) / fn;
}

}  // namespace loc
