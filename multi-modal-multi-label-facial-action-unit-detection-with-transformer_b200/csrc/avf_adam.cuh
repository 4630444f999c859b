// The Adam / AdamW update of four consecutive elements (torch.optim.Adam of train.py:334: L2 coupled into the gradient; decoupled =
// AdamW), shared by adam_kernel (avf_train.cu) and the fused all-reduce + Adam kernel (avf_peer.cu).
#pragma once
#include <cuda_bf16.h>

namespace avf {

struct AdamParams {
  float lr, b1, b2, eps, wd, inv_bc1, inv_sqrt_bc2, grad_scale;
  int decoupled;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamParams& a) {
  float gr = g * a.grad_scale;
  if (a.decoupled) p *= 1.f - a.lr * a.wd;
  else gr = fmaf(a.wd, p, gr);
  m = a.b1 * m + (1.f - a.b1) * gr;
  v = a.b2 * v + (1.f - a.b2) * gr * gr;
  const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
  p -= a.lr * a.inv_bc1 * (m / denom);
}

}  // namespace avf
