// Split path of one step (trials sharded over GPUs with an external all-reduce): phase A | local reduce | phase B.
#include "step_kernels.cuh"
#include "kernels.cuh"

__device__ __forceinline__ unsigned base_masks(const StepParams& p) {
  return 1u | ((p.flags & VJF_FLAG_WARMUP) ? 0u : 2u) | 4u;
}

// ---- split path (trials sharded over GPUs): phase A | local reduce | <all-reduce by the caller> | phase B ----
__global__ void __launch_bounds__(VJF_NT, 1) vjf_phase_a_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) float sm[];
  phase_a(p, sm, 0, base_masks(p));
}
__global__ void __launch_bounds__(VJF_NT, 1) vjf_reduce_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) float sm[];
  phase_b1(p, sm, p.partials, p.nslots, false, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(VJF_NT, 1) vjf_phase_b_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) float sm[];
  unsigned fin = 0;
  for (int i = 0; i < 3; ++i)
    if (isfinite(p.reduced[p.ps + i])) fin |= 1u << i;
  const unsigned need = base_masks(p);
  // without the in-kernel redo a non-finite term cannot be separated from the gradient: skip the SGD
  // step (the reference's "RuntimeError -> skip" branch, vjf/model.py:212-214) and flag it
  if ((fin & need) == need) sgd_from_reduced(p, blockIdx.x, gridDim.x);
  if (blockIdx.x == 0) {
    __syncthreads();
    phase_b2(p, sm, 0, fin);
  }
}


// The two halves of phase B as separate launches (wide-observation path, wide.cu): the SGD step on every CTA, and the serial
// part (losses, running variances, RLS: one CTA) on a side stream beside the next step's forward GEMM.
__global__ void __launch_bounds__(VJF_NT, 1) vjf_sgd_kernel(const __grid_constant__ StepParams p) {
  unsigned fin = 0;
  for (int i = 0; i < 3; ++i)
    if (isfinite(p.reduced[p.ps + i])) fin |= 1u << i;
  const unsigned need = base_masks(p);
  if ((fin & need) == need) sgd_from_reduced(p, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(VJF_NT, 1) vjf_rls_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) float sm[];
  unsigned fin = 0;
  for (int i = 0; i < 3; ++i)
    if (isfinite(p.reduced[p.ps + i])) fin |= 1u << i;
  phase_b2(p, sm, 0, fin);
}
