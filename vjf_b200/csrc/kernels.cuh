// Kernels defined in their own translation units (k_persistent.cu, k_split.cu) and launched from step.cu / aux.cu.
#pragma once
#include "common.cuh"

__global__ void __launch_bounds__(VJF_NT, 1) vjf_persistent_kernel(const __grid_constant__ StepParams p);
__global__ void __launch_bounds__(VJF_NT, 1) vjf_phase_a_kernel(const __grid_constant__ StepParams p);
__global__ void __launch_bounds__(VJF_NT, 1) vjf_reduce_kernel(const __grid_constant__ StepParams p);
__global__ void __launch_bounds__(VJF_NT, 1) vjf_phase_b_kernel(const __grid_constant__ StepParams p);
