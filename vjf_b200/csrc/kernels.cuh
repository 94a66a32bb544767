// Kernels defined in their own translation units (k_persistent.cu, k_split.cu) and launched from step.cu / aux.cu.
#pragma once
#include "common.cuh"

__global__ void __launch_bounds__(VJF_NT, 1) vjf_persistent_kernel(const __grid_constant__ StepParams p);
__global__ void __launch_bounds__(VJF_NT, 1) vjf_phase_a_kernel(const __grid_constant__ StepParams p);
__global__ void __launch_bounds__(VJF_NT, 1) vjf_reduce_kernel(const __grid_constant__ StepParams p);
__global__ void __launch_bounds__(VJF_NT, 1) vjf_phase_b_kernel(const __grid_constant__ StepParams p);

// throughput tile pipeline (k_tile.cu, tile_host.cu)
#include <cuda.h>
__global__ void __launch_bounds__(VJF_NT, 1) vjf_tile_kernel(const __grid_constant__ StepParams p, const __grid_constant__ CUtensorMap ymap);
// host: plan the tile pipeline for a launch (returns 1 when it applies and fills p.tp / the tensor map), launch it
int vjf_tile_plan(vjf_handle* h, StepParams& p, const void* y, int y_dtype, int T, int B, CUtensorMap* map, cudaStream_t stream);
int vjf_tile_launch(vjf_handle* h, StepParams& p, const CUtensorMap& map, cudaStream_t s);
int vjf_tile_create(vjf_handle* h);

// large n_rbf (bigr.cu): per-step launch sequence with the RBF contractions as tcgen05 GEMMs and a blocked multi-CTA factorisation
int vjf_bigr_create(vjf_handle* h);
void vjf_bigr_destroy(vjf_handle* h);
int vjf_bigr_time_loop(vjf_handle* h, StepParams& p, int T, int B, cudaStream_t s);
int vjf_plan_tiles_public(vjf_handle* h, StepParams& p, int B);
int vjf_internal_reduce(const StepParams& p, cudaStream_t s);

// wide observations (wide.cu; ydim above the tile pipeline's limit): per-step launch sequence with the two contractions over the
// observation columns as tcgen05 GEMMs over all trials
int vjf_wide_create(vjf_handle* h);
void vjf_wide_destroy(vjf_handle* h);
bool vjf_wide_applies(vjf_handle* h, const StepParams& p, int B);
int vjf_wide_time_loop(vjf_handle* h, StepParams& p, int T, int B, cudaStream_t s);
// are all observations exactly representable in tf32 (spike counts)?  (tile_host.cu; synchronises the stream)
int vjf_observations_exact(vjf_handle* h, const void* y, size_t n, cudaStream_t s, bool* exact);
int vjf_tile_mode_get();  // vjf_set_tile_mode: 1 = general persistent kernel only (no tile pipeline, no wide-observation path)
