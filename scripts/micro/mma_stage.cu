// Microbenchmark of the phase-A tile GEMMs in isolation (C2 shapes): cycles per call, one CTA of 512 threads.
#include <cstdio>
#include "../../vjf_b200/csrc/mma.cuh"
void vjf_set_error(const char*, ...) {}
long long g_vjf_launches = 0;
__global__ void __launch_bounds__(512, 1) k(float* W1, float* bias, float* slot, long long* cyc, int reps) {
  extern __shared__ __align__(16) float sm[];
  const int K1 = 206, K1p = 228, H = 64, Hp = 68, Gp = 72, R = 50, Rp = 68, ldu = 56, rows = 32, ldw = 72;
  float* in_s = sm; float* act = in_s + rows * K1p; float* G = act + rows * Hp; float* phi = G + rows * Gp; float* U = phi + rows * Rp;
  float* qp = U + 56 * ldu; float* W1s = qp + 512; float* inl = W1s + K1 * ldw; float* Gl = inl + rows * K1p; float* phil = Gl + rows * Gp;
  for (int i = threadIdx.x; i < rows * K1p; i += 512) in_s[i] = (i % 7) * 0.1f;
  for (int i = threadIdx.x; i < rows * Gp; i += 512) G[i] = (i % 5) * 0.01f;
  for (int i = threadIdx.x; i < rows * Rp; i += 512) phi[i] = (i % 3) * 0.2f;
  for (int i = threadIdx.x; i < 56 * ldu; i += 512) U[i] = (i % 11) * 0.05f;
  for (int i = threadIdx.x; i < K1 * ldw; i += 512) W1s[i] = (i % 13) * 0.01f;
  for (int i = threadIdx.x; i < rows * K1p; i += 512) inl[i] = 1e-4f * (i % 3);
  for (int i = threadIdx.x; i < rows * Gp; i += 512) Gl[i] = 1e-5f;
  for (int i = threadIdx.x; i < rows * Rp; i += 512) phil[i] = 1e-5f;
  __syncthreads();
  long long t[6];
  for (int r = 0; r < reps; ++r) {
    t[0] = clock64();
    mma_linear_fwd(in_s, inl, K1p, K1, W1s, ldw, bias, H, act, Hp, rows, true);
    __syncthreads(); t[1] = clock64();
    mma_linear_fwd(in_s, nullptr, K1p, K1, W1s, ldw, bias, H, act, Hp, rows, true);
    __syncthreads(); t[2] = clock64();
    mma_wgrad(in_s, inl, K1p, K1, G, Gl, Gp, H, rows, slot, true);
    __syncthreads(); t[3] = clock64();
    mma_quadform(phi, phil, Rp, U, ldu, R, rows, qp, 1);
    __syncthreads(); t[4] = clock64();
    mma_gram(phi, phil, Rp, R, rows, slot + 20000, true);
    __syncthreads(); t[5] = clock64();
  }
  if (threadIdx.x == 0) for (int i = 0; i < 5; ++i) cyc[i] = t[i + 1] - t[i];
}
int main() {
  float *W1, *bias, *slot; long long* cyc;
  cudaMalloc(&W1, 206 * 64 * 4); cudaMalloc(&bias, 256); cudaMalloc(&slot, 1 << 20); cudaMalloc(&cyc, 64);
  cudaMemset(W1, 0, 206 * 64 * 4); cudaMemset(bias, 0, 256);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int it = 0; it < 2; ++it) { k<<<1, 512, 200 * 1024>>>(W1, bias, slot, cyc, 3); cudaDeviceSynchronize(); }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  long long h[5]; cudaMemcpy(h, cyc, 40, cudaMemcpyDeviceToHost);
  const char* n[5] = {"linear_fwd (W1 in smem)", "linear_fwd (A split on the fly)", "wgrad", "quadform", "gram"};
  for (int i = 0; i < 5; ++i) printf("%-26s %lld cycles (%.2f us @1.9GHz)\n", n[i], h[i], h[i] / 1900.0);
  return 0;
}
