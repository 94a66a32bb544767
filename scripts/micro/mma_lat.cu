// Microbenchmark: latency / throughput of legacy mma.sync (tf32 m16n8k8, bf16 m16n8k16) and FFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void mma_tf32(float c[4], const uint32_t a[4], const uint32_t b[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float c[4], const uint32_t a[4], const uint32_t b[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CHAINS, int KIND>
__global__ void k(float* out, long long* cyc, int iters) {
  uint32_t a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
  float c[CHAINS][4];
  for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { if (KIND == 0) mma_tf32(c[i], a, b); else mma_bf16(c[i], a, b); }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
__global__ void kf(float* out, long long* cyc, int iters) {
  float c[CHAINS]; float a = threadIdx.x * 1e-3f, b = 1.0001f;
  for (int i = 0; i < CHAINS; ++i) c[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = fmaf(c[i], b, a);
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < CHAINS; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  auto run = [&](const char* name, auto kern, int chains, int threads) {
    kern<<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    kern<<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps=%2d chains=%d : %.1f cycles per instr per warp-chain-step, %.2f instr/cycle/SM\n", name, threads / 32, chains,
           (double)h / iters, (double)iters * chains * (threads / 32) / h);
  };
  run("tf32 m16n8k8", k<1, 0>, 1, 32);   run("tf32 m16n8k8", k<2, 0>, 2, 32);  run("tf32 m16n8k8", k<4, 0>, 4, 32);  run("tf32 m16n8k8", k<8, 0>, 8, 32);
  run("tf32 m16n8k8", k<4, 0>, 4, 128);  run("tf32 m16n8k8", k<4, 0>, 4, 512); run("tf32 m16n8k8", k<8, 0>, 8, 512); run("tf32 m16n8k8", k<1, 0>, 1, 512);
  run("bf16 m16n8k16", k<1, 1>, 1, 32);  run("bf16 m16n8k16", k<4, 1>, 4, 32); run("bf16 m16n8k16", k<4, 1>, 4, 512); run("bf16 m16n8k16", k<8, 1>, 8, 512);
  run("ffma", kf<1>, 1, 32); run("ffma", kf<8>, 8, 32); run("ffma", kf<8>, 8, 512); run("ffma", kf<16>, 16, 512);
  return 0;
}
