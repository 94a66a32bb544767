// Microbenchmark of the LDL^T sweep inner loop: which part of the per-column iteration costs what.
#include <cstdio>
#include <cuda_runtime.h>
template <int RPW, int CPL, int NW, int VARIANT>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int R) {
  __shared__ float colbuf[2][NW * RPW + 64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float v[RPW][CPL];
  for (int ri = 0; ri < RPW; ++ri) for (int ci = 0; ci < CPL; ++ci) v[ri][ci] = 1.0f + 0.001f * (tid + ri + ci);
  int jj[CPL];
  for (int ci = 0; ci < CPL; ++ci) { int j = lane + 32 * ci; jj[ci] = j < R ? j : -1; }
  if (tid < NW * RPW) { colbuf[0][tid] = 2.0f + tid * 1e-3f; colbuf[1][tid] = 2.0f; }
  __syncthreads();
  long long t0 = clock64();
  int kb = 0;
  if (warp < NW) {
    for (int rep = 0; rep < 20; ++rep)
    for (int k = 0; k < R; ++k) {
      const float* cb = colbuf[kb];
      float* cbn = colbuf[kb ^ 1];
      const float piv = cb[k];
      float tk[RPW], cj[CPL];
#pragma unroll
      for (int ri = 0; ri < RPW; ++ri) tk[ri] = cb[warp + NW * ri];
#pragma unroll
      for (int ci = 0; ci < CPL; ++ci) { const float x = cb[lane + 32 * ci]; cj[ci] = (jj[ci] > k) ? x : 0.f; }
      float ninv;
      if (VARIANT == 1) ninv = -piv;                // no reciprocal
      else if (VARIANT == 2) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(piv)); ninv = -r; }
      else ninv = -__frcp_rn(piv);
#pragma unroll
      for (int ri = 0; ri < RPW; ++ri) tk[ri] *= ninv * 1e-3f;
#pragma unroll
      for (int ri = 0; ri < RPW; ++ri)
#pragma unroll
        for (int ci = 0; ci < CPL; ++ci) v[ri][ci] = fmaf(tk[ri], cj[ci], v[ri][ci]);
      const int kn = k + 1;
      if (VARIANT != 3) {
        if (lane == (kn & 31)) {
          const int cn = kn >> 5;
#pragma unroll
          for (int ri = 0; ri < RPW; ++ri) {
            float x = v[ri][0];
#pragma unroll
            for (int ci = 1; ci < CPL; ++ci) x = (cn == ci) ? v[ri][ci] : x;
            cbn[warp + NW * ri] = 2.0f + 1e-6f * x;
          }
        }
      }
      if (VARIANT == 4) __syncthreads();
      else if (VARIANT == 5) { /* no barrier at all */ }
      else asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
      kb ^= 1;
    }
  }
  long long t1 = clock64();
  float s = 0; for (int ri = 0; ri < RPW; ++ri) for (int ci = 0; ci < CPL; ++ci) s += v[ri][ci];
  out[tid] = s;
  if (tid == 0) *cyc = (t1 - t0);
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const int R = 50;
  auto run = [&](const char* name, auto kern) {
    kern<<<1, 512>>>(out, cyc, R); cudaDeviceSynchronize();
    kern<<<1, 512>>>(out, cyc, R); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %.0f cycles / column\n", name, (double)h / (20.0 * R));
  };
  run("16 warps x7 rows, frcp_rn, named bar", k<7, 2, 16, 0>);
  run("16 warps x7 rows, no rcp", k<7, 2, 16, 1>);
  run("16 warps x7 rows, rcp.approx", k<7, 2, 16, 2>);
  run("16 warps x7 rows, no publish", k<7, 2, 16, 3>);
  run("16 warps x7 rows, __syncthreads", k<7, 2, 16, 4>);
  run("16 warps x7 rows, no barrier", k<7, 2, 16, 5>);
  run("8 warps x13 rows, frcp_rn, named bar", k<13, 2, 8, 0>);
  run("8 warps x13 rows, no barrier", k<13, 2, 8, 5>);
  run("4 warps x26 rows, frcp_rn, named bar", k<26, 2, 4, 0>);
  run("4 warps x26 rows, rcp.approx", k<26, 2, 4, 2>);
  run("4 warps x26 rows, no barrier", k<26, 2, 4, 5>);
  run("2 warps x52 rows, rcp.approx", k<52, 2, 2, 2>);
  return 0;
}
