// Microbenchmark: LDL^T sweep with shuffle multipliers, 1 column per barrier vs 2 columns per barrier.
#include <cstdio>
#include <cuda_runtime.h>
#define NW 16
template <int RPW, int CPL, int PR, int NB>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int R) {
  __shared__ float colbuf[2][2][NW * RPW + 64];
  __shared__ float dvec[128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float v[RPW][CPL];
  for (int ri = 0; ri < RPW; ++ri) for (int ci = 0; ci < CPL; ++ci) v[ri][ci] = 1.0f + 0.001f * (tid + ri + ci);
  int jj[CPL];
  for (int ci = 0; ci < CPL; ++ci) { int j = lane + 32 * ci; jj[ci] = j < R ? j : -1; }
  for (int i = tid; i < 2 * 2 * (NW * RPW + 64); i += 512) (&colbuf[0][0][0])[i] = 2.0f + (i % 7) * 1e-3f;
  __syncthreads();
  long long t0 = clock64();
  int kb = 0;
  for (int rep = 0; rep < 20; ++rep) {
    if (NB == 1) {
      for (int k = 0; k < R; ++k) {
        const float* cb = colbuf[kb][0]; float* cbn = colbuf[kb ^ 1][0];
        const float piv = cb[k];
        float tk[RPW], cj[CPL];
#pragma unroll
        for (int ci = 0; ci < CPL; ++ci) { const float x = cb[lane + 32 * ci]; cj[ci] = (jj[ci] > k) ? x : 0.f; }
#pragma unroll
        for (int ri = 0; ri < RPW; ++ri) tk[ri] = __shfl_sync(0xffffffffu, v[ri][0], k & 31);
        float rinv; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(piv));
        rinv = fmaf(rinv, fmaf(-piv, rinv, 1.0f), rinv);
        if (tid == 0) dvec[k] = piv;
        const float ninv = -rinv * 1e-3f;
#pragma unroll
        for (int ri = 0; ri < RPW; ++ri) tk[ri] *= ninv;
#pragma unroll
        for (int ri = 0; ri < RPW; ++ri)
#pragma unroll
          for (int ci = 0; ci < CPL; ++ci) v[ri][ci] = fmaf(tk[ri], cj[ci], v[ri][ci]);
        if (lane == ((k + 1) & 31)) {
#pragma unroll
          for (int ri = 0; ri < PR; ++ri) cbn[warp + NW * ri] = 2.0f + 1e-6f * v[ri][0];
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        kb ^= 1;
      }
    } else {
      for (int k = 0; k + 1 < R; k += 2) {
        const float* ca = colbuf[kb][0]; const float* cbb = colbuf[kb][1];
        float* na = colbuf[kb ^ 1][0]; float* nb = colbuf[kb ^ 1][1];
        const float a_k = ca[k], a_k1 = ca[k + 1], b_k1 = cbb[k + 1];
        float c1[CPL], c2[CPL];
        float ar[RPW], br[RPW];
#pragma unroll
        for (int ci = 0; ci < CPL; ++ci) { c1[ci] = ca[lane + 32 * ci]; c2[ci] = cbb[lane + 32 * ci]; }
#pragma unroll
        for (int ri = 0; ri < RPW; ++ri) { ar[ri] = __shfl_sync(0xffffffffu, v[ri][0], k & 31); br[ri] = __shfl_sync(0xffffffffu, v[ri][0], (k + 1) & 31); }
        float inv1; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1) : "f"(a_k));
        inv1 = fmaf(inv1, fmaf(-a_k, inv1, 1.0f), inv1);
        const float m = a_k1 * inv1;
        const float d1 = fmaf(-m, a_k1, b_k1) + 3.0f;
        float inv2; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv2) : "f"(d1));
        inv2 = fmaf(inv2, fmaf(-d1, inv2, 1.0f), inv2);
        if (tid == 0) { dvec[k] = a_k; dvec[k + 1] = d1; }
#pragma unroll
        for (int ci = 0; ci < CPL; ++ci) {
          const float aj = c1[ci], bj = c2[ci];
          c1[ci] = (jj[ci] > k) ? aj : 0.f;
          c2[ci] = (jj[ci] > k + 1) ? fmaf(-aj, m, bj) : 0.f;
        }
#pragma unroll
        for (int ri = 0; ri < RPW; ++ri) {
          const float u = ar[ri] * inv1 * 1e-3f;
          const float w = fmaf(-u, a_k1, br[ri]) * inv2 * 1e-3f;
#pragma unroll
          for (int ci = 0; ci < CPL; ++ci) v[ri][ci] = fmaf(-u, c1[ci], fmaf(-w, c2[ci], v[ri][ci]));
        }
        if (lane == ((k + 2) & 31)) {
#pragma unroll
          for (int ri = 0; ri < PR; ++ri) na[warp + NW * ri] = 2.0f + 1e-6f * v[ri][0];
        }
        if (lane == ((k + 3) & 31)) {
#pragma unroll
          for (int ri = 0; ri < PR; ++ri) nb[warp + NW * ri] = 2.0f + 1e-6f * v[ri][0];
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        kb ^= 1;
      }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int ri = 0; ri < RPW; ++ri) for (int ci = 0; ci < CPL; ++ci) s += v[ri][ci];
  out[tid] = s + dvec[3];
  if (tid == 0) *cyc = (t1 - t0);
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const int R = 50;
  auto run = [&](const char* name, auto kern) {
    kern<<<1, 512>>>(out, cyc, R); cudaDeviceSynchronize();
    kern<<<1, 512>>>(out, cyc, R); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %.0f cycles / column, %.2f us per 50-column sweep @1.9GHz\n", name, (double)h / (20.0 * R), (double)h / 20.0 / 1900.0);
  };
  run("1 column / barrier (shuffle multipliers)", k<7, 2, 4, 1>);
  run("2 columns / barrier", k<7, 2, 4, 2>);
  return 0;
}
