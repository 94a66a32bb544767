// Rows next to the hot path: whole-trajectory RLS re-initialisation and forecast.
#include <algorithm>

#include "step_kernels.cuh"

// ------------------------------------------------------------------------------------------
// RBFDS.initialize / LinearRegression.initialize  (vjf/model.py:379-388, vjf/module.py:144-150)
// The reference evaluates velocity() on all N=(T-1)B samples, which forms an (N,N) matrix
// (module.py:76); here phi^T phi, phi^T dx and sum dx^2 are streamed over N in tiles.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VJF_NT, 1)
vjf_rls_stats_kernel(const __grid_constant__ StepParams p, const float* xs, const float* xt, const float* uu, long long N, int xt_is_target) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = p.d, u = p.u, du = p.du, R = p.R, Rp = p.Rp;
  float* phi_s = sm + p.s_phi; float* xu_s = sm + p.s_xu; float* dx_s = sm + p.s_dx;
  float* c_s = sm + p.s_c; float* iw_s = sm + p.s_iw; float* red_s = sm + p.s_red;
  // shared parameters of the RBF features
  for (int i = tid; i < p.R * p.du; i += VJF_NT) c_s[i] = p.state[p.lay.centroid + i];
  for (int i = tid; i < p.R; i += VJF_NT) { const float w = expf(p.state[p.lay.logwidth + i]); iw_s[i] = -0.5f / (w * w); }
  __syncthreads();
  float* slot = p.partials + (size_t)blockIdx.x * p.PS;
  const int TB = VJF_TB_MAX;
  const long long ntiles = (N + TB - 1) / TB;
  bool first = true;
  float sdx = 0.f;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = tile * TB;
    const int nb = (N - b0 < (long long)TB) ? (int)(N - b0) : TB;
    const int rows = (nb + 3) & ~3;
    for (int i = tid; i < rows * du; i += VJF_NT) {
      const int b = i / du, k = i - b * du;
      float v = 0.f;
      if (b < nb) v = (k < d) ? xs[(b0 + b) * d + k] : uu[(b0 + b) * u + (k - d)];
      xu_s[i] = v;
    }
    for (int i = tid; i < rows * d; i += VJF_NT) {
      const int b = i / d;
      float v = 0.f;
      if (b < nb) { v = xt_is_target ? xt[b0 * d + i] : xt[b0 * d + i] - xs[b0 * d + i]; sdx = fmaf(v, v, sdx); }
      dx_s[i] = v;
    }
    __syncthreads();
    for (int i = tid; i < rows * Rp; i += VJF_NT) {
      const int b = i / Rp, k = i - b * Rp;
      float v = 0.f;
      if (k < R && b < nb) {
        float d2 = 0.f;
        for (int c = 0; c < du; ++c) { const float df = xu_s[b * du + c] - c_s[k * du + c]; d2 = fmaf(df, df, d2); }
        v = expf(d2 * iw_s[k]);
      }
      phi_s[i] = v;
    }
    __syncthreads();
    float* Ap = slot + p.pa;
    const int kb = (R + 3) >> 2;  // column blocks of 4 that hold at least one real column (phi rows are padded to Rp >= 4 kb)
    for (int it = tid; it < R * kb; it += VJF_NT) {
      const int kp = it % R, k0 = (it / R) << 2;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int b = 0; b < rows; ++b) {
        const float4 x = *reinterpret_cast<const float4*>(phi_s + b * Rp + k0);
        const float g = phi_s[b * Rp + kp];
        a0 = fmaf(x.x, g, a0); a1 = fmaf(x.y, g, a1); a2 = fmaf(x.z, g, a2); a3 = fmaf(x.w, g, a3);
      }
      float* o = Ap + (size_t)k0 * R + kp;
      acc_store(o, a0, first);
      if (k0 + 1 < R) acc_store(o + R, a1, first);
      if (k0 + 2 < R) acc_store(o + 2 * R, a2, first);
      if (k0 + 3 < R) acc_store(o + 3 * R, a3, first);
    }
    float* bp = slot + p.pb;
    for (int i = tid; i < R * d; i += VJF_NT) {
      const int r = i / d, k = i - r * d;
      float s = 0.f;
      for (int b = 0; b < nb; ++b) s = fmaf(phi_s[b * Rp + r], dx_s[b * d + k], s);
      acc_store(bp + i, s, first);
    }
    first = false;
    __syncthreads();
  }
  if (first) {
    for (int i = p.pa + tid; i < p.PS; i += VJF_NT) slot[i] = 0.f;
  }
  const float s = warp_sum(sdx);
  if (lane == 0) red_s[warp] = s;
  __syncthreads();
  if (tid < VJF_NSCAL) {
    float tot = 0.f;
    if (tid == SC_SDX) for (int w = 0; w < VJF_NWARP; ++w) tot += red_s[w];
    slot[p.ps + tid] = tot;
  }
}

__global__ void __launch_bounds__(VJF_NT, 1) vjf_rls_finish_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) float sm[];
  phase_b2(p, sm, 0, 7u);
}

int vjf_internal_reduce(const StepParams& p, cudaStream_t s);  // step.cu

// shared-memory plan of the streaming statistics kernel (a full 32-row tile: phi / xu / dx / shared RBF parameters) and of phase B2
static int plan_stats(vjf_handle* h, StepParams& p) {
  {
    size_t off = 0;
    auto take = [&](size_t n) { size_t at = off; off = (off + n + 3) & ~(size_t)3; return (int)at; };
    p.s_phi = take((size_t)VJF_TB_MAX * p.Rp);
    p.s_xu = take((size_t)VJF_TB_MAX * p.du);
    p.s_dx = take((size_t)VJF_TB_MAX * p.d);
    p.U_in_smem = 0; p.s_U = 0; p.dec_in_smem = 0; p.s_dec = 0;
    p.s_W = take((size_t)p.R * p.d);
    p.s_c = take((size_t)p.R * p.du);
    p.s_iw = take((size_t)p.R);
    p.s_red = take(VJF_NWARP * VJF_NSCAL + 64);
    size_t b2 = 2 * 16 * 17 + p.R + 4 + 3 * ((size_t)p.d * p.R + 4) + 16 + 2 * VJF_NWARP + 8 + (size_t)p.R * (p.R | 1) + 64;
    if (p.R > 128) b2 = (size_t)(2 * p.R + p.d) * p.ldm + ((p.R + 3) & ~3) + 4 + 2 * VJF_NWARP + 8;
    if (p.rls64) b2 = std::max(b2, vjf_rls64_floats(p));
    p.s_total = (int)std::max(std::max(off, b2), (size_t)1024);
    if ((size_t)p.s_total * 4 > h->smem_limit) { vjf_set_error("n_rbf=%d too large for this build", p.R); return -1; }
  }
  if (!h->aux_attr_set) {  // per handle, i.e. per device
    VJF_CUDA_OK(cudaFuncSetAttribute(vjf_rls_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
    VJF_CUDA_OK(cudaFuncSetAttribute(vjf_rls_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
    h->aux_attr_set = 1;
  }
  return 0;
}

extern "C" int vjf_rls_initialize(vjf_handle* h, int64_t N, const float* xs, const float* xt, const float* u, void* stream) {
  if (!h || !xs || !xt || N < 1) { vjf_set_error("bad argument"); return -1; }
  if (h->cfg.udim > 0 && !u) { vjf_set_error("udim=%d but u is NULL", h->cfg.udim); return -1; }
  if (N >= ((int64_t)1 << 31)) { vjf_set_error("N too large"); return -1; }
  StepParams p = h->base;
  if (plan_stats(h, p)) return -1;
  const int ntiles = (int)((N + VJF_TB_MAX - 1) / VJF_TB_MAX);
  p.nslots = std::min(ntiles, h->max_slots);
  p.Bglobal = (int)N; p.B = (int)N;
  p.flags = VJF_FLAG_UPDATE; p.init_mode = 1; p.red_begin = p.pa; p.T = 1; p.losses = nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)p.s_total * sizeof(float);
  vjf_rls_stats_kernel<<<p.nslots, VJF_NT, smem, s>>>(p, xs, xt, u, (long long)N, 0);
  if (vjf_internal_reduce(p, s)) return -2;
  vjf_rls_finish_kernel<<<1, VJF_NT, smem, s>>>(p);
  g_vjf_launches += 2;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// LinearRegression.kalman  (vjf/module.py:114-142): weight-space Kalman update with diffusion.
// The reference builds H = phi (N x R), R_obs = v I_N and runs kalman.predict + kalman.joseph_update AS WRITTEN
// (vjf/kalman.py:15-50, :102-145; S^-1 applied twice), whose innovation covariance S is N x N.  With M = H Lhat and
// f(M M^T + v I) M = M f(M^T M + v I) every N x N object collapses onto R x R ones built from the streamed statistics
// A = phi^T phi and b = phi^T target:
//   Vhat = U U^T + q I,  Lhat = chol(Vhat)                               (predict; U = w_chol, any square root)
//   Am = Lhat^T A Lhat,  C = Am + v I,  Z = C^-2 Am
//   w_mean' = W + Lhat C^-2 Lhat^T (b - A W)
//   V' = Lhat [ (I - Z)(I - Z)^T + v Z C^-2 ] Lhat^T,   w_chol' = chol(V')   (lower triangular, as linalg.cholesky returns)
// One CTA, fp64, matrices in a global workspace (R <= a few hundred: ~10 R^3 flops, not a hot path).
// ------------------------------------------------------------------------------------------
namespace wk {
__device__ __forceinline__ void sync() { __syncthreads(); }
// C = op(A) * op(B) (+ alpha on the diagonal), all n x n row-major
template <bool TA, bool TB>
__device__ void mm(double* Cm, const double* A, const double* B, int n, double diag = 0.0) {
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    const int r = i / n, c = i - r * n;
    double s = (r == c) ? diag : 0.0;
    for (int k = 0; k < n; ++k) s = fma(TA ? A[k * n + r] : A[r * n + k], TB ? B[c * n + k] : B[k * n + c], s);
    Cm[i] = s;
  }
  sync();
}
// in-place lower Cholesky (upper triangle zeroed); returns false on a non-positive pivot
__device__ bool chol(double* A, int n, int* flag) {
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      double s = A[j * n + j];
      for (int k = 0; k < j; ++k) s -= A[j * n + k] * A[j * n + k];
      if (!(s > 0.0)) *flag = 1;
      A[j * n + j] = sqrt(s);
    }
    sync();
    if (*flag) return false;
    const double dj = A[j * n + j];
    for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / dj;
      A[j * n + i] = 0.0;
    }
    sync();
  }
  return true;
}
// X = (L L^T)^-1 (n x n), one column per thread
__device__ void chol_inverse(double* X, const double* L, int n) {
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    for (int i = 0; i < n; ++i) {  // forward: L z = e_c
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * X[k * n + c];
      X[i * n + c] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {  // backward: L^T x = z
      double s = X[i * n + c];
      for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * X[k * n + c];
      X[i * n + c] = s / L[i * n + i];
    }
  }
  sync();
}
}  // namespace wk

__global__ void __launch_bounds__(VJF_NT, 1)
vjf_weight_kalman_kernel(const __grid_constant__ StepParams p, double* ws, double v, double q) {
  const int R = p.R, d = p.d, n2 = R * R;
  __shared__ int flag;
  if (threadIdx.x == 0) flag = 0;
  __syncthreads();
  double* A = ws; double* Lh = ws + n2; double* T1 = ws + 2 * n2; double* Am = ws + 3 * n2; double* Ci = ws + 4 * n2;
  double* Ci2 = ws + 5 * n2; double* Z = ws + 6 * n2; double* T2 = ws + 7 * n2; double* rhs = ws + 8 * n2; double* t3 = rhs + R * d;
  float* st = p.state;
  const float* U = st + p.lay.w_chol;
  const float* W = st + p.lay.w_mean;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) { A[i] = (double)p.reduced[p.pa + i]; T1[i] = (double)U[i]; }
  __syncthreads();
  wk::mm<false, true>(Lh, T1, T1, R, q);  // Vhat = U U^T + q I
  if (!wk::chol(Lh, R, &flag)) { if (threadIdx.x == 0) atomicOr(p.status, VJF_ST_CHOL_FAILED); return; }
  wk::mm<false, false>(T1, A, Lh, R);      // A Lhat
  wk::mm<true, false>(Am, Lh, T1, R);      // Am = Lhat^T A Lhat
  for (int i = threadIdx.x; i < n2; i += blockDim.x) T2[i] = Am[i] + ((i / R == i % R) ? v : 0.0);  // C
  __syncthreads();
  if (!wk::chol(T2, R, &flag)) { if (threadIdx.x == 0) atomicOr(p.status, VJF_ST_CHOL_FAILED); return; }
  wk::chol_inverse(Ci, T2, R);             // C^-1
  wk::mm<false, false>(Ci2, Ci, Ci, R);    // C^-2
  wk::mm<false, false>(Z, Ci2, Am, R);     // Z = C^-2 Am
  // rhs = Lhat^T (b - A W)  (R x d)
  for (int i = threadIdx.x; i < R * d; i += blockDim.x) {
    const int r = i / d, c = i - r * d;
    double s = (double)p.reduced[p.pb + i];
    for (int k = 0; k < R; ++k) s -= A[r * R + k] * (double)W[k * d + c];
    t3[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * d; i += blockDim.x) {
    const int r = i / d, c = i - r * d;
    double s = 0.0;
    for (int k = 0; k < R; ++k) s += Lh[k * R + r] * t3[k * d + c];
    rhs[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * d; i += blockDim.x) {  // t3 = C^-2 rhs
    const int r = i / d, c = i - r * d;
    double s = 0.0;
    for (int k = 0; k < R; ++k) s += Ci2[r * R + k] * rhs[k * d + c];
    t3[i] = s;
  }
  __syncthreads();
  // inner = (I - Z)(I - Z)^T + v Z C^-2   -> T1
  wk::mm<false, false>(T1, Z, Ci2, R);
  for (int i = threadIdx.x; i < n2; i += blockDim.x) T2[i] = ((i / R == i % R) ? 1.0 : 0.0) - Z[i];
  __syncthreads();
  wk::mm<false, true>(A, T2, T2, R);
  for (int i = threadIdx.x; i < n2; i += blockDim.x) A[i] += v * T1[i];
  __syncthreads();
  wk::mm<false, false>(T1, Lh, A, R);
  wk::mm<false, true>(T2, T1, Lh, R);      // V' = Lhat inner Lhat^T
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {  // symmetrise before the factorisation
    const int r = i / R, c = i - r * R;
    if (c < r) { const double m = 0.5 * (T2[i] + T2[c * R + r]); A[i] = m; A[c * R + r] = m; } else if (c == r) A[i] = T2[i];
  }
  __syncthreads();
  if (!wk::chol(A, R, &flag)) { if (threadIdx.x == 0) atomicOr(p.status, VJF_ST_CHOL_FAILED); return; }
  for (int i = threadIdx.x; i < n2; i += blockDim.x) st[p.lay.w_chol + i] = (float)A[i];
  for (int i = threadIdx.x; i < R * d; i += blockDim.x) {  // w_mean' = W + Lhat t3
    const int r = i / d, c = i - r * d;
    double s = (double)W[i];
    for (int k = 0; k < R; ++k) s += Lh[r * R + k] * t3[k * d + c];
    rhs[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * d; i += blockDim.x) st[p.lay.w_mean + i] = (float)rhs[i];
}

extern "C" int vjf_weight_kalman(vjf_handle* h, int64_t N, const float* xs, const float* target, const float* u, float v,
                                 float diffusion, void* stream) {
  if (!h || !xs || !target || N < 1) { vjf_set_error("bad argument"); return -1; }
  if (h->cfg.udim > 0 && !u) { vjf_set_error("udim=%d but u is NULL", h->cfg.udim); return -1; }
  if (!(diffusion >= 0.f)) { vjf_set_error("diffusion needs to be non-negative"); return -1; }  // module.py:127
  if (!(v > 0.f)) { vjf_set_error("noise variance v must be positive"); return -1; }
  if (N >= ((int64_t)1 << 31)) { vjf_set_error("N too large"); return -1; }
  StepParams p = h->base;
  if (plan_stats(h, p)) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t need = ((size_t)8 * p.R * p.R + 2 * (size_t)p.R * p.d) * sizeof(double);
  if (h->wk_ws_sz < need) {  // grow-only workspace owned by the handle
    VJF_CUDA_OK(cudaStreamSynchronize(s));
    if (h->wk_ws) cudaFree(h->wk_ws);
    h->wk_ws = nullptr; h->wk_ws_sz = 0;
    VJF_CUDA_OK(cudaMalloc(&h->wk_ws, need));
    h->wk_ws_sz = need;
  }
  const int ntiles = (int)((N + VJF_TB_MAX - 1) / VJF_TB_MAX);
  p.nslots = std::min(ntiles, h->max_slots);
  p.Bglobal = (int)N; p.B = (int)N;
  p.flags = VJF_FLAG_UPDATE; p.init_mode = 1; p.red_begin = p.pa; p.T = 1; p.losses = nullptr;
  const size_t smem = (size_t)p.s_total * sizeof(float);
  vjf_rls_stats_kernel<<<p.nslots, VJF_NT, smem, s>>>(p, xs, target, u, (long long)N, 1);
  if (vjf_internal_reduce(p, s)) return -2;
  vjf_weight_kalman_kernel<<<1, VJF_NT, 0, s>>>(p, h->wk_ws, (double)v, (double)diffusion);
  g_vjf_launches += 2;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// forecast: RBFDS.forecast with sampled weights (vjf/model.py:342-361, vjf/module.py:70-73)
// ------------------------------------------------------------------------------------------
// w_t = W + w_chol eps_t   (module.py:71)
__global__ void vjf_forecast_weights_kernel(const float* st, Lay lay, int R, int d, const float* w_eps, float* w_all) {
  const int t = blockIdx.x;
  const float* U = st + lay.w_chol;
  const float* W = st + lay.w_mean;
  const float* e = w_eps + (size_t)t * R * d;
  for (int i = threadIdx.x; i < R * d; i += blockDim.x) {
    const int r = i / d, c = i - r * d;
    float s = W[i];
    for (int k = 0; k < R; ++k) s = fmaf(U[r * R + k], e[k * d + c], s);
    w_all[(size_t)t * R * d + i] = s;
  }
}

// one warp per trial: x[t+1] = x[t] + phi([x[t], u[t]]) w_t (+ state noise)   (model.py:357-359)
__global__ void vjf_forecast_rollout_kernel(const float* st, Lay lay, int R, int d, int u, int n_step, int B, float* x,
                                            const float* uu, const float* w_all, const float* x_eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int b = warp, du = d + u;
  const float* cen = st + lay.centroid;
  const float* lw = st + lay.logwidth;
  const float s_noise = expf(0.5f * st[lay.tr_logvar]);
  float xc[VJF_MAX_XDIM];
#pragma unroll
  for (int k = 0; k < VJF_MAX_XDIM; ++k) xc[k] = (k < d) ? x[(size_t)b * d + k] : 0.f;
  for (int t = 0; t < n_step; ++t) {
    const float* w = w_all + (size_t)t * R * d;
    float acc[VJF_MAX_XDIM];
#pragma unroll
    for (int k = 0; k < VJF_MAX_XDIM; ++k) acc[k] = 0.f;
    for (int r = lane; r < R; r += 32) {
      float d2 = 0.f;
#pragma unroll
      for (int k = 0; k < VJF_MAX_XDIM; ++k)
        if (k < d) { const float df = xc[k] - cen[r * du + k]; d2 = fmaf(df, df, d2); }
      for (int k = 0; k < u; ++k) { const float df = uu[((size_t)t * B + b) * u + k] - cen[r * du + d + k]; d2 = fmaf(df, df, d2); }
      const float wd = expf(lw[r]);
      const float ph = expf(-0.5f * d2 / (wd * wd));
#pragma unroll
      for (int k = 0; k < VJF_MAX_XDIM; ++k)
        if (k < d) acc[k] = fmaf(ph, w[r * d + k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < VJF_MAX_XDIM; ++k)
      if (k < d) {
        float v = xc[k] + warp_sum(acc[k]);
        if (x_eps) v += x_eps[((size_t)t * B + b) * d + k] * s_noise;
        xc[k] = v;
        if (lane == 0) x[((size_t)(t + 1) * B + b) * d + k] = v;
      }
  }
}

// y = decoder(x)  (model.py:323)
__global__ void vjf_decode_kernel(const float* st, Lay lay, int D, int d, long long n, const float* x, float* y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * D) return;
  const long long r = i / D;
  const int j = (int)(i - r * D);
  float s = st[lay.dec_b + j];
  for (int k = 0; k < d; ++k) s = fmaf(st[lay.dec_w + k * D + j], x[r * d + k], s);
  y[i] = s;
}

extern "C" int vjf_forecast(vjf_handle* h, int32_t n_step, int32_t B, float* x, float* yhat, const float* u,
                            const float* w_eps, const float* x_eps, void* stream) {
  if (!h || !x || !w_eps || n_step < 1 || B < 1) { vjf_set_error("bad argument"); return -1; }
  if (h->cfg.udim > 0 && !u) { vjf_set_error("udim=%d but u is NULL", h->cfg.udim); return -1; }
  const StepParams& p = h->base;
  cudaStream_t s = (cudaStream_t)stream;
  // sampled weights of every step: a grow-only workspace owned by the handle (nothing is allocated per call once it is large enough)
  const size_t need = (size_t)n_step * p.R * p.d * sizeof(float);
  if (h->fc_w_sz < need) {
    VJF_CUDA_OK(cudaStreamSynchronize(s));
    if (h->fc_w) cudaFree(h->fc_w);
    h->fc_w = nullptr; h->fc_w_sz = 0;
    VJF_CUDA_OK(cudaMalloc(&h->fc_w, need));
    h->fc_w_sz = need;
  }
  float* w_all = h->fc_w;
  vjf_forecast_weights_kernel<<<n_step, 256, 0, s>>>(h->state, p.lay, p.R, p.d, w_eps, w_all);
  const int wpb = 4;
  vjf_forecast_rollout_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, s>>>(h->state, p.lay, p.R, p.d, p.u, n_step, B, x, u, w_all, x_eps);
  g_vjf_launches += 2;
  if (yhat) {
    const long long n = (long long)(n_step + 1) * B;
    vjf_decode_kernel<<<(unsigned)((n * p.D + 255) / 256), 256, 0, s>>>(h->state, p.lay, p.D, p.d, n, x, yhat);
    ++g_vjf_launches;
  }
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}
