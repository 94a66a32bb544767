// Rows next to the hot path: whole-trajectory RLS re-initialisation and forecast.
#include <algorithm>

#include "step_kernels.cuh"

// ------------------------------------------------------------------------------------------
// RBFDS.initialize / LinearRegression.initialize  (vjf/model.py:379-388, vjf/module.py:144-150)
// The reference evaluates velocity() on all N=(T-1)B samples, which forms an (N,N) matrix
// (module.py:76); here phi^T phi, phi^T dx and sum dx^2 are streamed over N in tiles.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VJF_NT, 1)
vjf_rls_stats_kernel(const __grid_constant__ StepParams p, const float* xs, const float* xt, const float* uu, long long N) {
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
      if (b < nb) { v = xt[b0 * d + i] - xs[b0 * d + i]; sdx = fmaf(v, v, sdx); }
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

extern "C" int vjf_rls_initialize(vjf_handle* h, int64_t N, const float* xs, const float* xt, const float* u, void* stream) {
  if (!h || !xs || !xt || N < 1) { vjf_set_error("bad argument"); return -1; }
  if (h->cfg.udim > 0 && !u) { vjf_set_error("udim=%d but u is NULL", h->cfg.udim); return -1; }
  if (N >= ((int64_t)1 << 31)) { vjf_set_error("N too large"); return -1; }
  StepParams p = h->base;
  // smem plan of a full 32-row tile (only phi / xu / dx / shared parameters are used) and of phase B2
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
    p.s_total = (int)std::max(std::max(off, b2), (size_t)1024);
    if ((size_t)p.s_total * 4 > h->smem_limit) { vjf_set_error("n_rbf=%d too large for this build", p.R); return -1; }
  }
  if (!h->aux_attr_set) {  // per handle, i.e. per device
    VJF_CUDA_OK(cudaFuncSetAttribute(vjf_rls_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
    VJF_CUDA_OK(cudaFuncSetAttribute(vjf_rls_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
    h->aux_attr_set = 1;
  }
  const int ntiles = (int)((N + VJF_TB_MAX - 1) / VJF_TB_MAX);
  p.nslots = std::min(ntiles, h->max_slots);
  p.Bglobal = (int)N; p.B = (int)N;
  p.flags = VJF_FLAG_UPDATE; p.init_mode = 1; p.red_begin = p.pa; p.T = 1; p.losses = nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)p.s_total * sizeof(float);
  vjf_rls_stats_kernel<<<p.nslots, VJF_NT, smem, s>>>(p, xs, xt, u, (long long)N);
  if (vjf_internal_reduce(p, s)) return -2;
  vjf_rls_finish_kernel<<<1, VJF_NT, smem, s>>>(p);
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
