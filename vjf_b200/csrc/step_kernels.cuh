// Device code of the VJF filter + learning step (sm_100a).  Reference semantics: vjf/model.py:179-221.
//
//   phase A  (all CTAs, trial-parallel)  forward, one-sample ELBO, hand-derived backward, RLS statistics
//            of a tile of trials; every CTA leaves its sums in its own slot of the partial buffer.
//   phase B1 (all CTAs)  deterministic two-level reduction of the slots, gradient clip + SGD on the slice
//            each CTA owns.
//   phase B2 (CTA 0)     loss read-out, running noise variances, RLS: one augmented LDL^T sweep that yields
//            chol(P'), L^-1 g and L^-T together, then W' = L^-T (L^-1 g).
#pragma once
#include "common.cuh"
#include "mma.cuh"
#include "umma.cuh"

// ------------------------------------------------------------------------------------------
// slot accumulation helpers (the tile GEMMs live in mma.cuh / umma.cuh)
// ------------------------------------------------------------------------------------------

// A CTA's slot is written by that CTA only: the first tile stores, later tiles add with a fire-and-forget
// reduction (same thread, same address => program order, so the sum order stays deterministic).
__device__ __forceinline__ void acc_store(float* p, float v, bool first) {
  if (first) *p = v;
  else atomicAdd(p, v);
}

// db[n] (+)= sum_b G[b][n]
__device__ __forceinline__ void tile_colsum(const float* G, const float* Gl, int ldg, int N, int rows, float* db, bool first) {
  for (int n = threadIdx.x; n < N; n += VJF_NT) {
    float s = 0.f;
    for (int b = 0; b < rows; ++b) s += G[b * ldg + n] + Gl[b * ldg + n];
    acc_store(db + n, s, first);
  }
}

// ------------------------------------------------------------------------------------------
// phase A: one tile of trials
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_y(const StepParams& p, size_t idx) {
  return p.y_dtype == VJF_Y_U8 ? (float)reinterpret_cast<const unsigned char*>(p.y)[idx]
                               : reinterpret_cast<const float*>(p.y)[idx];
}

// Decoder + likelihood + their gradients for one tile; DX >= d is the compile-time bound of the state loops.
template <int DX>
static __device__ __forceinline__ void decoder_stage(const StepParams& p, float* sm, int nb, bool first, bool r_on, float lam,
                                                     float p_lam, float e_nlam, const float* dw, const float* db, float* slot,
                                                     float* sc, int t = 0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D, d = p.d, K1p = p.K1p, Dp = p.Dp;
  const float* in_s = sm + p.s_in; float* g_s = sm + p.s_g; const float* xt_s = sm + p.s_xt; float* gxt_s = sm + p.s_gxt;
  for (int b = warp; b < nb; b += VJF_NWARP) {
    float xt[DX], gx[DX];
#pragma unroll
    for (int k = 0; k < DX; ++k) { xt[k] = (k < d) ? xt_s[b * d + k] : 0.f; gx[k] = 0.f; }
    const float* yb = in_s + b * K1p;
    const float* ybl = sm + p.s_inl + b * K1p;
    float* gb = g_s + b * Dp;
    for (int j = lane; j < D; j += 32) {
      float w[DX];
      float eta = db[j];
#pragma unroll
      for (int k = 0; k < DX; ++k) { w[k] = (DX == d || k < d) ? dw[k * D + j] : 0.f; eta = fmaf(w[k], xt[k], eta); }
      const float yv = p.in_split ? yb[j] + ybl[j] : yb[j];  // the staged observations may be a (hi, lo) pair
      float g;
      if (p.lik == VJF_LIK_GAUSSIAN) {
        // gaussian_loss(y, eta, lambda), functional.py:55-75 ; update's mse, likelihood.py:36-37
        const float r = yv - eta;
        const float rs = yv * p_lam - eta * p_lam;
        const float mse = rs * rs;
        if (!isfinite(mse)) sc[SC_BADMSE] += 1.f;
        sc[SC_RECON] += 0.5f * (mse + lam);
        sc[SC_SSE] = fmaf(r, r, sc[SC_SSE]);
        g = -r * e_nlam;
        // d/dlambda = 0.5 (1 - r^2 e^-lambda) ; accumulated into the lik_logvar gradient slot
        sc[6] += r_on ? 0.5f * (1.0f - r * r * e_nlam) : 0.f;
      } else {
        // poisson_nll_loss(clamp(eta, max=10), y, log_input=True), likelihood.py:60-62
        const float ec = fminf(eta, 10.0f);
        const float ex = expf(ec);
        sc[SC_RECON] += ex - yv * ec;
        g = (eta <= 10.0f) ? (ex - yv) : 0.f;
        if (eta != eta) { sc[SC_RECON] = eta; g = eta; }  // NaN propagates like torch.clamp
      }
      g = r_on ? g : 0.f;
      gb[j] = g;
#pragma unroll
      for (int k = 0; k < DX; ++k) gx[k] = fmaf(g, w[k], gx[k]);
    }
#pragma unroll
    for (int k = 0; k < DX; ++k) {
      const float s = warp_sum(gx[k]);
      if (lane == 0 && k < d) gxt_s[b * d + k] = s;
    }
  }
  VJF_STAMP(p, t, 41);
  __syncthreads();
  VJF_STAMP(p, t, 42);
  float* gdw = slot + p.lay.dec_w;
  float* gdb = slot + p.lay.dec_b;
  for (int j = tid; j < D; j += VJF_NT) {
    float accb = 0.f, acc[DX];
#pragma unroll
    for (int k = 0; k < DX; ++k) acc[k] = 0.f;
    for (int b = 0; b < nb; ++b) {
      const float g = g_s[b * Dp + j];
      accb += g;
#pragma unroll
      for (int k = 0; k < DX; ++k)
        if (DX == d || k < d) acc[k] = fmaf(g, xt_s[b * d + k], acc[k]);
    }
    acc_store(gdb + j, accb, first);
#pragma unroll
    for (int k = 0; k < DX; ++k)
      if (k < d) acc_store(gdw + k * D + j, acc[k], first);
  }
}

// masks: bit0 recon term on, bit1 dynamics term on (already combined with !warm_up), bit2 entropy term on.
// part: PART_BOTH runs the whole tile; PART_FRONT / PART_BACK run the halves of the overlapped schedule:
//   front = everything that does not need the RLS outputs of the previous step (observation staging, RBF
//           features, recognition forward, decoder + likelihood + their gradients, RLS statistics)
//   back  = dynamics read-out with the fresh w_mean / w_chol / state noise, ELBO dynamics + entropy terms,
//           backward through the recognition network
#define PART_BOTH 0
#define PART_FRONT 1
#define PART_BACK 2
// Per-CTA context of the overlapped persistent schedule (one tile per trial CTA): the Blackwell asynchronous machinery.
//   CX_UMMA: layer-1 weight gradient on tcgen05 / TMEM (umma.cuh)
//   CX_TMA : recognition layer-1 weight + decoder staged by TMA bulk copies (waited just before they are needed, not at
//            the start of the tile), and the observation tile of step t+1 prefetched (cp.async) at the end of the back
//            half of step t, so that its HBM latency hides behind the reduction / barriers
// mbarriers live in shared memory at s_flag + 2 (umma), + 4 (front weights: w_phase) and + 6 (back: w_chol / w_mean, y_phase);
// *_phase = parity to wait for.
#define CX_UMMA 1
#define CX_TMA 2
struct TileCtx { uint32_t tmem, umma_phase, w_phase, y_phase; int y_ready_t, flags, consts_staged, eps_ready_t, early_ok; };
typedef TileCtx UmmaCtx;

// bytes the front prologue moves by TMA
static __device__ __forceinline__ uint32_t tma_head_bytes(const StepParams& p) { return 16u * (((uint32_t)p.H[p.L - 1] * p.d + 3u) >> 2); }
static __device__ __forceinline__ uint32_t tma_back_bytes(const StepParams& p) {
  return (p.U_in_smem ? (uint32_t)((p.R + 7) & ~7) * p.ldu * 4u : 0u) + 16u * (((uint32_t)p.R * p.d + 3u) >> 2);
}
static __device__ __forceinline__ uint32_t tma_front_bytes(const StepParams& p) {
  return (p.W1_in_smem ? (uint32_t)p.K1 * p.ldw1 * 4u : 0u) + (p.dec_in_smem ? (uint32_t)(p.d + 1) * p.D * 4u : 0u) + 2u * tma_head_bytes(p) +
         16u * (((uint32_t)p.d + 3u) >> 2);
}

// one thread: w_chol (row-padded mirror, pads zero) and w_mean by TMA, completion on the mbarrier the back half waits on
static __device__ __forceinline__ void issue_back_tma(const StepParams& p, float* sm) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + p.s_flag + 6);
  asm volatile("fence.proxy.async.global;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(bar, tma_back_bytes(p));
  if (p.U_in_smem) tma_bulk_g2s(sm + p.s_U, p.u_mirror, (uint32_t)((p.R + 7) & ~7) * p.ldu * 4u, bar);
  tma_bulk_g2s(sm + p.s_W, p.state + p.lay.w_mean, 16u * (((uint32_t)p.R * p.d + 3u) >> 2), bar);
}

static __device__ __forceinline__ void phase_a_tile(const StepParams& p, float* sm, int t, int tile, bool first, unsigned masks, int part,
                                                    TileCtx* cx = nullptr) {
  TileCtx* uc = (cx && (cx->flags & CX_UMMA)) ? cx : nullptr;
  const bool tma = cx && (cx->flags & CX_TMA);
  const bool cx_overlap = p.overlap && part != PART_BOTH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D, d = p.d, u = p.u, R = p.R, du = p.du, E = p.E, L = p.L;
  const int K1 = p.K1, K1p = p.K1p, Rp = p.Rp, Gp = p.Gp;
  const int b0 = tile * p.TB;
  const int nb = min(p.TB, p.B - b0);
  const int rows = (nb + 15) & ~15;
  float* in_s = sm + p.s_in;   float* phi_s = sm + p.s_phi;
  float* inl_s = sm + p.s_inl; float* phil_s = sm + p.s_phil;  // lo parts of the presplit (hi, lo) pairs
  float* gpal = sm + p.s_gpal; float* gpbl = sm + p.s_gpbl;
  float* gpa = sm + p.s_gpa;   float* gpb = sm + p.s_gpb;   float* eps_s = sm + p.s_eps;
  float* xu_s = sm + p.s_xu;   float* xt_s = sm + p.s_xt;   float* mt_s = sm + p.s_mt;
  float* lt_s = sm + p.s_lt;   float* pm_s = sm + p.s_pm;   float* dx_s = sm + p.s_dx;
  float* gxt_s = sm + p.s_gxt; float* gmt_s = sm + p.s_gmt; float* glt_s = sm + p.s_glt;
  float* plv_s = sm + p.s_plv; float* W_s = sm + p.s_W;     float* c_s = sm + p.s_c;
  float* iw_s = sm + p.s_iw;   float* red_s = sm + p.s_red; float* qp_s = sm + p.s_qp;
  float* scf_s = sm + p.s_scf;
  const float* hm_s = sm + p.s_hm; const float* hv_s = sm + p.s_hv;  // head weights [H_L][d], then head_v_b at hv_s[H_L*d]
  int* flag_s = reinterpret_cast<int*>(sm + p.s_flag);
  float* st = p.state;
  const float* dw = p.dec_in_smem ? (sm + p.s_dec) : (st + p.lay.dec_w);        // [d][D]
  const float* db = p.dec_in_smem ? (sm + p.s_dec + d * D) : (st + p.lay.dec_b);  // [D]
  float* slot = p.partials + (size_t)blockIdx.x * p.PS;
  const bool r_on = masks & 1u, d_on = masks & 2u, h_on = masks & 4u;
  const int HL = p.H[L - 1], ldh = p.Hp[L - 1];
  const float* hL = sm + p.s_act[L - 1];
  float sc[VJF_NSCAL];
#pragma unroll
  for (int i = 0; i < VJF_NSCAL; ++i) sc[i] = 0.f;

  if (part != PART_BACK) {
    // ---- S0: stage the tile: in = [y | u | m_s | l_s | 0] (vjf/recognition.py:32-37), eps, zero pads ----
    {
      const size_t row0 = (size_t)t * p.B + b0;
      const int Ep = K1p - D;  // u, m_s, l_s and the zero pad
      const bool prior = (t == 0) && (p.flags & VJF_FLAG_PRIOR_Q0);
      const float* qm = (t == 0) ? p.q0m : p.mu + (size_t)(t - 1) * p.B * d;
      const float* ql = (t == 0) ? p.q0l : p.logvar + (size_t)(t - 1) * p.B * d;
      // observation tile prefetched (cp.async) during the previous back half?  (a prefetch for another step -- the redo
      // path re-runs the front half of the same step -- is drained and dropped)
      bool y_there = false;
      if (tma && cx->y_ready_t >= 0) {
        y_there = (cx->y_ready_t == t);
        cx->y_ready_t = -1;
        if (!y_there) { cp_async_wait_all(); __syncthreads(); }
      }
      if (y_there) {
        // fast path of the overlapped schedule: the observations are already in place (prefetch), the previous
        // posterior of this tile is still in shared memory (mt_s / lt_s of the previous front half), and the rest is
        // spread over all threads instead of one warp per trial
        for (int i = tid; i < nb * Ep; i += VJF_NT) {
          const int b = i / Ep, e = i - b * Ep;
          float v = 0.f;
          if (e < u) v = p.u_in[(row0 + b) * u + e];
          else if (e < u + d) v = mt_s[b * d + e - u];
          else if (e < E) v = lt_s[b * d + e - u - d];
          in_s[b * K1p + D + e] = v;
        }
        if (p.eps) {
          for (int i = tid; i < nb * 2 * d; i += VJF_NT) {
            const int b = i / (2 * d), k = i - b * 2 * d;
            const float* e0 = p.eps + ((size_t)t * 2 * p.B + b0 + b) * d;
            eps_s[i] = (k < d) ? e0[k] : e0[(size_t)p.B * d + k - d];
          }
        } else if (cx->eps_ready_t != t) {  // (normally drawn during the trial-barrier wait, see draw_noise_tile)
          const int nblk = (d + 3) >> 2;
          for (int i = tid; i < nb * 2 * nblk; i += VJF_NT) {
            const int b = i / (2 * nblk), r = i - b * 2 * nblk, which = r / nblk, blk = r - which * nblk;
            float z[4];
            philox_normal4(p.seed, p.step0 + t, p.trial_offset + b0 + b, which, blk, z);
            for (int k = 0; k < 4; ++k)
              if (blk * 4 + k < d) eps_s[b * 2 * d + which * d + blk * 4 + k] = z[k];
          }
        }
      }
      for (int b = (y_there ? nb : 0) + warp; b < rows; b += VJF_NWARP) {
        float* dst = in_s + b * K1p;
        if (b < nb) {
          if ((D & 3) == 0) {
            if (p.y_dtype == VJF_Y_U8) {
              const uchar4* src = reinterpret_cast<const uchar4*>(reinterpret_cast<const unsigned char*>(p.y) + (row0 + b) * D);
              for (int j = lane; j < (D >> 2); j += 32) {
                const uchar4 v = src[j];
                *reinterpret_cast<float4*>(dst + 4 * j) = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
              }
            } else {
              const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.y) + (row0 + b) * D);
              for (int j = lane; j < (D >> 2); j += 32) *reinterpret_cast<float4*>(dst + 4 * j) = src[j];
            }
          } else {
            for (int j = lane; j < D; j += 32) dst[j] = load_y(p, (row0 + b) * D + j);
          }
          for (int e = lane; e < Ep; e += 32) {
            float v = 0.f;
            if (e < u) v = p.u_in[(row0 + b) * u + e];
            else if (e < u + d) v = prior ? st[p.lay.prior_mean + e - u] : qm[(size_t)(b0 + b) * d + e - u];
            else if (e < E) v = prior ? st[p.lay.prior_logvar + e - u - d] : ql[(size_t)(b0 + b) * d + e - u - d];
            dst[D + e] = v;
          }
          if (p.eps) {
            const float* e0 = p.eps + ((size_t)t * 2 * p.B + b0 + b) * d;
            for (int k = lane; k < 2 * d; k += 32) eps_s[b * 2 * d + k] = (k < d) ? e0[k] : e0[(size_t)p.B * d + k - d];
          } else {
            const int nblk = (d + 3) >> 2;
            if (lane < 2 * nblk) {
              const int which = lane / nblk, blk = lane - which * nblk;
              float z[4];
              philox_normal4(p.seed, p.step0 + t, p.trial_offset + b0 + b, which, blk, z);
              for (int k = 0; k < 4; ++k)
                if (blk * 4 + k < d) eps_s[b * 2 * d + which * d + blk * 4 + k] = z[k];
            }
          }
        } else {
          // pad rows of everything that is summed over the rows of the tile
          for (int j = lane; j < K1p; j += 32) dst[j] = 0.f;
          if (!p.ext) for (int j = lane; j < Rp; j += 32) { phi_s[b * Rp + j] = 0.f; phil_s[b * Rp + j] = 0.f; }
          for (int j = lane; j < Gp; j += 32) { gpa[b * Gp + j] = 0.f; gpb[b * Gp + j] = 0.f; gpal[b * Gp + j] = 0.f; gpbl[b * Gp + j] = 0.f; }
          for (int j = lane; j < d; j += 32) { dx_s[b * d + j] = 0.f; gmt_s[b * d + j] = 0.f; glt_s[b * d + j] = 0.f; xt_s[b * d + j] = 0.f; }
        }
      }
    }
    VJF_STAMP(p, t, 22);
    cp_async_wait_all();  // shared parameters staged by the prologue (no-op after the first tile)
    __syncthreads();

    VJF_STAMP(p, t, 8);
    if (first && !p.ext && !(tma && cx->consts_staged)) {  // finish the staged RBF widths: -1/(2 w^2)
      for (int i = tid; i < R; i += VJF_NT) { const float w = expf(iw_s[i]); iw_s[i] = -0.5f / (w * w); }
    }
    if (tma) cx->consts_staged = 1;
    // ---- S1: xs = m_s + eps1 * exp(l_s / 2) (vjf/util.py:11-13); xu = [xs, u] (util.py:38-49) ----
    for (int i = tid; i < nb * du; i += VJF_NT) {
      const int b = i / du, k = i - b * du;
      float v;
      if (k < d) v = p.ext ? p.ext_xs[(size_t)(b0 + b) * d + k]  // large n_rbf (bigr.cu): drawn by the feature kernel
                           : in_s[b * K1p + D + u + k] + eps_s[b * 2 * d + k] * expf(0.5f * in_s[b * K1p + D + u + d + k]);
      else v = in_s[b * K1p + D + (k - d)];
      xu_s[b * du + k] = v;
    }
    __syncthreads();

    VJF_STAMP(p, t, 9);
    // ---- S2: RBF features phi = exp(-0.5 |xu - c|^2 / w^2) (vjf/functional.py:11-22), stored as a (hi, lo) pair;
    //      the staged input matrix is split in place the same way (its extras were consumed by S1) ----
    for (int b = warp; b < nb && !p.ext; b += VJF_NWARP) {
      for (int k = lane; k < Rp; k += 32) {
        float v = 0.f;
        if (k < R) {
          float d2 = 0.f;
          for (int c = 0; c < du; ++c) { const float df = xu_s[b * du + c] - c_s[k * du + c]; d2 = fmaf(df, df, d2); }
          v = expf(d2 * iw_s[k]);
        }
        const float h = tf32_hi(v);
        phi_s[b * Rp + k] = h;
        phil_s[b * Rp + k] = v - h;
      }
    }
    if (p.in_split) presplit_inplace(in_s, inl_s, rows * K1p);
    __syncthreads();

    VJF_STAMP(p, t, 11);
    // ---- S4: recognition MLP (vjf/recognition.py:31-42) on the tensor cores ----
    if (tma) {  // layer-1 weight, decoder and head weights: TMA bulk copies issued by the prologue
      mbar_wait(reinterpret_cast<uint64_t*>(sm + p.s_flag + 4), cx->w_phase);
      cx->w_phase ^= 1u;
    }
    {
      const float* A = in_s; int lda = K1p, K = K1;
      for (int l = 0; l < L; ++l) {
        float* out = sm + p.s_act[l];
        const bool w_sm = (l == 0) && p.W1_in_smem;
        mma_linear_fwd(A, (l == 0 && p.in_split) ? inl_s : nullptr, lda, K, w_sm ? (sm + p.s_W1) : (st + p.lay.mlp_w[l]), w_sm ? p.ldw1 : p.H[l], st + p.lay.mlp_b[l], p.H[l], out,
                       p.Hp[l], rows, true);
        // zero the remaining pad columns (beyond roundup(H,8)) read by the weight-gradient fragments
        const int h8 = (p.H[l] + 7) & ~7, h16 = (p.H[l] + 15) & ~15;
        if (h16 > h8)
          for (int i = tid; i < rows * (h16 - h8); i += VJF_NT) out[(i / (h16 - h8)) * p.Hp[l] + h8 + i % (h16 - h8)] = 0.f;
        __syncthreads();
        A = out; lda = p.Hp[l]; K = p.H[l];
      }
    }
    if (uc) {
      // tcgen05 path of the layer-1 weight gradient (back half): its B operand is the input matrix with the trials as the
      // K dimension, in the canonical K-major layout [(b/4)][k1][b%4] as a (hi, lo) pair.  The staged W1 is dead after S4,
      // so the pair is built in its place here, off the critical path.
      const int NK = p.umma_nk, nq = rows >> 2;
      float* bh = sm + p.s_W1;
      float* bl = bh + NK * rows;
      for (int i = tid; i < nq * NK; i += VJF_NT) {
        const int q = i / NK, k = i - q * NK;
        float h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = in_s[(4 * q + j) * K1p + k];
          if (p.in_split) { h[j] = x; l[j] = inl_s[(4 * q + j) * K1p + k]; }
          else { h[j] = tf32_hi(x); l[j] = x - h[j]; }
          // a spare (zero-pad) input column becomes a column of ones: its product with g_pre is the bias gradient
          if (k == NK - 1 && k >= K1) { h[j] = (4 * q + j < nb) ? 1.0f : 0.f; l[j] = 0.f; }
        }
        *reinterpret_cast<float4*>(bh + 4 * i) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(bl + 4 * i) = make_float4(l[0], l[1], l[2], l[3]);
      }
    }
    VJF_STAMP(p, t, 12);
    // heads: m_t = W_m h (no bias), l_t = W_v h + b_v ; then xt = m_t + eps2 exp(l_t/2), dx = xt - xs
    for (int b = warp; b < nb; b += VJF_NWARP) {
      float my_m = 0.f, my_lv = 0.f;
      for (int k = 0; k < d; ++k) {
        float m = 0.f, lv = 0.f;
        const float* wm = hm_s + k;
        const float* wv = hv_s + k;
        for (int n = lane; n < HL; n += 32) { const float h = hL[b * ldh + n]; m = fmaf(h, wm[n * d], m); lv = fmaf(h, wv[n * d], lv); }
        m = warp_sum(m); lv = warp_sum(lv);
        if (lane == k) { my_m = m; my_lv = lv + hv_s[HL * d + k]; }
      }
      if (lane < d) {
        const int i = b * d + lane;
        mt_s[i] = my_m; lt_s[i] = my_lv;
        const float x = my_m + eps_s[b * 2 * d + d + lane] * expf(0.5f * my_lv);
        xt_s[i] = x;
        const float dxv = x - xu_s[b * du + lane];
        dx_s[i] = dxv;
        sc[SC_SDX] = fmaf(dxv, dxv, sc[SC_SDX]);
        // posterior of this step -> trajectory (returned by filter / fit, model.py:218-221, :305-307)
        p.mu[((size_t)t * p.B + b0 + b) * d + lane] = my_m;
        p.logvar[((size_t)t * p.B + b0 + b) * d + lane] = my_lv;
      }
    }
    __syncthreads();

    VJF_STAMP(p, t, 13);
    // ---- S5/S6: decoder eta = D xt + bias (model.py:29-30), likelihood terms, dloss/deta (times B), g_xt = g_eta D,
    //      decoder gradients.  Specialised on the state dimension so that xt and the accumulators live in registers.
    // w_chol / w_mean of the previous step are usually published by the time the decoder stage is done (the RLS CTA is
    // ahead): their TMA is then issued before the Gram stage (below), so that the copies land behind it instead of in
    // front of the back half
    const bool early_try = tma && part == PART_FRONT && t > 0 && cx->early_ok && tid == 0;
    if (early_try) *reinterpret_cast<int*>(sm + p.s_flag + 10) = 0;
    // Gaussian likelihood: its logvar is updated by the RLS CTA (GaussianLikelihood.update) concurrently with this front
    // half in the overlapped schedule -- wait until the value of the previous step is final, then read it past L1
    float lam = 0.f;
    if (p.lik == VJF_LIK_GAUSSIAN) {
      if (cx_overlap && t > 0) wait_counter(p.ctrl + 4, (unsigned)t);
      lam = __ldcg(st + p.lay.lik_logvar);
    }
    const float e_nlam = expf(-lam), p_lam = expf(-0.5f * lam);
    switch (d) {
      case 1: decoder_stage<1>(p, sm, nb, first, r_on, lam, p_lam, e_nlam, dw, db, slot, sc); break;
      case 2: decoder_stage<2>(p, sm, nb, first, r_on, lam, p_lam, e_nlam, dw, db, slot, sc); break;
      case 3: decoder_stage<3>(p, sm, nb, first, r_on, lam, p_lam, e_nlam, dw, db, slot, sc, t); break;
      case 4: decoder_stage<4>(p, sm, nb, first, r_on, lam, p_lam, e_nlam, dw, db, slot, sc); break;
      default: if (d <= 8) decoder_stage<8>(p, sm, nb, first, r_on, lam, p_lam, e_nlam, dw, db, slot, sc);
               else decoder_stage<16>(p, sm, nb, first, r_on, lam, p_lam, e_nlam, dw, db, slot, sc);
    }

    VJF_STAMP(p, t, 18);
    if (early_try && ld_acquire_u32(p.ctrl + 5) >= (unsigned)t) {  // published: the copies land behind the Gram stage
      issue_back_tma(p, sm);
      *reinterpret_cast<int*>(sm + p.s_flag + 10) = 1;
    }
    // ---- S9: RLS sufficient statistics (vjf/module.py:94-96, unscaled): A += phi^T phi, b += phi^T dx ----
    if (p.ext) {  // large n_rbf: the statistics are two GEMMs over all trials (bigr.cu); hand over dx = xt - xs
      for (int i = tid; i < nb * d; i += VJF_NT) p.ext_dx[(size_t)b0 * d + i] = dx_s[i];
    } else {
    mma_gram(phi_s, phil_s, Rp, R, rows, slot + p.pa, first);
    {
      float* bp = slot + p.pb;
      for (int i = tid; i < R * d; i += VJF_NT) {
        const int r = i / d, k = i - r * d;
        float s = 0.f;
        for (int b = 0; b < nb; ++b) s = fmaf(phi_s[b * Rp + r] + phil_s[b * Rp + r], dx_s[b * d + k], s);
        acc_store(bp + i, s, first);
      }
    }
    }
    VJF_STAMP(p, t, 19);
    if (part == PART_FRONT) {
      // park the scalar sums of the front half until the back half finishes the tile
#pragma unroll
      for (int i = 0; i < VJF_NSCAL; ++i) {
        const float s = warp_sum(sc[i]);
        if (lane == 0) red_s[warp * VJF_NSCAL + i] = s;
      }
      __syncthreads();
      if (tid < VJF_NSCAL) {
        float s = 0.f;
        for (int w = 0; w < VJF_NWARP; ++w) s += red_s[w * VJF_NSCAL + tid];
        scf_s[tid] = s;
      }
      __syncthreads();
      return;
    }
    __syncthreads();
  }

  // =============================== back half ===============================
  if (part == PART_BACK) {
    if (tma) {  // w_chol / w_mean: TMA bulk copies of the back prologue
      mbar_wait(reinterpret_cast<uint64_t*>(sm + p.s_flag + 6), cx->y_phase);
      cx->y_phase ^= 1u;
    } else {
      cp_async_wait_all();  // w_chol / w_mean staged by the back prologue
      __syncthreads();
    }
  }
  VJF_STAMP(p, t, 10);
  if (first && p.U_in_smem && t == 0) {  // is w_chol upper triangular?  Checked once per launch: every later w_chol is
                                           // the L^-T this kernel computed itself (or unchanged)
    const float* U_s = sm + p.s_U;
    int nz = 0;
    for (int r = warp; r < R; r += VJF_NWARP)
      for (int c = lane; c < r; c += 32) nz |= (U_s[r * p.ldu + c] != 0.f);
    nz = __syncthreads_or(nz);
    if (tid == 0) *flag_s = nz;
    __syncthreads();
  }
  // ---- S3: dynamics read-out (vjf/module.py:75-77, model.py:338): p_mean = xs + phi W ;
  //      p_logvar = log |phi w_chol|^2  (the diagonal of the reference's (B,B) product) ----
  if (p.ext) {
    // large n_rbf: p_mean and p_logvar come from the feature kernel and the phi w_chol GEMM
  } else if (p.U_in_smem) {
    mma_quadform(phi_s, phil_s, Rp, sm + p.s_U, p.ldu, R, rows, qp_s, *flag_s == 0);
  } else {
    const float* U = st + p.lay.w_chol;
    for (int b = warp; b < nb; b += VJF_NWARP) {
      float q = 0.f;
      for (int k = lane; k < R; k += 32) {
        float fl = 0.f;
        const float* ph = phi_s + b * Rp;
        const float* pl = phil_s + b * Rp;
        for (int j = 0; j < R; ++j) fl = fmaf(ph[j] + pl[j], U[j * R + k], fl);
        q = fmaf(fl, fl, q);
      }
      q = warp_sum(q);
      if (lane == 0) qp_s[b] = q;
    }
  }
  VJF_STAMP(p, t, 49);
  if (p.ext) {
    for (int i = tid; i < nb * d; i += VJF_NT) pm_s[i] = p.ext_pm[(size_t)b0 * d + i];
  } else
  for (int i = tid; i < nb * d; i += VJF_NT) {  // one thread per (trial, state dim): four independent chains of length R/4
    const int b = i / d, k = i - b * d;
    const float* ph = phi_s + b * Rp;
    const float* pl = phil_s + b * Rp;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = 0;
    for (; r + 3 < R; r += 4) {
      s0 = fmaf(ph[r] + pl[r], W_s[r * d + k], s0);
      s1 = fmaf(ph[r + 1] + pl[r + 1], W_s[(r + 1) * d + k], s1);
      s2 = fmaf(ph[r + 2] + pl[r + 2], W_s[(r + 2) * d + k], s2);
      s3 = fmaf(ph[r + 3] + pl[r + 3], W_s[(r + 3) * d + k], s3);
    }
    for (; r < R; ++r) s0 = fmaf(ph[r] + pl[r], W_s[r * d + k], s0);
    pm_s[i] = xu_s[b * du + k] + ((s0 + s1) + (s2 + s3));
  }
  VJF_STAMP(p, t, 50);
  __syncthreads();
  VJF_STAMP(p, t, 51);
  // p_logvar from the n-tile partial sums
  if (tid < nb) {
    if (p.ext) plv_s[tid] = p.ext_plv[b0 + tid];
    else {
    float q = qp_s[tid];
    if (p.U_in_smem) { const int nt = (R + 7) >> 3; for (int n = 1; n < nt; ++n) q += qp_s[n * rows + tid]; }
    plv_s[tid] = logf(q);
    }
  }
  __syncthreads();

  VJF_STAMP(p, t, 15);
  // state-noise logvar of the previous step.  Overlapped schedule: there is no grid barrier at the end of a step -- the RLS
  // CTA publishes "step t-1 complete" (ctrl[3]) after its tail, and this is the first place that needs it
  if (cx_overlap && t > 0) wait_counter(p.ctrl + 3, (unsigned)t);
  const float gam = __ldcg(st + p.lay.tr_logvar);
  const float e_ngam = expf(-gam), p_gam = expf(-0.5f * gam);
  // ---- S7: dynamics NLL (functional.py:55-75 via model.py:390-391), entropy (functional.py:25-29),
  //      g_mt and g_lt (times B) ----
  for (int i = tid; i < nb * d; i += VJF_NT) {
    const int b = i / d, k = i - b * d;
    const float m = mt_s[i], lv = lt_s[i], pm = pm_s[i], plv = plv_s[b];
    const float e2 = eps_s[b * 2 * d + d + k], gx = gxt_s[i];
    const float df = pm * p_gam - m * p_gam;
    const float mse = df * df;
    if (!isfinite(mse)) sc[SC_BADMSE] += 1.f;
    const float tr = expf(plv + lv - gam);
    sc[SC_DYN] += 0.5f * (mse + gam) + 0.5f * tr;
    sc[SC_ENT] += 0.5f * lv;
    float gm = gx, gl = 0.5f * gx * e2 * expf(0.5f * lv);
    if (h_on) gl -= 0.5f;
    if (d_on) { gm += (m - pm) * e_ngam; gl += 0.5f * tr; }
    gmt_s[i] = gm; glt_s[i] = gl;
  }
  __syncthreads();

  VJF_STAMP(p, t, 16);
  // ---- S8: backward through the heads and the MLP ----
  {
    // head weight gradients [H_L][d] (input-major) and logvar-head bias
    for (int i = tid; i < HL * d; i += VJF_NT) {
      const int n = i / d, k = i - n * d;
      float am = 0.f, av = 0.f;
      for (int b = 0; b < nb; ++b) { const float h = hL[b * ldh + n]; am = fmaf(h, gmt_s[b * d + k], am); av = fmaf(h, glt_s[b * d + k], av); }
      acc_store(slot + p.lay.head_m_w + i, am, first);
      acc_store(slot + p.lay.head_v_w + i, av, first);
    }
    if (tid >= VJF_NT - 32 && lane < d) {
      float s = 0.f;
      for (int b = 0; b < nb; ++b) s += glt_s[b * d + lane];
      acc_store(slot + p.lay.head_v_b + lane, s, first);
    }
    // g_pre of the last hidden layer: (g_mt W_m + g_lt W_v) * (1 - h^2); pad columns up to a multiple of 8 are zero
    const int H8 = (HL + 7) & ~7;
    if (uc) {
      // tcgen05 path: G^T [64 x rows] in the canonical K-major layout (K = trials), pads zero
      for (int b = warp; b < rows; b += VJF_NWARP) {
        for (int n = lane; n < 64; n += 32) {
          float v = 0.f;
          if (n < HL && b < nb) {
            float s = 0.f;
            for (int k = 0; k < d; ++k) { s = fmaf(gmt_s[b * d + k], hm_s[n * d + k], s); s = fmaf(glt_s[b * d + k], hv_s[n * d + k], s); }
            const float h = hL[b * ldh + n];
            v = s * (1.0f - h * h);
          }
          const float vh = tf32_hi(v);
          const int o = umma_canon(n, b, 64);
          gpa[o] = vh;
          gpal[o] = v - vh;
        }
      }
    } else
    for (int b = warp; b < nb; b += VJF_NWARP) {
      for (int n = lane; n < H8; n += 32) {
        float v = 0.f;
        if (n < HL) {
          float s = 0.f;
          for (int k = 0; k < d; ++k) { s = fmaf(gmt_s[b * d + k], hm_s[n * d + k], s); s = fmaf(glt_s[b * d + k], hv_s[n * d + k], s); }
          const float h = hL[b * ldh + n];
          v = s * (1.0f - h * h);
        }
        const float vh = tf32_hi(v);
        gpa[b * Gp + n] = vh;
        gpal[b * Gp + n] = v - vh;
      }
    }
    __syncthreads();
    VJF_STAMP(p, t, 17);
    float* gcur = gpa; float* gnext = gpb; float* gcurl = gpal; float* gnextl = gpbl;
    if (uc) {
      // bias gradient = column sums of g_pre, then the weight gradient on tcgen05 / TMEM
      const bool bias_col = p.umma_nk - 1 >= K1;  // the ones column exists: the bias gradient comes out of the MMA
      if (!bias_col && tid < HL) {
        float s = 0.f;
        for (int b = 0; b < nb; ++b) { const int o = umma_canon(tid, b, 64); s += gpa[o] + gpal[o]; }
        acc_store(slot + p.lay.mlp_b[0] + tid, s, first);
      }
      float* bh = sm + p.s_W1;
      umma_wgrad(gpa, gpal, bh, bh + p.umma_nk * rows, p.umma_nk, K1, HL, rows, uc->tmem, 0u,
                 reinterpret_cast<uint64_t*>(sm + p.s_flag + 2), uc->umma_phase, slot + p.lay.mlp_w[0],
                 bias_col ? slot + p.lay.mlp_b[0] : nullptr, bh, first, &p, t);
      uc->umma_phase ^= 1u;
      VJF_STAMP(p, t, 43);
    } else
    for (int l = L - 1; l >= 0; --l) {
      const float* Aprev = (l == 0) ? in_s : (sm + p.s_act[l - 1]);
      const int ldp = (l == 0) ? K1p : p.Hp[l - 1];
      const int Kl = (l == 0) ? K1 : p.H[l - 1];
      mma_wgrad(Aprev, (l == 0 && p.in_split) ? inl_s : nullptr, ldp, Kl, gcur, gcurl, Gp, p.H[l], rows, slot + p.lay.mlp_w[l], first);
      tile_colsum(gcur, gcurl, Gp, p.H[l], nb, slot + p.lay.mlp_b[l], first);
      if (l > 0) {
        const float* Wl = st + p.lay.mlp_w[l];  // [Kl][H_l]
        const int N = p.H[l], K8 = (Kl + 7) & ~7;
        for (int b = warp; b < nb; b += VJF_NWARP) {
          for (int k = lane; k < K8; k += 32) {
            float v = 0.f;
            if (k < Kl) {
              float s = 0.f;
              for (int n = 0; n < N; ++n) s = fmaf(gcur[b * Gp + n] + gcurl[b * Gp + n], Wl[k * N + n], s);
              const float h = Aprev[b * ldp + k];
              v = s * (1.0f - h * h);
            }
            const float vh = tf32_hi(v);
            gnext[b * Gp + k] = vh;
            gnextl[b * Gp + k] = v - vh;
          }
        }
        __syncthreads();
        float* tmp = gcur; gcur = gnext; gnext = tmp;
        tmp = gcurl; gcurl = gnextl; gnextl = tmp;
      }
    }
  }

  // ---- scalar sums of the tile ----
#pragma unroll
  for (int i = 0; i < VJF_NSCAL; ++i) {
    const float s = warp_sum(sc[i]);
    if (lane == 0) red_s[warp * VJF_NSCAL + i] = s;
  }
  __syncthreads();
  VJF_STAMP(p, t, 44);
  if (tid < VJF_NSCAL) {
    float s = (part == PART_BACK) ? scf_s[tid] : 0.f;
    for (int w = 0; w < VJF_NWARP; ++w) s += red_s[w * VJF_NSCAL + tid];
    if (tid == 6) acc_store(slot + p.lay.lik_logvar, s, first);  // Gaussian d loss / d lambda (times B)
    else acc_store(slot + p.ps + tid, s, first);
    // a loss-term partial that is not comfortably finite: tell the grid (through barrier 1) to take the exact check of the sums
    if (tid < 3 && !(fabsf(s) < 1e30f)) *reinterpret_cast<volatile int*>(sm + p.s_flag + 8) = 1;  // handed to barrier 1 (grid_barrier_flag)
  }
  __syncthreads();
  VJF_STAMP(p, t, 45);
  if (tma && part == PART_BACK && t + 1 < p.T && p.y_dtype != VJF_Y_U8 && (D & 3) == 0) {
    // prefetch the observations of step t+1 into the (now dead) input matrix; the copies land while the grid reduces
    const float* src = reinterpret_cast<const float*>(p.y) + ((size_t)(t + 1) * p.B + b0) * D;
    const int c4 = D >> 2;
    for (int i = tid; i < nb * c4; i += VJF_NT) { const int b = i / c4, c = (i - b * c4) << 2; cp_async16(in_s + b * K1p + c, src + (size_t)b * D + c); }
    cp_async_commit();
    cx->y_ready_t = t + 1;
  }
  VJF_STAMP(p, t, 20);
}

// Asynchronous staging (cp.async) of the parameters the tiles of a step share.  STAGE_FRONT: what the front
// half reads (RBF centres/widths, recognition layer-1 weight, head weights, decoder) -- final once the SGD
// step of the previous time step is done.  STAGE_BACK: w_mean / w_chol -- final once its RLS is done.
#define STAGE_FRONT 1
#define STAGE_BACK 2
static __device__ void phase_a_prologue(const StepParams& p, float* sm, int what, TileCtx* cx = nullptr) {
  const int tid = threadIdx.x;
  const bool tma = cx && (cx->flags & CX_TMA);
  const float* st = p.state;
  const int HL = p.H[p.L - 1];
  if (what & STAGE_FRONT) {
    // RBF centres and widths are not trained (vjf/module.py:115-130: requires_grad=False) and the RLS does not touch
    // them: with a dedicated tile per CTA they are staged once per launch
    if (!p.ext && (!tma || !cx->consts_staged)) {
      stage_async(sm + p.s_c, p.du, st + p.lay.centroid, p.du, p.R, p.du, tid, VJF_NT);
      stage_async(sm + p.s_iw, p.R, st + p.lay.logwidth, p.R, 1, p.R, tid, VJF_NT);
    }
    if (!tma) {
      stage_async(sm + p.s_hm, p.d, st + p.lay.head_m_w, p.d, HL, p.d, tid, VJF_NT);
      stage_async(sm + p.s_hv, p.d, st + p.lay.head_v_w, p.d, HL, p.d, tid, VJF_NT);
      stage_async(sm + p.s_hv + HL * p.d, p.d, st + p.lay.head_v_b, p.d, 1, p.d, tid, VJF_NT);
    }
    if (tma) {
      // the two big ones by TMA bulk copies issued by one thread, completion on an mbarrier that the tile waits on right
      // before S4 -- S0..S2 run while the data is in flight.  The layer-1 weight comes from its row-padded mirror
      // (StepParams::w1_mirror, kept up to date by the SGD step), so that ONE copy lands it in the bank-conflict-free
      // layout the MMA fragments want.  The sources were written by other CTAs through the generic proxy (ordered by
      // the grid barrier): fence.proxy.async orders the async-proxy reads after them.
      const uint32_t bytes = tma_front_bytes(p);
      if (tid == 0) {
        uint64_t* bar = reinterpret_cast<uint64_t*>(sm + p.s_flag + 4);
        VJF_STAMP(p, 0, 47);
        asm volatile("fence.proxy.async.global;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        VJF_STAMP(p, 0, 48);
        mbar_expect_tx(bar, bytes);
        if (p.W1_in_smem) tma_bulk_g2s(sm + p.s_W1, p.w1_mirror, (uint32_t)p.K1 * p.ldw1 * 4u, bar);
        if (p.dec_in_smem) {
          tma_bulk_g2s(sm + p.s_dec, st + p.lay.dec_w, (uint32_t)p.d * p.D * 4u, bar);
          tma_bulk_g2s(sm + p.s_dec + p.d * p.D, st + p.lay.dec_b, (uint32_t)p.D * 4u, bar);
        }
        // head weights [H_L][d] and the logvar-head bias (sizes rounded up to 16 bytes: the state segments and the
        // shared arrays are padded)
        const uint32_t hb = tma_head_bytes(p);
        tma_bulk_g2s(sm + p.s_hm, st + p.lay.head_m_w, hb, bar);
        tma_bulk_g2s(sm + p.s_hv, st + p.lay.head_v_w, hb, bar);
        tma_bulk_g2s(sm + p.s_hv + HL * p.d, st + p.lay.head_v_b, 16u * ((p.d + 3) >> 2), bar);
      }
    } else {
      if (p.dec_in_smem) {
        stage_async(sm + p.s_dec, p.D, st + p.lay.dec_w, p.D, p.d, p.D, tid, VJF_NT);
        stage_async(sm + p.s_dec + p.d * p.D, p.D, st + p.lay.dec_b, p.D, 1, p.D, tid, VJF_NT);
      }
      if (p.W1_in_smem) stage_async(sm + p.s_W1, p.ldw1, st + p.lay.mlp_w[0], p.H[0], p.K1, p.H[0], tid, VJF_NT);
    }
  }
  if ((what & STAGE_BACK) && tma) {
    // w_chol from its row-padded mirror (kept current by the RLS commit; pads are zero) and w_mean: two TMA bulk copies,
    // completion on the mbarrier the back half waits on
    if (tid == 0) issue_back_tma(p, sm);
  } else if ((what & STAGE_BACK) && !p.ext) {
    stage_async(sm + p.s_W, p.d, st + p.lay.w_mean, p.d, p.R, p.d, tid, VJF_NT);
    if (p.U_in_smem) {
      float* U_s = sm + p.s_U;
      const int Rk = (p.R + 7) & ~7, ldu = p.ldu;
      stage_async(U_s, ldu, st + p.lay.w_chol, p.R, p.R, p.R, tid, VJF_NT);
      // zero padding (disjoint from the async destinations)
      for (int i = tid; i < Rk * ldu; i += VJF_NT) { const int r = i / ldu, c = i - r * ldu; if (r >= p.R || c >= p.R) U_s[i] = 0.f; }
    }
  }
  cp_async_commit();
}

// whole phase A for the non-overlapped schedule (also the split multi-GPU path): every CTA walks its tiles
static __device__ void phase_a(const StepParams& p, float* sm, int t, unsigned masks) {
  VJF_STAMP(p, t, 7);
  __syncthreads();  // the previous phase is done with the shared memory that is re-planned here
  phase_a_prologue(p, sm, STAGE_FRONT | STAGE_BACK);
  VJF_STAMP(p, t, 21);
  bool first = true;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    phase_a_tile(p, sm, t, tile, first, masks, PART_BOTH);
    first = false;
  }
  if (first) {  // a CTA without tiles still owns a slot: zero it
    cp_async_wait_all();
    float* slot = p.partials + (size_t)blockIdx.x * p.PS;
    for (int i = threadIdx.x; i < p.PS; i += VJF_NT) slot[i] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// phase B1: reduce the slots (two-level, fixed order), clip + SGD on the owned slice
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sgd_applies(const StepParams& p, int e) {
  if (!(p.flags & VJF_FLAG_SGD)) return false;
  if (e == p.lay.lik_logvar) return p.lik == VJF_LIK_GAUSSIAN;
  if ((p.flags & VJF_FLAG_DECODER_FROZEN) && e >= p.lay.dec_w && e < p.lay.dec_b + p.D) return false;
  return e < p.lay.n_train;
}

// reduce `src` ([nslots][PS]) into p.reduced; when apply is set also take the SGD step (vjf/model.py:210-211)
// Chunks of 128 consecutive elements (32 float4 columns).  A CTA pass reduces one chunk: warp w sums slots
// w, w+16, ... with 128-bit loads (all of a warp's loads are in flight together), the 16 partial sums are
// combined in a fixed order through shared memory, and warp 0 writes the result and takes the SGD step.
// Chunks are walked from the END of the vector, so the RLS statistics and the loss sums (which live at the
// end) are reduced first; when `signal` is given, every finished statistics chunk bumps it so that the RLS CTA
// can start the factorisation without waiting for the gradient part.
static __device__ void phase_b1(const StepParams& p, float* sm, const float* src, int nslots, bool apply, int cta, int nctas,
                                unsigned* signal = nullptr, unsigned epoch = 0, unsigned my_fin = 7u, unsigned need = 0u) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float invB = 1.0f / (float)p.Bglobal;
  float4* red = reinterpret_cast<float4*>(sm + p.s_b1);  // [16][32] float4
  const int c_lo = p.red_begin >> 7, nchunks = (p.PS + 127) >> 7;
  const int stat_chunk0 = p.pa >> 7;  // chunks >= this one hold RLS statistics / loss sums
  // Rounds of nctas chunks from the end of the vector; odd rounds are dealt out in the opposite CTA order, so that the few
  // chunks beyond a whole number of rounds land on the CTAs that hold gradient chunks, not on those that hold the statistics
  // chunks the RLS waits for.  A CTA that owns a chunk in two consecutive rounds reduces them side by side, eight warps
  // each, instead of one after the other: the phase lasts one chunk time for every CTA.
  const int top = nchunks - 1;
  for (int k = 0; top - k * nctas >= c_lo; k += 2) {
    const int chA = top - k * nctas - cta, chB = top - (k + 2) * nctas + 1 + cta;  // round k forwards, round k + 1 backwards
    const bool vA = chA >= c_lo, vB = chB >= c_lo;
    if (!vA && !vB) continue;
    const bool two = vA && vB;
    const int wpg = two ? VJF_NWARP / 2 : VJF_NWARP;   // warps per chunk
    const int grp = warp / wpg, wi = warp - grp * wpg;
    const int ch = two ? (grp ? chB : chA) : (vA ? chA : chB);
    const int e0 = (ch << 7) + (lane << 2);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e0 < p.PS) {
      const float* q = src + e0;
      if (!two) {
        // up to 10 slots per warp (nslots <= 160): issue every load before the first add, one L2 round trip in all
        float4 v[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const int c = warp + j * VJF_NWARP;
          v[j] = (c < nslots) ? *reinterpret_cast<const float4*>(q + (size_t)c * p.PS) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 10; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
        for (int c = warp + 10 * VJF_NWARP; c < nslots; c += VJF_NWARP) {
          const float4 v0 = *reinterpret_cast<const float4*>(q + (size_t)c * p.PS);
          acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
        }
      } else {
        // eight warps per chunk: two rounds of ten slots each (nslots <= 160)
#pragma unroll 1
        for (int base = wi; base < nslots; base += 10 * (VJF_NWARP / 2)) {
          float4 v[10];
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            const int c = base + j * (VJF_NWARP / 2);
            v[j] = (c < nslots) ? *reinterpret_cast<const float4*>(q + (size_t)c * p.PS) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 10; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
        }
      }
    }
    red[warp * 32 + lane] = acc;
    __syncthreads();
    if (wi == 0) {
      float4 t = red[warp * 32 + lane];
      for (int w = 1; w < wpg; ++w) { const float4 v = red[(warp + w) * 32 + lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
      if (epoch) {
        // ---- in-kernel all-reduce over NVLink peer memory: push this rank's chunk into every rank's inbox, raise the
        //      (source rank, chunk) flag there, wait for the same chunk of every rank, add in rank order ----
        const int par = epoch & 1, nchx = p.PSx >> 7;
        const size_t flag_off = (size_t)p.world * 2 * p.PSx;
        if (e0 < p.PS)
          for (int r = 0; r < p.world; ++r)
            *reinterpret_cast<float4*>(p.peer[r] + (size_t)(p.rank * 2 + par) * p.PSx + e0) = t;
        // the warp barrier orders every lane's stores before the flag lanes' st.release.sys (cumulativity): no separate
        // system-scope fence is needed, and it would cost a second NVLink round trip
        __syncwarp();
        // The flag carries (epoch, finite mask of this rank's three ELBO sums): every CTA learns the global decision
        // "a term is non-finite somewhere" from the flags it polls anyway, identically on every rank, before it applies
        // the SGD step of its chunk (vjf/model.py:138-145, :212-214: such a step is skipped, as in the split path).
        unsigned got = 7u;
        if (lane < p.world) {
          st_release_sys_u32(reinterpret_cast<unsigned*>(p.peer[lane] + flag_off) + p.rank * nchx + ch, (epoch << 3) | (my_fin & 7u));
          const unsigned* wf = reinterpret_cast<const unsigned*>(p.peer[p.rank] + flag_off) + lane * nchx + ch;
          const long long t0 = clock64();
          unsigned v, spins = 0;
          while (((v = ld_acquire_sys_u32(wf)) >> 3) < epoch) {
            // a lost peer: give up after ~3 s, and at once when another CTA already gave up (sticky status bit)
            if (clock64() - t0 > 6000000000ll) { atomicOr(p.status, (unsigned)VJF_ST_COMM_TIMEOUT); break; }
            if ((++spins & 1023u) == 0 && (*reinterpret_cast<volatile unsigned*>(p.status) & VJF_ST_COMM_TIMEOUT)) break;
          }
          got = ((v >> 3) >= epoch) ? (v & 7u) : 0u;
        }
        got = __reduce_and_sync(0xffffffffu, got);
        // after a time-out nothing is applied any more: the replicas would diverge silently (the host raises)
        if ((got & need) != need || (*reinterpret_cast<volatile unsigned*>(p.status) & VJF_ST_COMM_TIMEOUT)) apply = false;
        __syncwarp();  // the polling lanes' ld.acquire.sys + this barrier order the inbox reads below after the peers' data
        if (e0 < p.PS) {
          t = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int r = 0; r < p.world; ++r) {
            const float4 v = ld_volatile_f4(p.peer[p.rank] + (size_t)(r * 2 + par) * p.PSx + e0);
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
          }
        }
      }
      if (e0 < p.PS) {
      *reinterpret_cast<float4*>(p.reduced + e0) = t;
      if (apply && e0 < p.lay.n_train) {
        const float tv[4] = {t.x, t.y, t.z, t.w};
        // the row-padded mirror of the layer-1 weight (TMA source) is kept current; e0 and H are multiples of 4, so the
        // four elements of a lane share a row
        const int w0 = e0 - p.lay.mlp_w[0];
        float* mir = (p.use_tma && w0 >= 0 && w0 < p.K1 * p.H[0]) ? p.w1_mirror + (w0 / p.H[0]) * p.ldw1 + (w0 % p.H[0]) : nullptr;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (sgd_applies(p, e0 + i)) {
            const float v = p.state[e0 + i] - p.lr * clip1(tv[i] * invB);
            p.state[e0 + i] = v;
            if (mir) mir[i] = v;
            if (p.tp.on) {  // tile pipeline: the layer-1 weight and bias live on as tcgen05 operand images
              const int wi = e0 + i - p.lay.mlp_w[0], bi = e0 + i - p.lay.mlp_b[0];
              if (wi >= 0 && wi < p.K1 * p.H[0]) w1k_store(p, wi / p.H[0], wi % p.H[0], v);
              else if (bi >= 0 && bi < p.H[0]) w1k_store(p, p.K1, bi, v);
            }
          }
      }
      }
    }
    __syncthreads();
    if (signal && tid == 0) {
      const unsigned n = ((vA && chA >= stat_chunk0) ? 1u : 0u) + ((vB && chB >= stat_chunk0) ? 1u : 0u);
      if (n) { __threadfence(); atomicAdd(signal, n); }
    }
  }
}

// SGD from an already reduced vector (multi-GPU split path, after the all-reduce)
static __device__ void sgd_from_reduced(const StepParams& p, int cta, int nctas) {
  const float invB = 1.0f / (float)p.Bglobal;
  for (int e = cta * VJF_NT + threadIdx.x; e < p.lay.n_train; e += nctas * VJF_NT)
    if (sgd_applies(p, e)) p.state[e] -= p.lr * clip1(p.reduced[e] * invB);
}

// ------------------------------------------------------------------------------------------
// phase B2 (one CTA): losses, running variances, RLS
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float running_var_f(float acc_var, float acc_n, float new_var, float new_n, float cap,
                                               float* n_out) {
  // vjf/util.py:20-35 ; f1, f2 formed in double like the reference's Python floats, applied in fp32
  const double a = fmin((double)acc_n, (double)cap), tot = a + (double)new_n;
  const float f1 = (float)(a / tot), f2 = (float)((double)new_n / tot);
  *n_out = (float)tot;
  return f1 * acc_var + f2 * new_var;
}

static __device__ double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < VJF_NWARP; ++w) s += red[w];
  __syncthreads();
  return s;
}

// ---- LinearRegression.rls (vjf/module.py:89-102), register-resident ------------------------------------
// Work matrix rows: [0,R) lower triangle of P' = P + A/v ; [R,R+d) g^T with g = P W + b/v ; [R+d,2R+d) I.
// A right-looking LDL^T sweep over the R columns of P' applies the same eliminations to the appended
// rows, i.e. performs the forward substitutions L^-1 g and L^-1 I on the fly; after scaling column k by
// 1/sqrt(pivot_k) the three row groups hold chol(P'), (L^-1 g)^T and L^-T = w_chol.  Row r lives in warp
// r % 16 (register slot r / 16), column j in lane j % 32 (slot j / 32): the sweep touches shared memory only
// to broadcast the current column.  Returns false (block-uniform) if a pivot is not positive.
#ifdef VJF_DEBUG_STAMPS
__device__ long long g_sweep_ticks[160];  // development aid: clock64 at every column of the last sweep
#endif

// One range [k0, k1) of the LDL^T sweep with compile-time register slots: CN = slot of column k, CNP = slot of
// column k + 1.  Multipliers of a warp's own rows come from a warp shuffle (element (r, k) lives in lane k % 32 of
// the warp that owns row r); only the P'-rows' entries of the next column travel through shared memory.
template <int CPL, int RPW, int PR, int NW, int CN, int CNP>
__device__ __forceinline__ bool ldl_sweep_range(float (&v)[RPW][CPL], const int (&jj)[CPL], float* colbuf, float* dvec, int k0, int k1,
                                                int R, int& kb) {
  constexpr int NRP = NW * RPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = k0; k < k1; ++k) {
    const float* cb = colbuf + kb * NRP;
    float* cbn = colbuf + (kb ^ 1) * NRP;
#ifdef VJF_DEBUG_STAMPS
    if (threadIdx.x == 0) g_sweep_ticks[k] = clock64();
#endif
    const float piv = cb[k];
    float tk[RPW], cj[CPL];
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) { const float x = cb[lane + 32 * ci]; cj[ci] = (jj[ci] > k) ? x : 0.f; }
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) tk[ri] = __shfl_sync(0xffffffffu, v[ri][CN], k & 31);
    if (!(piv > 0.f)) return false;
    float rinv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(piv));
    rinv = fmaf(rinv, fmaf(-piv, rinv, 1.0f), rinv);  // one Newton step: < 1 ulp
    if (threadIdx.x == 0) dvec[k] = piv;
    const float ninv = -rinv;
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) tk[ri] *= ninv;
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) {
#pragma unroll
      for (int ci = 0; ci < CPL; ++ci) v[ri][ci] = fmaf(tk[ri], cj[ci], v[ri][ci]);
    }
    if (lane == ((k + 1) & 31)) {
#pragma unroll
      for (int ri = 0; ri < PR; ++ri) cbn[warp + NW * ri] = v[ri][CNP];  // rows >= R in these slots are never read
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");  // only the sweep warps
    kb ^= 1;
  }
  return true;
}

template <int CPL, int RPW, int NW>
static __device__ bool rls_factor_regs(const StepParams& p, float* sm, float iv, const float* A, const float* bv, int t,
                                       double* resid_out, const unsigned* wait_ctr = nullptr, unsigned wait_val = 0, bool* noise_done = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, d = p.d, NR = 2 * R + d;
  // NW warps take part (row r lives in warp r % NW)
  constexpr int NRP = NW * RPW;                 // rows incl. padding: every (warp, slot) pair is a row
  constexpr int PR = (RPW + 1) / 2;             // register slots that can hold P rows (R <= NR / 2)
  const bool active_warp = warp < NW;
  const int ldq = R | 1;                        // odd row stride: conflict-free column walks
  float* colbuf = sm;                           // [2][NRP]
  float* dvec = colbuf + 2 * NRP;               // [R]
  float* zbuf = dvec + ((R + 3) & ~3);          // [d][R]   g, later z = L^-1 g
  float* Ws = zbuf + ((d * R + 3) & ~3);        // [R][d]   old W, later W'
  float* bs = Ws + ((d * R + 3) & ~3);          // [R][d]   b staged
  float* misc = bs + ((d * R + 3) & ~3);        // [0] fail flag
  double* dred = reinterpret_cast<double*>(misc + 4);  // [NWARP] (8-byte aligned: all sizes above are multiples of 4 floats)
  float* Qs = misc + 4 + 2 * VJF_NWARP + 4;     // [R][ldq]  P (for g = P W), later w_chol (for W' = w_chol z)
  float* st = p.state;
  const float* P = st + p.lay.w_precision;
  float v[RPW][CPL];
  float v0[PR][CPL];                            // P' rows kept for the commit
  float a_in[PR][CPL];                          // lower triangle of A (kept for the residual at the end)
  // ---- part 1, independent of this step's statistics (runs while the RLS CTA would otherwise wait for them): old P into
  //      registers and shared memory, old W, and P W ----
#pragma unroll
  for (int ri = 0; ri < PR; ++ri) {
    const int r = active_warp ? warp + NW * ri : (1 << 20);
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) {
      const int j = lane + 32 * ci;
      v[ri][ci] = ((r < R) && (j < R)) ? P[r * R + j] : 0.f;
    }
  }
  for (int i = tid; i < R * d; i += VJF_NT) Ws[i] = st[p.lay.w_mean + i];
#pragma unroll
  for (int ri = 0; ri < PR; ++ri) {
    const int r = active_warp ? warp + NW * ri : (1 << 20);
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) {
      const int j = lane + 32 * ci;
      if (r < R && j < R) Qs[r * ldq + j] = v[ri][ci];
    }
  }
#pragma unroll
  for (int ri = PR; ri < RPW; ++ri)
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) v[ri][ci] = 0.f;
  if (tid == 0) misc[0] = 0.f;
  __syncthreads();
  // (P W)[r][c] (old P, old W: module.py:93); one thread per output, parked in zbuf across the wait
  for (int i = tid; i < R * d; i += VJF_NT) {
    const int r = i / d, c = i - r * d;
    float s0 = 0.f, s1 = 0.f;
    const float* q = Qs + r * ldq;
    int j = 0;
    for (; j + 1 < R; j += 2) { s0 = fmaf(q[j], Ws[j * d + c], s0); s1 = fmaf(q[j + 1], Ws[(j + 1) * d + c], s1); }
    if (j < R) s0 = fmaf(q[j], Ws[j * d + c], s0);
    zbuf[c * R + r] = s0 + s1;
  }
  // ---- part 2: this step's statistics.  P' = P + A/v on and below the diagonal ; g = P W + b/v ----
  if (wait_ctr) wait_counter(wait_ctr, wait_val);
  VJF_STAMP(p, t, 30);
#pragma unroll
  for (int ri = 0; ri < PR; ++ri) {
    const int r = active_warp ? warp + NW * ri : (1 << 20);
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) {
      const int j = lane + 32 * ci;
      a_in[ri][ci] = ((r < R) && (j < R) && j <= r) ? __ldcg(A + r * R + j) : 0.f;
    }
  }
  for (int i = tid; i < R * d; i += VJF_NT) {  // same thread <-> element mapping as above: no barrier needed in between
    const int r = i / d, c = i - r * d;
    const float b = __ldcg(bv + i);
    bs[i] = b;
    zbuf[c * R + r] = fmaf(b, iv, zbuf[c * R + r]);
  }
#pragma unroll
  for (int ri = 0; ri < PR; ++ri)
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) {
      v[ri][ci] = fmaf(a_in[ri][ci], iv, v[ri][ci]);
      v0[ri][ci] = v[ri][ci];
    }
  VJF_STAMP(p, t, 25);
  __syncthreads();
#pragma unroll
  for (int ri = 0; ri < RPW; ++ri) {
    const int r = active_warp ? warp + NW * ri : (1 << 20);
    if (r >= R) {
#pragma unroll
      for (int ci = 0; ci < CPL; ++ci) {
        const int j = lane + 32 * ci;
        float x = 0.f;
        if (r < R + d) x = (j < R) ? zbuf[(r - R) * R + j] : 0.f;
        else if (r < NR) x = (j == r - R - d) ? 1.0f : 0.f;
        v[ri][ci] = x;
      }
    }
  }
  // publish column 0
  if (lane == 0 && active_warp) {
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) colbuf[warp + NW * ri] = v[ri][0];
  }
  int jj[CPL];  // column index of each slot, -1 when it is not a column of P'
#pragma unroll
  for (int ci = 0; ci < CPL; ++ci) { const int j = lane + 32 * ci; jj[ci] = (j < R) ? j : -1; }
  __syncthreads();
  VJF_STAMP(p, t, 26);
  // ---- the sweep.  No row predicates are needed: rows above the pivot only touch their (unused) upper
  //      triangle, and an identity row whose column has not been reached has a zero multiplier.  Every
  //      thread forms 1/pivot itself from the broadcast column, so the only cross-warp traffic per column
  //      is the publication of the next column. ----
  int kb = 0;
  bool fail = false;
  if (active_warp) {
    bool ok = true;
    if (CPL == 2) {
      ok = ldl_sweep_range<CPL, RPW, PR, NW, 0, 0>(v, jj, colbuf, dvec, 0, min(R, 31), R, kb);
      if (ok && R > 31) ok = ldl_sweep_range<CPL, RPW, PR, NW, 0, (CPL > 1 ? 1 : 0)>(v, jj, colbuf, dvec, 31, 32, R, kb);
      if (ok && R > 32) ok = ldl_sweep_range<CPL, RPW, PR, NW, (CPL > 1 ? 1 : 0), (CPL > 1 ? 1 : 0)>(v, jj, colbuf, dvec, 32, R, R, kb);
    } else {
#define VJF_SW(CNv, CNPv, a, b) if (ok && R > (a)) ok = ldl_sweep_range<CPL, RPW, PR, NW, (CNv) < CPL ? (CNv) : 0, (CNPv) < CPL ? (CNPv) : 0>(v, jj, colbuf, dvec, (a), min(R, (b)), R, kb);
      VJF_SW(0, 0, 0, 31) VJF_SW(0, 1, 31, 32) VJF_SW(1, 1, 32, 63) VJF_SW(1, 2, 63, 64)
      VJF_SW(2, 2, 64, 95) VJF_SW(2, 3, 95, 96) VJF_SW(3, 3, 96, 128)
#undef VJF_SW
    }
    fail = !ok;
    if (tid == 0) misc[0] = fail ? 1.f : 0.f;
  }
  __syncthreads();
  fail = misc[0] != 0.f;
  VJF_STAMP(p, t, 27);
  if (fail) return false;
  // ---- scale the columns ; z and w_chol = L^-T into shared memory ; W' = w_chol z ; publish ; residual ; THEN the rest of the
  //      commit.  What the trial CTAs of the next step wait for (w_chol / w_mean, in the tile pipeline their operand images)
  //      is written first -- the images by whole 128-byte rows from shared memory instead of one scattered store per element
  //      from the register layout -- and everything nobody waits for (w_pchol, w_precision, w_chol of the tile pipeline)
  //      after the residual. ----
  float* Lout = st + p.lay.w_pchol;
  float* Uout = st + p.lay.w_chol;
  float* Pout = st + p.lay.w_precision;
  float sdv[CPL], isd[CPL];
#pragma unroll
  for (int ci = 0; ci < CPL; ++ci) {
    const int j = lane + 32 * ci;
    sdv[ci] = (j < R) ? sqrtf(dvec[j]) : 1.f;
    isd[ci] = 1.0f / sdv[ci];
  }
  const bool images = p.tp.on != 0;
#pragma unroll
  for (int ri = 0; ri < RPW; ++ri) {
    const int r = active_warp ? warp + NW * ri : (1 << 20);
    if (r >= NR || r < R) continue;
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) {
      const int j = lane + 32 * ci;
      if (j >= R) continue;
      const float x = v[ri][ci] * isd[ci];
      if (r < R + d) {
        zbuf[(r - R) * R + j] = x;
      } else {
        const int c = r - R - d;
        const float uv = (j >= c) ? x : 0.f;
        if (!images) {
          Uout[c * R + j] = uv;
          if (p.use_tma) p.u_mirror[c * p.ldu + j] = uv;  // row-padded mirror: the TMA source of the next back half
        }
        Qs[c * ldq + j] = uv;
      }
    }
  }
  __syncthreads();
  VJF_STAMP(p, t, 28);
  // W'[c][i] = sum_{j >= c} w_chol[c][j] z[i][j] ; one thread per output
  float* Wout = st + p.lay.w_mean;
  for (int i = tid; i < R * d; i += VJF_NT) {
    const int c = i / d, k = i - c * d;
    float s0 = 0.f, s1 = 0.f;
    const float* q = Qs + c * ldq;
    const float* z = zbuf + k * R;
    int j = c;
    for (; j + 1 < R; j += 2) { s0 = fmaf(q[j], z[j], s0); s1 = fmaf(q[j + 1], z[j + 1], s1); }
    if (j < R) s0 = fmaf(q[j], z[j], s0);
    const float w = s0 + s1;
    Wout[i] = w; Ws[i] = w;
  }
  __syncthreads();
  if (images) {
    // operand images of the tile pipeline ([w_chol^T ; w_mean^T], hi | lo, K-major SW128; see uk_store): a warp writes one
    // 128-byte image row (32 contraction indices of one output column) per store
    const int nchunk = (p.tp.Rk + 31) >> 5, NQ = p.tp.NQ, lo_off = p.tp.ukimg >> 2;
    for (int task = warp; task < (R + d) * nchunk; task += VJF_NWARP) {
      const int row = task / nchunk, chunk = task - row * nchunk;
      const int c = chunk * 32 + lane;
      const float val = (c < R) ? (row < R ? Qs[c * ldq + row] : Ws[c * d + (row - R)]) : 0.f;
      const int nq = row < R ? row : p.tp.Rk + (row - R);
      const int off = (chunk * NQ + nq) * 32 + ((((lane >> 2) ^ (nq & 7)) << 2) | (lane & 3));
      p.uk[off] = val;
      p.uk[lo_off + off] = val - tf32_trunc_f(val);
    }
    __syncthreads();
  }
  // overlapped schedule: w_chol / w_mean of this step are final -- let the trial CTAs stage them for the back half of the
  // next step while the residual / state-noise update below is still running
  if (p.overlap && !p.init_mode && tid == 0) { __threadfence(); st_release_gpu_u32(p.ctrl + 5, (unsigned)(t + 1)); }
  VJF_STAMP(p, t, 29);
  // ---- sum |dx - phi W'|^2 - S = <W', A W'> - 2 <W', b> from the statistics still in registers / shared memory ----
  double acc = 0.0;
#pragma unroll
  for (int ri = 0; ri < PR; ++ri) {
    const int r = active_warp ? warp + NW * ri : (1 << 20);
    if (r < R) {
#pragma unroll
      for (int ci = 0; ci < CPL; ++ci) {
        const int j = lane + 32 * ci;
        if (j <= r) {  // lower triangle of the symmetric A, off-diagonal counted twice
          float w2 = 0.f;
          for (int k = 0; k < d; ++k) w2 = fmaf(Ws[r * d + k], Ws[j * d + k], w2);
          acc += (double)((j < r ? 2.0f : 1.0f) * a_in[ri][ci]) * (double)w2;
        }
      }
    }
  }
  for (int i = tid; i < R * d; i += VJF_NT) acc -= 2.0 * (double)Ws[i] * (double)bs[i];
  *resid_out = block_sum_d(acc, dred);
  // Overlapped schedules: the trial CTAs of the next step wait for the state-noise variance (ctrl[3]) right after the dynamics
  // read-out -- update it (vjf/model.py:373-377) and publish it HERE, before the part of the commit nobody waits for.  (The
  // caller finds *noise_done set and only reads out the losses; its own release of ctrl[3] repeats the same value.)
  if (noise_done) {
    *noise_done = false;
    if (p.overlap && !p.init_mode && p.lik == VJF_LIK_POISSON) {
      if (tid == 0) {
        const float Bf = (float)p.Bglobal, gam = st[p.lay.tr_logvar];
        const double tot = *resid_out + (double)p.reduced[p.ps + SC_SDX];
        const float mse = (float)(fmax(tot, 0.0) / ((double)p.Bglobal * (double)d));
        float n_new;
        const float var = running_var_f(expf(gam), st[p.lay.tr_n], mse, Bf, 500.f, &n_new);
        st[p.lay.tr_logvar] = logf(var);
        st[p.lay.tr_n] = n_new;
        __threadfence();
        st_release_gpu_u32(p.ctrl + 3, (unsigned)(t + 1));
      }
      *noise_done = true;
    }
  }
  // ---- the rest of the commit: w_pchol = L and (tile pipeline) w_chol from the registers, row by row; w_precision = P' both
  //      triangles through shared memory (the w_chol buffer is dead), so that the global rows are written whole ----
  __syncthreads();
#pragma unroll
  for (int ri = 0; ri < RPW; ++ri) {
    const int r = active_warp ? warp + NW * ri : (1 << 20);
    if (r >= NR || (r >= R && r < R + d)) continue;
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) {
      const int j = lane + 32 * ci;
      if (j >= R) continue;
      const float x = v[ri][ci] * isd[ci];
      if (r < R) {
        Lout[r * R + j] = (j < r) ? x : ((j == r) ? sdv[ci] : 0.f);
      } else if (images) {
        const int c = r - R - d;
        Uout[c * R + j] = (j >= c) ? x : 0.f;
      }
    }
  }
#pragma unroll
  for (int ri = 0; ri < PR; ++ri) {
    const int r = active_warp ? warp + NW * ri : (1 << 20);
    if (r >= R) continue;
#pragma unroll
    for (int ci = 0; ci < CPL; ++ci) {
      const int j = lane + 32 * ci;
      if (j <= r) { Qs[r * ldq + j] = v0[ri][ci]; Qs[j * ldq + r] = v0[ri][ci]; }
    }
  }
  __syncthreads();
  for (int i = tid; i < R * R; i += VJF_NT) { const int r = i / R, j = i - r * R; Pout[i] = Qs[r * ldq + j]; }
  return true;
}

// Fallback for n_rbf > 128: the same sweep with the work matrix in shared memory.
static __device__ bool rls_factor_smem(const StepParams& p, float* sm, float iv, const float* A, const float* bv) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, d = p.d, ldm = p.ldm;
  float* st = p.state;
  const float* P = st + p.lay.w_precision;
  const float* Wold = st + p.lay.w_mean;
  float* M = sm;                                   // [(2R+d)][ldm]
  float* dvec = sm + (size_t)(2 * R + d) * ldm;    // [R] pivots
  for (int i = tid; i < R * R; i += VJF_NT) {
    const int r = i / R, c = i - r * R;
    if (c <= r) M[r * ldm + c] = fmaf(A[i], iv, P[i]);
  }
  for (int i = tid; i < R * d; i += VJF_NT) {
    const int k = i / d, c = i - k * d;
    float s = 0.f;
    for (int j = 0; j < R; ++j) s = fmaf(P[k * R + j], Wold[j * d + c], s);
    M[(R + c) * ldm + k] = fmaf(bv[i], iv, s);
  }
  for (int i = tid; i < R * R; i += VJF_NT) {
    const int r = i / R, c = i - r * R;
    M[(R + d + r) * ldm + c] = (r == c) ? 1.0f : 0.f;
  }
  __syncthreads();
  float piv = M[0];
  float inv = 1.0f / piv;
  bool fail = !(piv > 0.f);
  for (int k = 0; k < R && !fail; ++k) {
    if (tid == 0) dvec[k] = piv;
    float pivn = 1.f;
    if (k + 1 < R) {
      const float l = M[(k + 1) * ldm + k];
      pivn = fmaf(-(l * inv), l, M[(k + 1) * ldm + k + 1]);
    }
    const int rend = R + d + k + 1;
    for (int r = k + 1 + warp; r < rend; r += VJF_NWARP) {
      const float tk = M[r * ldm + k] * inv;
      const int jend = (r < R) ? (r + 1) : R;
      for (int j = k + 1 + lane; j < jend; j += 32) {
        if (r == k + 1 && j == k + 1) continue;  // next pivot lives in registers
        M[r * ldm + j] = fmaf(-tk, M[j * ldm + k], M[r * ldm + j]);
      }
    }
    if (k + 1 < R && !(pivn > 0.f)) fail = true;
    piv = pivn;
    inv = 1.0f / pivn;
    __syncthreads();
  }
  if (fail) return false;
  __syncthreads();
  float* Lout = st + p.lay.w_pchol;
  float* Uout = st + p.lay.w_chol;
  float* Pout = st + p.lay.w_precision;
  for (int i = tid; i < R * R; i += VJF_NT) {
    const int r = i / R, c = i - r * R;
    const float sd = sqrtf(dvec[c]);
    Lout[i] = (c < r) ? M[r * ldm + c] / sd : ((c == r) ? sd : 0.f);
    Uout[i] = (c >= r) ? M[(R + d + r) * ldm + c] / sd : 0.f;
    if (p.use_tma) p.u_mirror[r * p.ldu + c] = Uout[i];
    if (p.tp.on) uk_store(p, c, r, Uout[i]);
    const int lo = (c <= r) ? i : (c * R + r);
    Pout[i] = fmaf(A[lo], iv, P[lo]);
  }
  for (int i = tid; i < R * d; i += VJF_NT) {
    const int k = i / d, c = i - k * d;
    M[(R + c) * ldm + k] = M[(R + c) * ldm + k] / sqrtf(dvec[k]);
  }
  __syncthreads();
  float* Wout = st + p.lay.w_mean;
  for (int i = tid; i < R * d; i += VJF_NT) {
    const int r = i / d, c = i - r * d;
    float s = 0.f;
    for (int k = r; k < R; ++k) s = fmaf(Uout[r * R + k], M[(R + c) * ldm + k], s);
    Wout[i] = s;
    if (p.tp.on) uk_store(p, p.tp.Rk + c, r, s);
  }
  __syncthreads();
  return true;
}

// Double-precision RLS for long runs (vjf_set_rls_precision(h, 64)).  The information-form recursion accumulates phi^T phi / v
// without forgetting (shrink = 1, vjf/model.py:371): after ~1e7 samples fp32 can neither absorb a step's increment into
// w_precision nor factorise it (the reference's fp32 run breaks the same way; its float64 configuration, script/example.py:12,
// does not).  Here w_precision lives in a double shadow `p.P64` (re-seeded from the fp32 state whenever the two disagree, i.e.
// after the caller loaded a state), the sweep of the augmented matrix [P'; g^T; I] runs in double in shared memory, and the
// fp32 state receives the rounded results.
static __device__ bool rls_factor_f64(const StepParams& p, float* smf, float ivf, const float* A, const float* bv) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, d = p.d, ldm = p.ldm;
  float* st = p.state;
  float* Pst = st + p.lay.w_precision;
  const float* Wold = st + p.lay.w_mean;
  double* P64 = p.P64;
  double* M = reinterpret_cast<double*>(smf);      // [(2R+d)][ldm]
  double* dvec = M + (size_t)(2 * R + d) * ldm;    // [R] pivots
  const double iv = (double)ivf;
  int mism = 0;
  for (int i = tid; i < R * R; i += VJF_NT) mism |= ((float)P64[i] != Pst[i]);
  if (__syncthreads_or(mism)) {
    for (int i = tid; i < R * R; i += VJF_NT) P64[i] = (double)Pst[i];
    __syncthreads();
  }
  for (int i = tid; i < R * R; i += VJF_NT) {
    const int r = i / R, c = i - r * R;
    if (c <= r) M[r * ldm + c] = fma((double)A[i], iv, P64[i]);
  }
  for (int i = tid; i < R * d; i += VJF_NT) {
    const int k = i / d, c = i - k * d;
    double s = 0.0;
    for (int j = 0; j < R; ++j) s = fma(P64[k * R + j], (double)Wold[j * d + c], s);
    M[(R + c) * ldm + k] = fma((double)bv[i], iv, s);
  }
  for (int i = tid; i < R * R; i += VJF_NT) {
    const int r = i / R, c = i - r * R;
    M[(R + d + r) * ldm + c] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  // 1/x from the fp32 reciprocal + two Newton steps in double (relative error ~1e-14; a double division is ~10x slower and
  // sits on the serial chain of the sweep)
  auto recip = [](double x) { double r = (double)__frcp_rn((float)x); r = fma(r, fma(-x, r, 1.0), r); return fma(r, fma(-x, r, 1.0), r); };
  double piv = M[0];
  bool fail = !(piv > 0.0 && piv < 1e37);
  double inv = recip(piv);
  for (int k = 0; k < R && !fail; ++k) {
    if (tid == 0) dvec[k] = piv;
    double pivn = 1.0;
    if (k + 1 < R) {
      const double l = M[(k + 1) * ldm + k];
      pivn = fma(-(l * inv), l, M[(k + 1) * ldm + k + 1]);
    }
    const int rend = R + d + k + 1;
    for (int r = k + 1 + warp; r < rend; r += VJF_NWARP) {
      const double tk = M[r * ldm + k] * inv;
      const int jend = (r < R) ? (r + 1) : R;
      for (int j = k + 1 + lane; j < jend; j += 32) {
        if (r == k + 1 && j == k + 1) continue;  // next pivot lives in registers
        M[r * ldm + j] = fma(-tk, M[j * ldm + k], M[r * ldm + j]);
      }
    }
    if (k + 1 < R && !(pivn > 0.0 && pivn < 1e37)) fail = true;
    piv = pivn;
    inv = recip(pivn);
    __syncthreads();
  }
  if (fail) return false;
  __syncthreads();
  float* Lout = st + p.lay.w_pchol;
  float* Uout = st + p.lay.w_chol;
  double* isdv = dvec + R;  // [R] 1/sqrt(pivot)
  for (int c = tid; c < R; c += VJF_NT) { const double sd = sqrt(dvec[c]); dvec[c] = sd; isdv[c] = 1.0 / sd; }
  __syncthreads();
  for (int i = tid; i < R * R; i += VJF_NT) {
    const int r = i / R, c = i - r * R;
    const double sd = dvec[c], isd = isdv[c];
    Lout[i] = (float)((c < r) ? M[r * ldm + c] * isd : ((c == r) ? sd : 0.0));
    const double uv = (c >= r) ? M[(R + d + r) * ldm + c] * isd : 0.0;
    M[(R + d + r) * ldm + c] = uv;  // each element is read and written by this thread only
    Uout[i] = (float)uv;
    if (p.use_tma) p.u_mirror[r * p.ldu + c] = (float)uv;
    if (p.tp.on) uk_store(p, c, r, (float)uv);
    const int lo = (c <= r) ? i : (c * R + r);
    const double pn = fma((double)A[lo], iv, P64[lo]);
    M[r * ldm + c] = pn;   // P' (full, symmetric) parked in the dead P rows until every thread has read the old P64
    Pst[i] = (float)pn;
  }
  for (int i = tid; i < R * d; i += VJF_NT) {
    const int k = i / d, c = i - k * d;
    M[(R + c) * ldm + k] = M[(R + c) * ldm + k] * isdv[k];
  }
  __syncthreads();
  for (int i = tid; i < R * R; i += VJF_NT) { const int r = i / R, c = i - r * R; P64[i] = M[r * ldm + c]; }
  float* Wout = st + p.lay.w_mean;
  for (int i = tid; i < R * d; i += VJF_NT) {
    const int r = i / d, c = i - r * d;
    double s = 0.0;
    for (int k = r; k < R; ++k) s = fma(M[(R + d + r) * ldm + k], M[(R + c) * ldm + k], s);
    Wout[i] = (float)s;
    if (p.tp.on) uk_store(p, p.tp.Rk + c, r, (float)s);
  }
  __syncthreads();
  return true;
}

// finmask: bit0 recon finite, bit1 dyn finite, bit2 entropy finite (from the reduced sums)
static __device__ void phase_b2(const StepParams& p, float* sm, int t, unsigned finmask, const unsigned* wait_ctr = nullptr,
                                unsigned wait_val = 0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, d = p.d;
  float* st = p.state;
  const float* red = p.reduced;
  const float* A = red + p.pa;
  const float* bv = red + p.pb;
  const float* scal = red + p.ps;
  const float Bf = (float)p.Bglobal;
  const bool warm = p.flags & VJF_FLAG_WARMUP;
  bool upd = p.flags & VJF_FLAG_UPDATE;
  if (p.world > 1 && (*reinterpret_cast<volatile unsigned*>(p.status) & VJF_ST_COMM_TIMEOUT)) upd = false;  // a peer was lost: commit nothing

  // Overlapped Poisson schedule: the caller hands over the statistics-ready counter instead of waiting itself.  When the
  // register factorisation runs, its statistics-independent part (old P, P W) is done first and the wait happens inside
  // it; the loss sums (one thread, nothing waits for them) are then written after the factorisation.
  const bool defer = wait_ctr && upd && !warm && p.R <= 128 && !p.init_mode && p.lik == VJF_LIK_POISSON && !p.rls64;
  if (wait_ctr && !defer) wait_counter(wait_ctr, wait_val);
  auto losses_and_likelihood = [&](int who) {
    if (p.world > 1) finmask = (isfinite(scal[SC_RECON]) ? 1u : 0u) | (isfinite(scal[SC_DYN]) ? 2u : 0u) | (isfinite(scal[SC_ENT]) ? 4u : 0u);
    if (tid == who && !p.init_mode) {
      unsigned stbits = 0;
      float l_recon = scal[SC_RECON] / Bf, l_dyn = scal[SC_DYN] / Bf, h = scal[SC_ENT] / Bf;
      if (!(finmask & 1u)) { l_recon = 0.f; stbits |= VJF_ST_RECON_NONFINITE; }
      if (!(finmask & 2u)) { l_dyn = 0.f; stbits |= VJF_ST_DYN_NONFINITE; }
      if (!(finmask & 4u)) { h = 0.f; stbits |= VJF_ST_ENTROPY_NONFINITE; }
      if (scal[SC_BADMSE] != 0.f) stbits |= VJF_ST_MSE_NONFINITE;
      float loss = l_recon - h;
      if (!warm) loss += l_dyn;
      if (p.losses) {
        float* o = p.losses + (size_t)t * 4;
        o[0] = loss; o[1] = -l_recon; o[2] = -l_dyn; o[3] = h;
      }
      if (stbits) atomicOr(p.status, stbits);
      // GaussianLikelihood.update (vjf/likelihood.py:28-40): reads the post-SGD logvar
      if (upd && p.lik == VJF_LIK_GAUSSIAN) {
        const float mse = scal[SC_SSE] / (Bf * (float)p.D);
        float n_new;
        const float var = running_var_f(expf(st[p.lay.lik_logvar]), st[p.lay.lik_n], mse, Bf, 1000.f, &n_new);
        st[p.lay.lik_logvar] = logf(var);
        st[p.lay.lik_n] = n_new;
      }
      // overlapped schedule: the front half of step t+1 runs concurrently and needs this logvar for its decoder stage --
      // publish "logvar of step t is final" (the trial CTAs wait for it there)
      if (p.overlap && p.lik == VJF_LIK_GAUSSIAN) { __threadfence(); st_release_gpu_u32(p.ctrl + 4, (unsigned)(t + 1)); }
    }
  };
  if (!defer) losses_and_likelihood(0);
  VJF_STAMP(p, t, 24);
  if (p.ext) return;  // large n_rbf: the RLS and the state-noise update run as their own kernels (bigr.cu)
  if (!upd) {
    if (p.overlap && !p.init_mode && tid == 0) { __threadfence(); st_release_gpu_u32(p.ctrl + 5, (unsigned)(t + 1)); }
    return;
  }
  const float gam = st[p.lay.tr_logvar];
  // rls(x, target, v): v = exp(state logvar) in a filter step (model.py:371); in initialize v is the
  // mean squared increment (model.py:384-385)
  const float iv = p.init_mode ? (Bf * (float)d) / scal[SC_SDX] : 1.0f / expf(gam);
  __syncthreads();
  bool have_resid = false, noise_done = false;  // noise_done: the factorisation already updated and published the state-noise variance
  double resid = 0.0;
  if (!warm) {
    bool ok;
    const int nr16 = (2 * R + d + VJF_NWARP - 1) / VJF_NWARP;  // rows per warp of the work matrix
    const int nr8 = (2 * R + d + 7) / 8;  // rows per warp with 8 sweep warps
    (void)nr16;
    (void)nr8;
    if (p.rls64) ok = rls_factor_f64(p, sm, iv, A, bv);
    else if (R <= 64 && nr16 <= 5) { ok = rls_factor_regs<2, 5, 16>(p, sm, iv, A, bv, t, &resid, defer ? wait_ctr : nullptr, wait_val, &noise_done); have_resid = ok; }
    else if (R <= 64 && nr16 <= 7) { ok = rls_factor_regs<2, 7, 16>(p, sm, iv, A, bv, t, &resid, defer ? wait_ctr : nullptr, wait_val, &noise_done); have_resid = ok; }
    else if (R <= 64) { ok = rls_factor_regs<2, 9, 16>(p, sm, iv, A, bv, t, &resid, defer ? wait_ctr : nullptr, wait_val, &noise_done); have_resid = ok; }
    else if (R <= 128 && nr16 <= 13) { ok = rls_factor_regs<4, 13, 16>(p, sm, iv, A, bv, t, &resid, defer ? wait_ctr : nullptr, wait_val, &noise_done); have_resid = ok; }
    else if (R <= 128) { ok = rls_factor_regs<4, 17, 16>(p, sm, iv, A, bv, t, &resid, defer ? wait_ctr : nullptr, wait_val, &noise_done); have_resid = ok; }
    else ok = rls_factor_smem(p, sm, iv, A, bv);
    if (!ok && tid == 0) atomicOr(p.status, (unsigned)VJF_ST_CHOL_FAILED);
    __syncthreads();
  }
  // every path (failed factorisation, shared-memory fallback, warm-up) ends with final RLS outputs here
  if (p.overlap && !p.init_mode && tid == 0) { __threadfence(); st_release_gpu_u32(p.ctrl + 5, (unsigned)(t + 1)); }

  // ---- state-noise running variance (vjf/model.py:373-377).  sum |dx - phi W'|^2 from the reduced
  //      statistics: S - 2 <W', b> + <W', A W'>, evaluated in double ----
  double tot;
  if (have_resid) {
    tot = resid + (double)scal[SC_SDX];
  } else {
    float* wbuf = sm;                                      // [R][d] current W
    const size_t doff = ((size_t)R * d + 5) & ~(size_t)1;  // 8-byte aligned
    double* dred = reinterpret_cast<double*>(sm + doff);   // [NWARP]
    for (int i = tid; i < R * d; i += VJF_NT) wbuf[i] = st[p.lay.w_mean + i];
    __syncthreads();
    double acc = 0.0;
    for (int r = warp; r < R; r += VJF_NWARP) {
      for (int j = lane; j <= r; j += 32) {  // lower triangle of the symmetric A, off-diagonal counted twice
        double w2 = 0.0;
        for (int k = 0; k < d; ++k) w2 += (double)wbuf[r * d + k] * (double)wbuf[j * d + k];
        acc += (j < r ? 2.0 : 1.0) * (double)A[r * R + j] * w2;
      }
    }
    for (int i = tid; i < R * d; i += VJF_NT) acc -= 2.0 * (double)wbuf[i] * (double)bv[i];
    tot = block_sum_d(acc, dred) + (double)scal[SC_SDX];
  }
  {
    if (tid == 0 && !noise_done) {
      const float mse = (float)(fmax(tot, 0.0) / ((double)p.Bglobal * (double)d));
      if (p.init_mode) { st[p.lay.tr_logvar] = logf(mse); return; }  // model.py:387-388
      float n_new;
      const float var = running_var_f(expf(gam), st[p.lay.tr_n], mse, Bf, 500.f, &n_new);
      st[p.lay.tr_logvar] = logf(var);
      st[p.lay.tr_n] = n_new;
    }
  }
  // deferred loss read-out: by a thread of another warp, concurrently with thread 0's state-noise update above
  if (defer) losses_and_likelihood(32);
}

// finite flags of the three ELBO terms from the slots (every CTA evaluates this identically)
static __device__ unsigned term_finite_mask(const StepParams& p, const float* src, int nslots, float* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned* flag = reinterpret_cast<unsigned*>(sm + p.s_b1);
  if (warp == 0) {
    unsigned m = 0;
    for (int i = 0; i < 3; ++i) {
      float s = 0.f;
      for (int c = lane; c < nslots; c += 32) s += src[(size_t)c * p.PS + p.ps + i];
      s = warp_sum(s);
      if (isfinite(s)) m |= 1u << i;
    }
    if (lane == 0) *flag = m;
  }
  __syncthreads();
  const unsigned m = *flag;
  __syncthreads();
  return m;
}
