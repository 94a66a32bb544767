// Large n_rbf (BASELINE.json configs[2], "C3": n_rbf = 1024, 16 384 trials): the two contractions of the RBF read-out that are
// O(B R^2) -- FL = phi w_chol (vjf/module.py:75-77) and phi^T phi (module.py:94-96) -- as tcgen05 GEMMs over ALL trials of the
// step, and LinearRegression.rls (module.py:89-102: Cholesky, cholesky_solve, inv(L^T)) as a blocked right-looking factorisation
// spread over every SM.  Below VJF_BIGR_MIN the whole step lives in one persistent kernel (k_persistent.cu / k_tile.cu) where
// w_chol fits in shared memory and the factorisation in one CTA's registers; here R x R is 4 MB and the step is a short
// sequence of launches per time step:
//
//   features   xs = reparametrize(q_{t-1}); phi (B x R) and its transpose; p_mean = xs + phi W          (model.py:112-113, :338)
//   GEMM 1     |phi w_chol|^2 row sums -> p_logvar                                                      (module.py:75-77)
//   phase A    the tile kernel of the split path with the dynamics read-out handed in (StepParams::ext): recognition, decoder,
//              likelihood, ELBO, hand-derived backward, slot sums                                       (model.py:97-154, :209)
//   reduce + phase B   clip + SGD, losses, GaussianLikelihood.update                                    (model.py:210-211)
//   GEMM 2     phi^T phi (lower tiles, split over the trials) ; phi^T dx                                (module.py:94-96)
//   factor     [P'; g^T; I] -> [L; (L^-1 g)^T; L^-T] by 64-column panels, W' = L^-T L^-1 g             (module.py:99-102)
//   residual   mean |dx - phi W'|^2 -> running state-noise variance                                     (model.py:373-377)
//
// GEMM kernel: C = A B^T with A (M x K) and B (N x K) row-major fp32.  128 x 256 output tile per CTA, fp32 accumulators in 256
// tensor-memory columns, K in chunks of 32 floats (one 128-byte swizzle atom) through a 2-stage ring: a producer warp issues the
// TMA tensor loads (SWIZZLE_128B, out-of-range rows / columns zero-filled), eight warps split the landed fp32 chunk in place
// into hi (what the tensor core keeps when it truncates to tf32) and lo = x - hi, one warp issues tcgen05.mma kind::tf32 three
// times per k-step (lo*hi, hi*lo, hi*hi: fp32-grade products), tcgen05.commit frees the stage.  The same eight warps read the
// accumulator back (tcgen05.ld) for the epilogue: row sums of squares (GEMM 1) or a plain tile store (GEMM 2).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdio>

#include <cuda.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

struct BigR {
  int Bmax, Bp, SK, NP;
  float *phi, *phiT, *xs, *pm, *plv, *dx, *part1, *Ut, *Apart, *bstat, *M, *Pnew;  // bstat: [BSPLIT][R][d] partial sums
  double* rpart;
  unsigned* ctl;  // [0] grid barrier, [1] fail flag
};

constexpr int BIGR_BSPLIT = 8;   // phi^T dx: the trials are split over this many CTAs per group of 8 RBFs (bytes in flight)
constexpr int BIGR_FT = 16;      // trials per CTA of the feature kernel

namespace bg {
constexpr int BM = 128, BN = 256;
constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, RAW_BYTES = A_BYTES + B_BYTES;
constexpr int STAGE_BYTES = 2 * RAW_BYTES;  // raw chunk + its lo companion
constexpr int STAGES = 2;
constexpr int NT = 320;                     // warp 0 TMA, warp 1 MMA, warps 2..9 split + epilogue
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;

struct Args {
  int M, N, K;
  int tri;    // B[n][k] = 0 for k > n (w_chol^T): K chunks past the tile's last column are skipped
  int sym;    // C symmetric: tiles strictly above the diagonal band are skipped
  int mode;   // 0: out[(2 * ntile + half) * ldo + row] = sum of squares over the tile's columns; 1: out[z][row][col] = C tile
  float* out;
  int ldo;
};

__device__ __forceinline__ uint64_t kmaj_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B: LBO 16 (unused), SBO 1024 (8 rows x 128 B)
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(16 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spins > (1u << 27)) __trap();  // a protocol error fails the launch instead of hanging the device
  }
}
__device__ __forceinline__ void arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tensor2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
}  // namespace bg

__global__ void __launch_bounds__(bg::NT, 1)
bigr_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const bg::Args g) {
  using namespace bg;
  if (g.sym && blockIdx.x < 2 * blockIdx.y) return;
  extern __shared__ unsigned char smraw[];
  unsigned char* sb = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + STAGES * STAGE_BYTES);  // full[2] split[2] empty[2] accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int NK = (g.K + 31) >> 5;
  if (g.tri) NK = min(NK, (min(n0 + BN, g.N) + 31) >> 5);
  const int per = (NK + gridDim.z - 1) / gridDim.z;
  const int kc0 = blockIdx.z * per, nk = max(0, min(NK, kc0 + per) - kc0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bars + s, 1); mbar_init(bars + 2 + s, 8); mbar_init(bars + 4 + s, 1); }
    mbar_init(bars + 6, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc256(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i & 1, u = i >> 1;
        if (u >= 1) wait(bars + 4 + s, (u - 1) & 1);
        unsigned char* raw = sb + s * STAGE_BYTES;
        mbar_expect_tx(bars + s, RAW_BYTES);
        tensor2d(raw, &mapA, (kc0 + i) * 32, m0, bars + s);
        tensor2d(raw + A_BYTES, &mapB, (kc0 + i) * 32, n0, bars + s);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_tf32(BM, BN);
    for (int i = 0; i < nk; ++i) {
      const int s = i & 1, u = i >> 1;
      wait(bars + 2 + s, u & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(sb + s * STAGE_BYTES), b_hi = a_hi + A_BYTES, a_lo = a_hi + RAW_BYTES, b_lo = a_lo + A_BYTES;
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // four k-steps of 8 inside the 128-byte atom
          const uint32_t o = j * 32;
          umma_tf32_ss(tmem, kmaj_desc(a_lo + o), kmaj_desc(b_hi + o), idesc, (i > 0 || j > 0) ? 1u : 0u);
          umma_tf32_ss(tmem, kmaj_desc(a_hi + o), kmaj_desc(b_lo + o), idesc, 1u);
          umma_tf32_ss(tmem, kmaj_desc(a_hi + o), kmaj_desc(b_hi + o), idesc, 1u);
        }
        umma_commit(bars + 4 + s);
        if (i == nk - 1) umma_commit(bars + 6);
      }
      __syncwarp();
    }
  } else {
    const int t8 = threadIdx.x - 64;
    for (int i = 0; i < nk; ++i) {
      const int s = i & 1, u = i >> 1;
      wait(bars + s, u & 1);
      const float4* raw = reinterpret_cast<const float4*>(sb + s * STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(sb + s * STAGE_BYTES + RAW_BYTES);
#pragma unroll 4
      for (int v = t8; v < RAW_BYTES / 16; v += 256) {
        const float4 x = raw[v];
        float4 l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
        l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
        l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        lo[v] = l;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) arrive(bars + 2 + s);
    }
    // ---- epilogue: this warp reaches the tensor-memory lanes of quarter warp % 4; the two warps of a quarter take 128 columns each
    if (nk > 0) { wait(bars + 6, 0); tc_fence_after(); }
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = m0 + q * 32 + lane;
    float ssq = 0.f;
    for (int cb = 0; cb < 4; ++cb) {
      const int col0 = half * 128 + cb * 32;
      float v[32];
      if (nk > 0) tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + col0, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (g.mode == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ssq = fmaf(v[j], v[j], ssq);
      } else if (row < g.M) {
        float* o = g.out + ((size_t)blockIdx.z * g.M + row) * g.ldo + n0 + col0;
        if (n0 + col0 + 32 <= g.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32; ++j) if (n0 + col0 + j < g.N) o[j] = v[j];
        }
      }
    }
    if (g.mode == 0 && row < g.M) g.out[(size_t)(blockIdx.y * 2 + half) * g.ldo + row] = ssq;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_free256(tmem);
}

// ------------------------------------------------------------------------------------------
// features: xs, phi, phi^T, p_mean  (vjf/util.py:11-13, vjf/functional.py:11-22, vjf/model.py:338)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) bigr_feat_kernel(const StepParams p, const BigR w, int t) {
  extern __shared__ __align__(16) float sm[];
  constexpr int FT = BIGR_FT;
  const int tid = threadIdx.x, R = p.R, d = p.d, u = p.u, du = p.du, ldt = R + 1;
  const int du4 = (du + 3) >> 2, ldx = 4 * du4;  // rows of xu padded with zeros to whole float4 groups
  float* xu_s = sm;                 // [FT][ldx]  (first: 16-byte aligned)
  float* tile = xu_s + FT * ldx;    // [FT][R + 1]
  float* Wt = tile + ((FT * ldt + 3) & ~3);  // [d][R]  w_mean transposed: conflict-free reads with the lanes over the RBFs
  const int b0 = blockIdx.x * FT, nb = min(FT, p.B - b0);
  const float* st = p.state;
  const bool prior = (t == 0) && (p.flags & VJF_FLAG_PRIOR_Q0);
  const float* qm = (t == 0) ? p.q0m : p.mu + (size_t)(t - 1) * p.B * d;
  const float* ql = (t == 0) ? p.q0l : p.logvar + (size_t)(t - 1) * p.B * d;
  for (int i = tid; i < FT * ldx; i += blockDim.x) {
    const int b = i / ldx, k = i - b * ldx;
    float v = 0.f;
    if (b < nb && k < du) {
      if (k < d) {
        const float m = prior ? st[p.lay.prior_mean + k] : qm[(size_t)(b0 + b) * d + k];
        const float l = prior ? st[p.lay.prior_logvar + k] : ql[(size_t)(b0 + b) * d + k];
        float e;
        if (p.eps) e = p.eps[((size_t)t * 2 * p.B + b0 + b) * d + k];
        else { float z[4]; philox_normal4(p.seed, p.step0 + t, p.trial_offset + b0 + b, 0, k >> 2, z); e = z[k & 3]; }
        v = m + e * expf(0.5f * l);
        w.xs[(size_t)(b0 + b) * d + k] = v;
      } else {
        v = p.u_in[((size_t)t * p.B + b0 + b) * u + (k - d)];
      }
    }
    xu_s[i] = v;
  }
  for (int i = tid; i < R * d; i += blockDim.x) { const int r = i / d, k = i - r * d; Wt[k * R + r] = st[p.lay.w_mean + i]; }
  __syncthreads();
  const float* cen = st + p.lay.centroid;
  const float* lw = st + p.lay.logwidth;
  if (du4 <= 4) {
    // two RBFs per thread with their centres in registers; a trial's inputs arrive as (broadcast) 128-bit shared-memory loads
    for (int rb = 0; rb < R; rb += 2 * blockDim.x) {
      const int r0 = rb + tid, r1 = rb + tid + blockDim.x;
      float4 c0[4], c1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float t0[4], t1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 4 * q + e;
          t0[e] = (q < du4 && k < du && r0 < R) ? cen[r0 * du + k] : 0.f;
          t1[e] = (q < du4 && k < du && r1 < R) ? cen[r1 * du + k] : 0.f;
        }
        c0[q] = make_float4(t0[0], t0[1], t0[2], t0[3]); c1[q] = make_float4(t1[0], t1[1], t1[2], t1[3]);
      }
      float iw0 = 0.f, iw1 = 0.f;
      if (r0 < R) { const float wd = expf(lw[r0]); iw0 = -0.5f / (wd * wd); }
      if (r1 < R) { const float wd = expf(lw[r1]); iw1 = -0.5f / (wd * wd); }
#pragma unroll 2
      for (int b = 0; b < FT; ++b) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < du4) {
            const float4 x = *reinterpret_cast<const float4*>(xu_s + b * ldx + 4 * q);
            float df;
            df = x.x - c0[q].x; a0 = fmaf(df, df, a0); df = x.y - c0[q].y; a0 = fmaf(df, df, a0);
            df = x.z - c0[q].z; a0 = fmaf(df, df, a0); df = x.w - c0[q].w; a0 = fmaf(df, df, a0);
            df = x.x - c1[q].x; a1 = fmaf(df, df, a1); df = x.y - c1[q].y; a1 = fmaf(df, df, a1);
            df = x.z - c1[q].z; a1 = fmaf(df, df, a1); df = x.w - c1[q].w; a1 = fmaf(df, df, a1);
          }
        const bool live = b < nb;
        const float v0 = live ? expf(a0 * iw0) : 0.f, v1 = live ? expf(a1 * iw1) : 0.f;
        if (r0 < R) { if (live) w.phi[(size_t)(b0 + b) * R + r0] = v0; tile[b * ldt + r0] = v0; }
        if (r1 < R) { if (live) w.phi[(size_t)(b0 + b) * R + r1] = v1; tile[b * ldt + r1] = v1; }
      }
    }
  } else
  for (int r = tid; r < R; r += blockDim.x) {
    float c[2 * VJF_MAX_XDIM];
    for (int k = 0; k < du; ++k) c[k] = cen[r * du + k];
    const float wd = expf(lw[r]);
    const float iw = -0.5f / (wd * wd);
    for (int b = 0; b < FT; ++b) {
      float v = 0.f;
      if (b < nb) {
        float d2 = 0.f;
        for (int k = 0; k < du; ++k) { const float df = xu_s[b * ldx + k] - c[k]; d2 = fmaf(df, df, d2); }
        v = expf(d2 * iw);
        w.phi[(size_t)(b0 + b) * R + r] = v;
      }
      tile[b * ldt + r] = v;
    }
  }
  __syncthreads();
  for (int i = tid; i < FT * R; i += blockDim.x) {
    const int r = i / FT, b = i % FT;
    if (b0 + b < w.Bp) w.phiT[(size_t)r * w.Bp + b0 + b] = tile[b * ldt + r];
  }
  // p_mean = xs + phi W: one warp per trial, lanes over the RBFs
  {
    const int b = tid >> 5, lane = tid & 31;
    float acc[VJF_MAX_XDIM];
#pragma unroll
    for (int k = 0; k < VJF_MAX_XDIM; ++k) acc[k] = 0.f;
    for (int r = lane; r < R; r += 32) {
      const float ph = tile[b * ldt + r];
#pragma unroll
      for (int k = 0; k < VJF_MAX_XDIM; ++k)
        if (k < d) acc[k] = fmaf(ph, Wt[k * R + r], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < VJF_MAX_XDIM; ++k) {
      const float sum = warp_sum(acc[k]);
      if (lane == 0 && k < d && b < nb) w.pm[(size_t)(b0 + b) * d + k] = xu_s[b * ldx + k] + sum;
    }
  }
}

// p_logvar = log sum_n FL^2 from the per-tile partial sums (fixed order)
__global__ void bigr_plv_kernel(const BigR w, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int j = 0; j < w.NP; ++j) s += w.part1[(size_t)j * w.Bmax + b];
  w.plv[b] = logf(s);
}

// w_chol^T (the B operand of GEMM 1, K-major) from the state
__global__ void bigr_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R) {
  __shared__ float t[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = y0 + j, c = x0 + threadIdx.x;
    t[j][threadIdx.x] = (r < R && c < R) ? in[(size_t)r * R + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = x0 + j, c = y0 + threadIdx.x;
    if (r < R && c < R) out[(size_t)r * R + c] = t[threadIdx.x][j];
  }
}

// b = phi^T dx (R x d): one warp per RBF, lanes over the trials; the dx chunk of 256 trials is staged once per CTA
__global__ void __launch_bounds__(256) bigr_bstat_kernel(const StepParams p, const BigR w) {
  __shared__ float dxs[256 * VJF_MAX_XDIM];
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, d = p.d;
  float acc[VJF_MAX_XDIM];
#pragma unroll
  for (int k = 0; k < VJF_MAX_XDIM; ++k) acc[k] = 0.f;
  const float* row = w.phiT + (size_t)min(r, p.R - 1) * w.Bp;
  const int per = (((p.B + BIGR_BSPLIT - 1) / BIGR_BSPLIT) + 255) & ~255;
  const int cbeg = blockIdx.y * per, cend = min(p.B, cbeg + per);
  for (int c0 = cbeg; c0 < cend; c0 += 256) {
    const int nb = min(256, cend - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * d; i += 256) dxs[i] = w.dx[(size_t)c0 * d + i];
    __syncthreads();
    float ph[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int b = lane + 32 * i; ph[i] = (b < nb) ? row[c0 + b] : 0.f; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int b = lane + 32 * i;
      if (b < nb) {
#pragma unroll
        for (int k = 0; k < VJF_MAX_XDIM; ++k)
          if (k < d) acc[k] = fmaf(ph[i], dxs[b * d + k], acc[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VJF_MAX_XDIM; ++k) {
    const float sum = warp_sum(acc[k]);
    if (lane == 0 && k < d && r < p.R) w.bstat[((size_t)blockIdx.y * p.R + r) * d + k] = sum;
  }
}

// work matrix of the factorisation: rows [0, R) lower triangle of P' = P + A / v, rows [R, R + d) g^T with g = P W + b / v,
// rows [R + d, 2R + d) the identity (vjf/module.py:89-97)
__global__ void bigr_setup_kernel(const StepParams p, const BigR w) {
  const int R = p.R, d = p.d;
  const float* st = p.state;
  const float iv = 1.0f / expf(st[p.lay.tr_logvar]);
  const float* P = st + p.lay.w_precision;
  const size_t n = (size_t)(2 * R + d) * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / R), c = (int)(i - (size_t)r * R);
    float v = 0.f;
    if (r < R) {
      const int hi = max(r, c), lo = min(r, c);
      float a = 0.f;
      for (int z = 0; z < w.SK; ++z) a += w.Apart[((size_t)z * R + hi) * R + lo];
      v = fmaf(a, iv, P[(size_t)hi * R + lo]);
      w.Pnew[i] = v;
      if (c > r) v = 0.f;
    } else if (r < R + d) {
      continue;  // g^T: bigr_g_kernel
    } else {
      v = (c == r - R - d) ? 1.0f : 0.f;
    }
    w.M[i] = v;
  }
}

// g = P W + b / v (old P, old W: vjf/module.py:93) into rows [R, R + d) of the work matrix: one warp per row of P
__global__ void __launch_bounds__(256) bigr_g_kernel(const StepParams p, const BigR w) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, R = p.R, d = p.d;
  if (c >= R) return;
  const float* st = p.state;
  const float iv = 1.0f / expf(st[p.lay.tr_logvar]);
  const float* prow = st + p.lay.w_precision + (size_t)c * R;
  const float* W = st + p.lay.w_mean;
  float acc[VJF_MAX_XDIM];
#pragma unroll
  for (int k = 0; k < VJF_MAX_XDIM; ++k) acc[k] = 0.f;
  for (int j = lane; j < R; j += 32) {
    const float pv = prow[j];
#pragma unroll
    for (int k = 0; k < VJF_MAX_XDIM; ++k)
      if (k < d) acc[k] = fmaf(pv, W[j * d + k], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < VJF_MAX_XDIM; ++k) {
    const float sum = warp_sum(acc[k]);
    if (lane == 0 && k < d) {
      float bsum = 0.f;
      for (int z = 0; z < BIGR_BSPLIT; ++z) bsum += w.bstat[((size_t)z * R + c) * d + k];
      w.M[(size_t)(R + k) * R + c] = fmaf(bsum, iv, sum);
    }
  }
}

// ------------------------------------------------------------------------------------------
// blocked right-looking Cholesky of the augmented matrix, 64-column panels, every SM (cooperative launch):
//   panel k:  L11 = chol(A11) (each CTA redundantly, shared memory) ; rows below and the appended rows: X <- X L11^-T (one warp per
//   row, forward substitution) ; grid barrier ; trailing update of 64 x 64 tiles C -= X_i X_j^T ; grid barrier.
// On exit rows [0,R) hold L, rows [R,R+d) hold (L^-1 g)^T, rows [R+d,2R+d) hold L^-T = w_chol; then the commit and W' = w_chol z.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) bigr_factor_kernel(const StepParams p, const BigR w) {
  extern __shared__ __align__(16) float sm[];
  constexpr int FB = 64, LD = 65;
  float* Ds = sm;             // [64][65] diagonal block -> L11
  float* As = Ds + FB * LD;   // [64][65]
  float* Bs = As + FB * LD;   // [64][65]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, d = p.d, ld = R;
  float* M = w.M;
  unsigned target = 0;
  const int nblk = (R + FB - 1) / FB;
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(w.ctl + 16);  // development aid: globaltimer of CTA 0 per panel phase
#define BIGR_STAMP(i) if (blockIdx.x == 0 && tid == 0 && kb < 32) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); stamps[kb * 4 + (i)] = t_; }
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb * FB, nbk = min(FB, R - k0);
    BIGR_STAMP(0)
    // ---- L11 = chol(A11) and U = L11^-T, every CTA redundantly.  The 64 x 64 block and an appended identity live in
    //      registers (16 x 16 threads x 4 x 4 elements each); per column only the current column travels through shared memory
    //      (double-buffered: one block barrier per column).  The appended rows receive the same column operations, i.e. they
    //      end as I L11^-T. ----
    const int ty = tid >> 4, tx = tid & 15;
    float dv[4][4], ev[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = 4 * ty + a, c = 4 * tx + b;
        dv[a][b] = (r < nbk && c <= r) ? __ldcg(M + (size_t)(k0 + r) * ld + k0 + c) : ((r == c) ? 1.0f : 0.f);
        ev[a][b] = (r == c) ? 1.0f : 0.f;
      }
    float* colD = As;           // [2][64]   (As / Bs are free until the trailing update)
    float* colE = As + 128;     // [2][64]
#pragma unroll 1
    for (int jb = 0; jb < 16; ++jb) {
#pragma unroll
      for (int cj = 0; cj < 4; ++cj) {
        const int j = 4 * jb + cj, buf = (cj & 1) * 64;
        if (tx == jb) {
#pragma unroll
          for (int a = 0; a < 4; ++a) { colD[buf + 4 * ty + a] = dv[a][cj]; colE[buf + 4 * ty + a] = ev[a][cj]; }
        }
        __syncthreads();
        const float piv = colD[buf + j];
        if (!(piv > 0.f) || !(piv < 1e37f)) {  // every CTA holds the same block: all of them leave together
          if (blockIdx.x == 0 && tid == 0) { atomicOr(p.status, (unsigned)VJF_ST_CHOL_FAILED); w.ctl[1] = 1u; }
          return;
        }
        float rs = rsqrtf(piv);
        rs = rs * fmaf(-0.5f * piv, rs * rs, 1.5f);  // one Newton step: < 1 ulp, off the slow division path
        float la[4], lb[4], le[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          la[a] = (4 * ty + a > j) ? colD[buf + 4 * ty + a] * rs : 0.f;
          lb[a] = (4 * tx + a > j) ? colD[buf + 4 * tx + a] * rs : 0.f;
          le[a] = colE[buf + 4 * ty + a] * rs;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) { dv[a][b] = fmaf(-la[a], lb[b], dv[a][b]); ev[a][b] = fmaf(-le[a], lb[b], ev[a][b]); }
        if (tx == jb) {
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int r = 4 * ty + a;
            dv[a][cj] = (r > j) ? la[a] : ((r == j) ? piv * rs : 0.f);
            ev[a][cj] = le[a];
          }
        }
      }
    }
    __syncthreads();
    float* Us = Bs;  // [64][65] U = L11^-T (upper triangular)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = 4 * ty + a, c = 4 * tx + b;
        Ds[r * LD + c] = (c <= r) ? dv[a][b] : 0.f;
        Us[r * LD + c] = (c >= r) ? ev[a][b] : 0.f;
      }
    __syncthreads();
    if (blockIdx.x == 0) {
      for (int i = tid; i < nbk * nbk; i += 256) {
        const int r = i / nbk, c = i - r * nbk;
        if (c <= r) M[(size_t)(k0 + r) * ld + k0 + c] = Ds[r * LD + c];
      }
    }
    BIGR_STAMP(1)
    // ---- panel rows [k0 + nbk, R + d + k0 + nbk) (the identity rows of later columns are still zero in this panel):
    //      X <- X L11^-T = X U, one warp per row ----
    const int r_lo = k0 + nbk, nrows = R + d;
    {
      const int per = (nrows + gridDim.x - 1) / gridDim.x;
      const int my0 = blockIdx.x * per, my1 = min(nrows, my0 + per);
      for (int rr = my0 + warp; rr < my1; rr += 8) {
        float* rowp = M + (size_t)(r_lo + rr) * ld + k0;
        const float x0 = (lane < nbk) ? __ldcg(rowp + lane) : 0.f, x1 = (lane + 32 < nbk) ? __ldcg(rowp + lane + 32) : 0.f;
        float y0 = 0.f, y1 = 0.f;
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
          const float xc = __shfl_sync(0xffffffffu, x0, c);
          y0 = fmaf(xc, Us[c * LD + lane], y0); y1 = fmaf(xc, Us[c * LD + lane + 32], y1);
        }
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
          const float xc = __shfl_sync(0xffffffffu, x1, c);
          y1 = fmaf(xc, Us[(c + 32) * LD + lane + 32], y1);  // U[c + 32][lane] = 0 below the diagonal block
        }
        if (lane < nbk) rowp[lane] = y0;
        if (lane + 32 < nbk) rowp[lane + 32] = y1;
      }
    }
    if (kb == nblk - 1) break;
    BIGR_STAMP(2)
    grid_barrier(w.ctl, target);
    BIGR_STAMP(3)
    // ---- trailing update: rows [r_lo, r_lo + R + d) x columns [r_lo, R) in 64 x 64 tiles ----
    {
      const int ncb = (R - r_lo + FB - 1) / FB, nrb = (nrows + FB - 1) / FB;
      for (int tl = blockIdx.x; tl < nrb * ncb; tl += gridDim.x) {
        const int rb = tl / ncb, cb = tl - rb * ncb;
        const int row0 = r_lo + rb * FB, col0 = r_lo + cb * FB;
        const int nr = min(FB, r_lo + nrows - row0), nc = min(FB, R - col0);
        if (row0 + nr - 1 < col0 && row0 + nr - 1 < R) continue;  // a tile of P' entirely above the diagonal
        __syncthreads();
        {
          float4 va[4], vb[4];
          const int c4 = (tid & 15) << 2;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = (tid >> 4) + 16 * i;
            va[i] = (r < nr && c4 < nbk) ? __ldcg(reinterpret_cast<const float4*>(M + (size_t)(row0 + r) * ld + k0 + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            vb[i] = (r < nc && c4 < nbk) ? __ldcg(reinterpret_cast<const float4*>(M + (size_t)(col0 + r) * ld + k0 + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = (tid >> 4) + 16 * i;
            float* ap = As + r * LD + c4; float* bp = Bs + r * LD + c4;
            ap[0] = va[i].x; ap[1] = va[i].y; ap[2] = va[i].z; ap[3] = va[i].w;
            bp[0] = vb[i].x; bp[1] = vb[i].y; bp[2] = vb[i].z; bp[3] = vb[i].w;
          }
        }
        __syncthreads();
        const int ty = tid >> 4, tx = tid & 15;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        for (int k = 0; k < FB; ++k) {
          float av[4], bv[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) { av[a] = As[(ty * 4 + a) * LD + k]; bv[a] = Bs[(tx * 4 + a) * LD + k]; }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        if (tx * 4 + 3 < nc) {  // R is a multiple of 4 and so are col0 and k0: whole float4 groups
          float4 cv[4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
            cv[a] = (ty * 4 + a < nr) ? __ldcg(reinterpret_cast<const float4*>(M + (size_t)(row0 + ty * 4 + a) * ld + col0 + tx * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int a = 0; a < 4; ++a)
            if (ty * 4 + a < nr)
              *reinterpret_cast<float4*>(M + (size_t)(row0 + ty * 4 + a) * ld + col0 + tx * 4) =
                  make_float4(cv[a].x - acc[a][0], cv[a].y - acc[a][1], cv[a].z - acc[a][2], cv[a].w - acc[a][3]);
        }
      }
    }
    grid_barrier(w.ctl, target);
  }
  grid_barrier(w.ctl, target);
  // ---- commit: w_pchol = L, w_chol = L^-T (upper triangular), its transpose for the next GEMM 1, w_precision = P' ----
  float* st = p.state;
  const size_t n2 = (size_t)R * R, gsz = (size_t)gridDim.x * 256, g0 = (size_t)blockIdx.x * 256 + tid;
  for (size_t i = g0; i < n2; i += gsz) {
    const int r = (int)(i / R), c = (int)(i - (size_t)r * R);
    st[p.lay.w_pchol + i] = (c <= r) ? __ldcg(M + i) : 0.f;
    const float uv = (c >= r) ? __ldcg(M + (size_t)(R + d + r) * ld + c) : 0.f;
    st[p.lay.w_chol + i] = uv;
    w.Ut[(size_t)c * R + r] = uv;
    st[p.lay.w_precision + i] = w.Pnew[i];
  }
  // W'[c][k] = sum_{j >= c} w_chol[c][j] z[k][j]: one warp per output
  {
    const int gw = (int)((g0) >> 5), nw = (int)(gsz >> 5);
    for (int o = gw; o < R * d; o += nw) {
      const int c = o / d, k = o - c * d;
      const float* urow = M + (size_t)(R + d + c) * ld;
      const float* zrow = M + (size_t)(R + k) * ld;
      float s = 0.f;
      for (int j = c + lane; j < R; j += 32) s = fmaf(__ldcg(urow + j), __ldcg(zrow + j), s);
      s = warp_sum(s);
      if (lane == 0) st[p.lay.w_mean + o] = s;
    }
  }
}

// sum_b |dx - phi W|^2 (vjf/model.py:373-374): one warp per trial, W staged in shared memory, per-CTA partial sums in double
__global__ void __launch_bounds__(256) bigr_resid_kernel(const StepParams p, const BigR w) {
  extern __shared__ __align__(16) float Ws[];  // [R][d]
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, R = p.R, d = p.d;
  for (int i = threadIdx.x; i < R * d; i += 256) Ws[i] = __ldcg(p.state + p.lay.w_mean + i);
  __syncthreads();
  double tot = 0.0;
  for (int b = blockIdx.x * 8 + warp; b < p.B; b += gridDim.x * 8) {
    float acc[VJF_MAX_XDIM];
#pragma unroll
    for (int k = 0; k < VJF_MAX_XDIM; ++k) acc[k] = 0.f;
    const float* ph = w.phi + (size_t)b * R;
    for (int r0 = 0; r0 < R; r0 += 256) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int r = r0 + lane + 32 * i; f[i] = (r < R) ? ph[r] : 0.f; }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + lane + 32 * i;
        if (r < R) {
#pragma unroll
          for (int k = 0; k < VJF_MAX_XDIM; ++k)
            if (k < d) acc[k] = fmaf(f[i], Ws[r * d + k], acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < VJF_MAX_XDIM; ++k) {
      const float sum = warp_sum(acc[k]);
      if (k < d) { const float e = w.dx[(size_t)b * d + k] - sum; tot += (double)e * (double)e; }
    }
  }
  if (lane == 0) red[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) { double sum = 0.0; for (int i = 0; i < 8; ++i) sum += red[i]; w.rpart[blockIdx.x] = sum; }
}

// running state-noise variance (vjf/model.py:374-377, vjf/util.py:20-35); the partial sums are added in a fixed order
__global__ void bigr_noise_kernel(const StepParams p, const BigR w, int nparts) {
  const int lane = threadIdx.x;
  double tot = 0.0;
  for (int i = lane; i < nparts; i += 32) tot += w.rpart[i];
  tot = warp_sum_d(tot);
  if (lane != 0) return;
  float* st = p.state;
  const float mse = (float)(tot / ((double)p.Bglobal * (double)p.d));
  const double a = fmin((double)st[p.lay.tr_n], 500.0), n = a + (double)p.Bglobal;
  const float f1 = (float)(a / n), f2 = (float)((double)p.Bglobal / n);
  st[p.lay.tr_logvar] = logf(f1 * expf(st[p.lay.tr_logvar]) + f2 * mse);
  st[p.lay.tr_n] = (float)n;
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn bigr_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// [rows][cols] fp32 row-major with row stride ld (floats): boxes of {32 columns, box_rows rows}, 128-byte swizzle
static int map2d(CUtensorMap* m, const float* ptr, int rows, int cols, int ld, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = bigr_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { vjf_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -2; }
  return 0;
}

int vjf_bigr_create(vjf_handle* h) {
  const StepParams& p = h->base;
  if (!bigr_encode_fn()) { vjf_set_error("cuTensorMapEncodeTiled is not available: n_rbf > %d needs the TMA tensor maps", VJF_BIGR_MIN); return -1; }
  if (p.R % 4 != 0) { vjf_set_error("n_rbf > %d must be a multiple of 4 (16-byte row pitch of the TMA tensor maps), got %d", VJF_BIGR_MIN, p.R); return -1; }
  BigR* w = (BigR*)calloc(1, sizeof(BigR));
  h->bigr = w;
  const size_t B = (size_t)h->cfg.max_trials, R = (size_t)p.R, d = (size_t)p.d;
  w->Bmax = (int)B; w->Bp = (int)((B + 3) & ~(size_t)3) + 32;  // + 128 bytes: a power-of-two row pitch would put every row of a tile on the same L2 slice / channel
  w->NP = 2 * (int)((R + bg::BN - 1) / bg::BN);
  w->SK = 7;
  auto alloc = [&](float** ptr, size_t n) { if (cudaMalloc(ptr, n * sizeof(float)) != cudaSuccess) return -1; return cudaMemset(*ptr, 0, n * sizeof(float)) == cudaSuccess ? 0 : -1; };
  int bad = 0;
  bad |= alloc(&w->phi, B * R); bad |= alloc(&w->phiT, R * (size_t)w->Bp); bad |= alloc(&w->xs, B * d); bad |= alloc(&w->pm, B * d);
  bad |= alloc(&w->plv, B); bad |= alloc(&w->dx, B * d); bad |= alloc(&w->part1, (size_t)w->NP * B); bad |= alloc(&w->Ut, R * R);
  bad |= alloc(&w->Apart, (size_t)w->SK * R * R); bad |= alloc(&w->bstat, BIGR_BSPLIT * R * d); bad |= alloc(&w->M, (2 * R + d) * R); bad |= alloc(&w->Pnew, R * R);
  if (bad) { vjf_set_error("n_rbf=%d, max_trials=%d: out of device memory for the large-n_rbf workspace", p.R, h->cfg.max_trials); return -2; }
  VJF_CUDA_OK(cudaMalloc(&w->rpart, 1024 * sizeof(double)));
  VJF_CUDA_OK(cudaMalloc(&w->ctl, 4096));
  VJF_CUDA_OK(cudaMemset(w->ctl, 0, 4096));
  VJF_CUDA_OK(cudaFuncSetAttribute(bigr_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bg::SMEM_BYTES));
  VJF_CUDA_OK(cudaFuncSetAttribute(bigr_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  VJF_CUDA_OK(cudaFuncSetAttribute(bigr_resid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<size_t>(h->smem_limit - 1024, R * d * sizeof(float) + 64)));
  VJF_CUDA_OK(cudaFuncSetAttribute(bigr_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 64 * 65 * 4));
  const size_t feat = (BIGR_FT * (R + 1) + BIGR_FT * (size_t)(p.du + 3) + 8 + R * d) * 4;
  if (feat > h->smem_limit) { vjf_set_error("n_rbf=%d too large for the feature tile in shared memory", p.R); return -1; }
  return 0;
}

void vjf_bigr_destroy(vjf_handle* h) {
  BigR* w = h->bigr;
  if (!w) return;
  cudaFree(w->phi); cudaFree(w->phiT); cudaFree(w->xs); cudaFree(w->pm); cudaFree(w->plv); cudaFree(w->dx); cudaFree(w->part1); cudaFree(w->Ut);
  cudaFree(w->Apart); cudaFree(w->bstat); cudaFree(w->M); cudaFree(w->Pnew); cudaFree(w->rpart); cudaFree(w->ctl);
  free(w);
  h->bigr = nullptr;
}

// development aid: VJF_BIGR_TIMING=1 prints the CUDA-event time of every launch of the last time step to stderr
struct BigrTimer {
  bool on; cudaStream_t s; cudaEvent_t ev[32]; const char* name[32]; int n;
  void mark(const char* nm) { if (!on || n >= 32) return; if (!ev[n]) cudaEventCreate(&ev[n]); cudaEventRecord(ev[n], s); name[n++] = nm; }
  void report() {
    if (!on || n < 2) return;
    cudaEventSynchronize(ev[n - 1]);
    for (int i = 0; i + 1 < n; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); fprintf(stderr, "  bigr %-10s %8.1f us\n", name[i], ms * 1e3f); }
  }
};

int vjf_bigr_time_loop(vjf_handle* h, StepParams& p0, int T, int B, cudaStream_t s) {
  static BigrTimer tm = {getenv("VJF_BIGR_TIMING") != nullptr, nullptr, {}, {}, 0};
  tm.s = s;
  BigR& w = *h->bigr;
  if (B > w.Bmax) { vjf_set_error("trials B=%d above max_trials=%d", B, w.Bmax); return -1; }
  if (p0.y_dtype != VJF_Y_F32 && p0.y_dtype != VJF_Y_U8) { vjf_set_error("unknown y dtype"); return -1; }
  StepParams pl = p0;
  if (vjf_plan_tiles_public(h, pl, B)) return -1;
  const int R = pl.R, d = pl.d;
  pl.ext_xs = w.xs; pl.ext_pm = w.pm; pl.ext_plv = w.plv; pl.ext_dx = w.dx;
  pl.Bglobal = B; pl.T = 1;
  CUtensorMap mPhi, mUt, mPhiT_a, mPhiT_b;
  if (map2d(&mPhi, w.phi, B, R, R, bg::BM) || map2d(&mUt, w.Ut, R, R, R, bg::BN) || map2d(&mPhiT_a, w.phiT, R, B, w.Bp, bg::BM) ||
      map2d(&mPhiT_b, w.phiT, R, B, w.Bp, bg::BN)) return -2;
  // w_chol^T from the state (the caller may have loaded a new state since the last launch)
  bigr_transpose_kernel<<<dim3((R + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, s>>>(h->state + pl.lay.w_chol, w.Ut, R);
  ++g_vjf_launches;
  const size_t feat_smem = (BIGR_FT * ((size_t)R + 1) + BIGR_FT * (size_t)(pl.du + 3) + 8 + (size_t)R * d) * 4;
  const size_t ysz = (pl.y_dtype == VJF_Y_U8) ? 1 : 4;
  const bool upd = pl.flags & VJF_FLAG_UPDATE, warm = pl.flags & VJF_FLAG_WARMUP;
  const int nb_grid = std::max(1, std::min(h->num_sms, (pl.lay.n_train + VJF_NT - 1) / VJF_NT));
  for (int t = 0; t < T; ++t) {
    // per-step view of the launch parameters: step t of the trajectory as a one-step launch
    StepParams p = pl;
    p.y = reinterpret_cast<const unsigned char*>(pl.y) + (size_t)t * B * pl.D * ysz;
    p.u_in = pl.u_in ? pl.u_in + (size_t)t * B * pl.u : nullptr;
    p.eps = pl.eps ? pl.eps + (size_t)t * 2 * B * d : nullptr;
    p.mu = pl.mu + (size_t)t * B * d; p.logvar = pl.logvar + (size_t)t * B * d;
    p.losses = pl.losses ? pl.losses + (size_t)t * 4 : nullptr;
    if (t > 0) { p.q0m = pl.mu + (size_t)(t - 1) * B * d; p.q0l = pl.logvar + (size_t)(t - 1) * B * d; p.flags = pl.flags & ~(uint32_t)VJF_FLAG_PRIOR_Q0; }
    p.step0 = pl.step0 + t;
    tm.n = 0;
    tm.mark("feat");
    bigr_feat_kernel<<<(B + BIGR_FT - 1) / BIGR_FT, 512, feat_smem, s>>>(p, w, 0);
    tm.mark("gemm1");
    {
      bg::Args g = {B, R, R, 1, 0, 0, w.part1, w.Bmax};
      bigr_gemm_kernel<<<dim3((B + bg::BM - 1) / bg::BM, (R + bg::BN - 1) / bg::BN, 1), bg::NT, bg::SMEM_BYTES, s>>>(mPhi, mUt, g);
    }
    tm.mark("plv");
    bigr_plv_kernel<<<(B + 255) / 256, 256, 0, s>>>(w, B);
    tm.mark("phase_a");
    vjf_phase_a_kernel<<<p.nslots, VJF_NT, (size_t)p.s_total * sizeof(float), s>>>(p);
    g_vjf_launches += 4;
    tm.mark("reduce");
    if (vjf_internal_reduce(p, s)) return -2;
    tm.mark("phase_b");
    {
      StepParams pb = p;
      pb.B = B;
      vjf_phase_b_kernel<<<nb_grid, VJF_NT, (size_t)p.s_total * sizeof(float), s>>>(pb);
      ++g_vjf_launches;
    }
    if (upd) {
      if (!warm) {
        tm.mark("gemm2");
        bg::Args g = {R, R, B, 0, 1, 1, w.Apart, R};
        bigr_gemm_kernel<<<dim3((R + bg::BM - 1) / bg::BM, (R + bg::BN - 1) / bg::BN, w.SK), bg::NT, bg::SMEM_BYTES, s>>>(mPhiT_a, mPhiT_b, g);
        tm.mark("bstat");
        bigr_bstat_kernel<<<dim3((R + 7) / 8, BIGR_BSPLIT), 256, 0, s>>>(p, w);
        tm.mark("g+setup");
        bigr_g_kernel<<<(R + 7) / 8, 256, 0, s>>>(p, w);  // reads the old w_precision: before the commit of the factorisation
        bigr_setup_kernel<<<h->num_sms * 4, 256, 0, s>>>(p, w);
        tm.mark("factor");
        VJF_CUDA_OK(cudaMemsetAsync(w.ctl, 0, 8, s));
        void* args[] = {(void*)&p, (void*)&w};
        VJF_CUDA_OK(cudaLaunchCooperativeKernel((void*)bigr_factor_kernel, dim3(h->num_sms), dim3(256), args, (size_t)3 * 64 * 65 * 4, s));
        g_vjf_launches += 5;
      }
      const int nparts = std::min(4 * h->num_sms, std::max(1, (B + 7) / 8));
      tm.mark("resid");
      bigr_resid_kernel<<<nparts, 256, (size_t)R * d * sizeof(float), s>>>(p, w);
      bigr_noise_kernel<<<1, 32, 0, s>>>(p, w, nparts);
      g_vjf_launches += 2;
    }
    tm.mark("end");
  }
  tm.report();
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}

// views of the large-n_rbf workspace (tests / profiling): 0 phi [B][R], 1 p_mean [B][d], 2 p_logvar [B], 3 phi^T phi partial sums
extern "C" float* vjf_bigr_buffer(vjf_handle* h, int32_t which) {
  if (!h || !h->bigr) return nullptr;
  switch (which) { case 4: return reinterpret_cast<float*>(h->bigr->ctl); case 0: return h->bigr->phi; case 1: return h->bigr->pm; case 2: return h->bigr->plv; case 3: return h->bigr->Apart; default: return nullptr; }
}
