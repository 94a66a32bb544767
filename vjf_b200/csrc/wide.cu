// Wide observations (BASELINE.json configs[3], "C4": ydim = 2000 Poisson, xdim = 8, 65 536 trials over 2/4/8 GPUs): above the
// tile pipeline's limit (ydim <= 480) a trial's observation row no longer fits next to the rest of a tile in shared memory, and
// the recognition layer-1 weight (ydim x H, 1 MB) and its gradient no longer fit in shared / tensor memory.  The two
// contractions over the observation columns become GEMMs over ALL trials of the step, everything else of the step stays one
// trial-parallel kernel:
//
//   W1^T images    y-rows of the layer-1 weight, transposed (K-major B operand) as raw + lo = x - trunc(x)
//   GEMM fwd       pre_y[trial][n] = y_t[trial][:] W1[:D][n]                        (vjf/recognition.py:38, the y-columns of cat(y, u, q))
//   mid kernel     reparametrise, RBF features, dynamics read-out, h = tanh(pre_y + [u | m | l] W1[D:] + b), heads, decoder +
//                  likelihood + ELBO + hand-derived backward (vjf/model.py:97-154, :209), RLS statistics (vjf/module.py:94-96);
//                  hands g_pre^T (raw + lo) to the second GEMM
//   GEMM dW        dW1[:D]^T[n][j] = sum_trials g_pre[trial][n] y_t[trial][j]        (backward of recognition.py:38)
//   reduce         slot sums + split-K partial sums -> the reduced vector in the standard layout; (sharded: push all-reduce over
//                  NVLink peer memory) ; vjf_phase_b_kernel: clip + SGD, losses, running variances, RLS (k_split.cu)
//
// GEMM kernel: a pure TMA -> tcgen05 pipeline (no in-kernel operand pass): 128 x 128 output tile, fp32 accumulators in tensor
// memory, K in chunks of 32 through a ring of shared-memory stages; warp 0 issues the TMA tensor loads, warp 1 the
// tcgen05.mma kind::tf32 (one per operand-image pair: the tensor core truncates fp32 to tf32, so hi*hi + hi*lo + lo*hi with
// lo = x - trunc(x) gives fp32-grade products; spike counts are exact in tf32 and need no lo image), warps 2-5 read the
// accumulator back.  The forward GEMM reads y_t K-major (SWIZZLE_128B boxes); the weight-gradient GEMM contracts over the
// trials and reads the SAME y_t rows as an MN-major operand: TMA boxes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B land exactly
// the SWIZZLE_128B_BASE32B image the tensor core wants (profiles/micro_r02.txt cases 5.0 and 9.1) -- no transposed copy of y.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdio>

#include <cuda.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

struct Wide {
  int Bmax, Bp, Dp, ZF, ZD, nslots, per;
  float *W1T, *W1Tlo, *pre, *GT, *GTlo, *dWT, *ylo;
  unsigned long long* stamps;     // development aid (VJF_WIDE_STAMPS=1): globaltimer stamps of the last step, see scripts/c4_time.py
  unsigned* flag; unsigned seq;   // completion counter of the side-stream RLS launches (device word, host copy)
  cudaStream_t side; cudaEvent_t ev_sgd, ev_rls;  // host side: the serial half of phase B runs beside the next step's forward GEMM
};

// the two halves of phase B as separate launches (k_split.cu)
__global__ void __launch_bounds__(VJF_NT, 1) vjf_sgd_kernel(const __grid_constant__ StepParams p);
__global__ void __launch_bounds__(VJF_NT, 1) vjf_rls_kernel(const __grid_constant__ StepParams p);

namespace wg {
constexpr int BM = 128, BN = 128;
constexpr int IMG = BM * 128;            // bytes of one operand image of a stage: 128 rows x 32 floats
constexpr int NT = 192;                  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int MAXST = 6;

struct Args {
  int M, N, K;       // C (M x N) = A (M x K) B^T
  int b_mn;          // B is read MN-major from a [K][N] row-major matrix (boxes of 32 columns x 32 rows, ATOM_32B swizzle)
  int has_alo, has_blo;
  int dual;          // 0: one 128 x 128 tile; 1: two row tiles (two A images share a B image); 2: two column tiles (two B images share A)
  int stages;
  float* out;        // out[z][row][col], row stride ldo
  int ldo;
};

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
         (1ull << 46) | ((uint64_t)type << 61);
}
__device__ __forceinline__ uint64_t kmaj(uint32_t a) { return desc(a, 16, 1024, 2); }     // K-major, SWIZZLE_128B
__device__ __forceinline__ uint64_t mnmaj(uint32_t a) { return desc(a, 4096, 512, 1); }   // MN-major, SWIZZLE_128B_BASE32B, 32-column chunks of [32 rows][128 B]
__device__ __forceinline__ uint32_t idesc(int M, int N, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spins > (1u << 27)) __trap();  // a protocol error fails the launch instead of hanging the device
  }
}
__device__ __forceinline__ void tensor2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
}  // namespace wg

__global__ void __launch_bounds__(wg::NT, 1)
wide_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapBlo, const wg::Args g) {
  using namespace wg;
  extern __shared__ unsigned char smraw[];
  unsigned char* sb = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
  const int nimg = 2 + g.has_alo + g.has_blo + (g.dual == 1 ? 1 + g.has_alo : (g.dual == 2 ? 1 + g.has_blo : 0)), stage_bytes = nimg * IMG, ST = g.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + ST * stage_bytes);  // full[ST] empty[ST] accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAXST + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM * (g.dual == 1 ? 2 : 1), n0 = blockIdx.y * BN * (g.dual == 2 ? 2 : 1);
  const int NK = (g.K + 31) >> 5;
  const int per = (NK + gridDim.z - 1) / gridDim.z;
  const int kc0 = blockIdx.z * per, nk = max(0, min(NK, kc0 + per) - kc0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) { mbar_init(bars + s, 1); mbar_init(bars + MAXST + s, 1); }
    mbar_init(bars + 2 * MAXST, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc256(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // image order inside a stage: A | B | A lo | B lo | second tile: its A (| A lo) or its B (| B lo)
  const int o_b = IMG, o_alo = 2 * IMG, o_blo = (2 + g.has_alo) * IMG;
  const int o_2 = (2 + g.has_alo + g.has_blo) * IMG, o_2lo = o_2 + IMG, ntile = g.dual ? 2 : 1;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % ST, u = i / ST;
        if (u >= 1) wait(bars + MAXST + s, (u - 1) & 1);
        unsigned char* st = sb + s * stage_bytes;
        const int k = (kc0 + i) * 32;
        mbar_expect_tx(bars + s, (uint32_t)stage_bytes);
        tensor2d(st, &mapA, k, m0, bars + s);
        if (g.has_alo) tensor2d(st + o_alo, &mapAlo, k, m0, bars + s);
        if (g.dual == 1) {
          tensor2d(st + o_2, &mapA, k, m0 + BM, bars + s);
          if (g.has_alo) tensor2d(st + o_2lo, &mapAlo, k, m0 + BM, bars + s);
        }
        for (int tl = 0; tl < (g.dual == 2 ? 2 : 1); ++tl) {
          const int nn = n0 + tl * BN, ob = tl ? o_2 : o_b, obl = tl ? o_2lo : o_blo;
          if (g.b_mn) {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
              tensor2d(st + ob + c * 4096, &mapB, nn + 32 * c, k, bars + s);
              if (g.has_blo) tensor2d(st + obl + c * 4096, &mapBlo, nn + 32 * c, k, bars + s);
            }
          } else {
            tensor2d(st + ob, &mapB, k, nn, bars + s);
            if (g.has_blo) tensor2d(st + obl, &mapBlo, k, nn, bars + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t id = idesc(BM, BN, g.b_mn);
    for (int i = 0; i < nk; ++i) {
      const int s = i % ST, u = i / ST;
      wait(bars + s, u & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t base = smem_u32(sb + s * stage_bytes);
        for (int tl = 0; tl < ntile; ++tl) {
          const uint32_t a_hi = base + ((tl && g.dual == 1) ? o_2 : 0), a_lo = base + ((tl && g.dual == 1) ? o_2lo : o_alo);
          const uint32_t b_hi = base + ((tl && g.dual == 2) ? o_2 : o_b), b_lo = base + ((tl && g.dual == 2) ? o_2lo : o_blo);
          const uint32_t dcol = tmem + tl * BN;
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // four k-steps of 8: +32 bytes inside the K-major 128-byte rows, +8 rows of the MN-major image
            const uint32_t oa = j * 32, ob = g.b_mn ? j * 1024 : j * 32;
            uint32_t acc = (i > 0 || j > 0) ? 1u : 0u;
            if (g.has_alo) { umma_tf32_ss(dcol, kmaj(a_lo + oa), g.b_mn ? mnmaj(b_hi + ob) : kmaj(b_hi + ob), id, acc); acc = 1u; }
            if (g.has_blo) { umma_tf32_ss(dcol, kmaj(a_hi + oa), g.b_mn ? mnmaj(b_lo + ob) : kmaj(b_lo + ob), id, acc); acc = 1u; }
            umma_tf32_ss(dcol, kmaj(a_hi + oa), g.b_mn ? mnmaj(b_hi + ob) : kmaj(b_hi + ob), id, acc);
          }
        }
        umma_commit(bars + MAXST + s);
        if (i == nk - 1) umma_commit(bars + 2 * MAXST);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue: warp w reaches the tensor-memory lanes of quarter w % 4 ----
    if (nk > 0) { wait(bars + 2 * MAXST, 0); tc_fence_after(); }
    const int q = warp & 3;
    for (int tl = 0; tl < ntile; ++tl)
    for (int cb = 0; cb < BN / 32; ++cb) {
      const int row = m0 + (g.dual == 1 ? tl * BM : 0) + q * 32 + lane;
      const int col0 = (g.dual == 2 ? tl * BN : 0) + cb * 32;
      float v[32];
      if (nk > 0) tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + tl * BN + cb * 32, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (row < g.M && n0 + col0 < g.N) {
        float* o = g.out + ((size_t)blockIdx.z * g.M + row) * g.ldo + n0 + col0;
        if (n0 + col0 + 32 <= g.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32; ++j) if (n0 + col0 + j < g.N) o[j] = v[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_free256(tmem);
}

// y-rows of the recognition layer-1 weight [K1][H] (input-major) -> W1^T [H][Dp] raw and lo = x - trunc(x)
__global__ void wide_w1t_kernel(const float* __restrict__ w1, float* __restrict__ wt, float* __restrict__ wtlo, int D, int H, int Dp) {
  __shared__ float t[32][33];
  const int j0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = j0 + r, n = n0 + threadIdx.x;
    t[r][threadIdx.x] = (j < D && n < H) ? w1[(size_t)j * H + n] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, j = j0 + threadIdx.x;
    if (n < H && j < Dp) {
      const float v = t[threadIdx.x][r];
      wt[(size_t)n * Dp + j] = v;
      wtlo[(size_t)n * Dp + j] = v - tf32_trunc_f(v);
    }
  }
}

// lo image of one step's observations (only when they are not exact in tf32)
__global__ void wide_ylo_kernel(const float4* __restrict__ y, float4* __restrict__ lo, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = y[i];
    lo[i] = make_float4(x.x - tf32_trunc_f(x.x), x.y - tf32_trunc_f(x.y), x.z - tf32_trunc_f(x.z), x.w - tf32_trunc_f(x.w));
  }
}

// ------------------------------------------------------------------------------------------
// mid kernel: everything of the trial-parallel phase except the two contractions over the observation columns
// ------------------------------------------------------------------------------------------
constexpr int WM_TB = 32;   // trials per sub-tile
constexpr int WM_NC = 4;    // observation columns per thread (ydim <= 4 * 512)

struct WideSm {  // float offsets
  int hs, gp, phi, U, Wm, cen, iw, W1e, hm, hv, dec, ex, eps, xu, xt, mt, lt, pm, dx, gxt, gmt, glt, plv, gxp, red, xt8, gm8, gl8, total;
  int HP, RP, EP, U_in;
};

static inline WideSm wide_plan(const StepParams& p, bool u_in) {
  WideSm s;
  int f = 0;
  auto take = [&](int n) { int at = f; f = (f + n + 3) & ~3; return at; };
  const int d = p.d, H = p.H[0], R = p.R, TB = WM_TB;
  s.HP = H + 1; s.RP = R + 1; s.EP = std::max(4, (p.E + 3) & ~3); s.U_in = u_in ? 1 : 0;
  s.hs = take(TB * s.HP); s.gp = take(TB * s.HP); s.phi = take(TB * s.RP);
  s.U = take(u_in ? R * R : 4);
  s.Wm = take(R * d); s.cen = take(R * p.du); s.iw = take(R);
  s.W1e = take((p.E + 1) * H);
  s.hm = take(H * d); s.hv = take(H * d + d);
  s.dec = take(12 * p.D);  // per observation column: 8 decoder weights (zero padded), bias, 3 pad
  s.ex = take(TB * s.EP); s.eps = take(TB * 2 * d); s.xu = take(TB * p.du);
  s.xt = take(TB * d); s.mt = take(TB * d); s.lt = take(TB * d); s.pm = take(TB * d); s.dx = take(TB * d);
  s.gxt = take(TB * d); s.gmt = take(TB * d); s.glt = take(TB * d); s.plv = take(TB);
  s.gxp = take(VJF_NWARP * TB * 8);
  s.red = take(VJF_NWARP * VJF_NSCAL + 16);
  s.xt8 = take(TB * 8); s.gm8 = take(TB * 8); s.gl8 = take(TB * 8);  // rows padded to 8: 128-bit broadcast loads
  s.total = f;
  return s;
}

__device__ __forceinline__ void wm_acc(float* p, float v, bool first) { *p = first ? v : *p + v; }  // a CTA's slot is its own

template <int LIK>
__global__ void __launch_bounds__(VJF_NT, 1) wide_mid_kernel(const StepParams p, const Wide w, const WideSm s, const unsigned wait_seq) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D, d = p.d, u = p.u, R = p.R, du = p.du, E = p.E, H = p.H[0], HP = s.HP, RP = s.RP, EP = s.EP, B = p.B;
  constexpr int TB = WM_TB;
  float* hs = sm + s.hs;   float* gp = sm + s.gp;   float* phi = sm + s.phi; float* U_s = sm + s.U;  float* Wm = sm + s.Wm;
  float* cen = sm + s.cen; float* iw = sm + s.iw;   float* W1e = sm + s.W1e; float* hm = sm + s.hm;  float* hv = sm + s.hv;
  float* dec = sm + s.dec; float* ex = sm + s.ex;   float* eps_s = sm + s.eps; float* xu = sm + s.xu; float* xt_s = sm + s.xt;
  float* mt_s = sm + s.mt; float* lt_s = sm + s.lt; float* pm_s = sm + s.pm; float* dx_s = sm + s.dx; float* gxt_s = sm + s.gxt;
  float* gmt_s = sm + s.gmt; float* glt_s = sm + s.glt; float* plv_s = sm + s.plv; float* gxp = sm + s.gxp; float* red_s = sm + s.red;
  float* xt8 = sm + s.xt8; float* gm8 = sm + s.gm8; float* gl8 = sm + s.gl8;
  float* st = p.state;
  float* slot = p.partials + (size_t)blockIdx.x * p.PS;
  const bool r_on = true, d_on = !(p.flags & VJF_FLAG_WARMUP), h_on = true;

  // ---- parameters of this step into shared memory ----
  for (int i = tid; i < R * du; i += VJF_NT) cen[i] = st[p.lay.centroid + i];
  for (int i = tid; i < R; i += VJF_NT) { const float wd = expf(st[p.lay.logwidth + i]); iw[i] = -0.5f / (wd * wd); }
  for (int i = tid; i < E * H; i += VJF_NT) W1e[i] = st[p.lay.mlp_w[0] + (size_t)D * H + i];
  for (int i = tid; i < H; i += VJF_NT) W1e[E * H + i] = st[p.lay.mlp_b[0] + i];
  for (int i = tid; i < H * d; i += VJF_NT) { hm[i] = st[p.lay.head_m_w + i]; hv[i] = st[p.lay.head_v_w + i]; }
  for (int i = tid; i < d; i += VJF_NT) hv[H * d + i] = st[p.lay.head_v_b + i];
  for (int k = 0; k < 8; ++k)
    for (int j = tid; j < D; j += VJF_NT) dec[j * 12 + k] = (k < d) ? st[p.lay.dec_w + k * D + j] : 0.f;
  for (int i = tid; i < D; i += VJF_NT) dec[i * 12 + 8] = st[p.lay.dec_b + i];
  for (int i = tid; i < TB * 8; i += VJF_NT) { xt8[i] = 0.f; gm8[i] = 0.f; gl8[i] = 0.f; }
  for (int i = tid; i < TB * EP; i += VJF_NT) ex[i] = 0.f;
  // What the serial half of the previous step (vjf_rls_kernel, side stream) writes -- w_chol, w_mean, the noise variances -- is
  // read only after its completion counter has reached wait_seq: with spike counts (no observation-noise parameter) that is
  // after the likelihood stage of the first sub-tile, so the RLS hides behind the forward GEMM and most of this kernel's work.
  // The grid leaves one SM free (host), so the RLS launch can always be scheduled; a lost launch times out instead of hanging.
  if (w.stamps && blockIdx.x == 0 && tid == 0) w.stamps[(p.step0 & 7) * 8 + 2] = (unsigned long long)gtime_ns();
  bool rls_ready = false;
  float lam = 0.f, e_nlam = 1.f, p_lam = 1.f, gam = 0.f, e_ngam = 1.f, p_gam = 1.f;
  auto wait_rls = [&]() {
    if (w.stamps && blockIdx.x == 0 && tid == 0) w.stamps[(p.step0 & 7) * 8 + 3] = (unsigned long long)gtime_ns();
    if (tid == 0 && wait_seq) {
      const long long t0 = clock64();
      while (ld_acquire_u32(w.flag) < wait_seq) {
        if (clock64() - t0 > 4000000000ll) { atomicOr(p.status, (unsigned)VJF_ST_COMM_TIMEOUT); break; }
        __nanosleep(64);
      }
      __threadfence();
    }
    if (w.stamps && blockIdx.x == 0 && tid == 0) w.stamps[(p.step0 & 7) * 8 + 4] = (unsigned long long)gtime_ns();
    __syncthreads();
    if (s.U_in) for (int i = tid; i < R * R; i += VJF_NT) U_s[i] = __ldcg(st + p.lay.w_chol + i);
    for (int i = tid; i < R * d; i += VJF_NT) Wm[i] = __ldcg(st + p.lay.w_mean + i);
    if (LIK == VJF_LIK_GAUSSIAN) lam = __ldcg(st + p.lay.lik_logvar);
    e_nlam = expf(-lam); p_lam = expf(-0.5f * lam);
    gam = __ldcg(st + p.lay.tr_logvar);
    e_ngam = expf(-gam); p_gam = expf(-0.5f * gam);
    rls_ready = true;
    __syncthreads();
  };
  __syncthreads();
  if (LIK == VJF_LIK_GAUSSIAN) wait_rls();  // the decoder stage needs the observation-noise variance of the previous step

  float acc_w[WM_NC][8], acc_b[WM_NC], sc[VJF_NSCAL];
#pragma unroll
  for (int c = 0; c < WM_NC; ++c) {
    acc_b[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc_w[c][k] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < VJF_NSCAL; ++i) sc[i] = 0.f;

  const int c_lo = blockIdx.x * w.per, c_hi = min(B, c_lo + w.per);
  const bool prior = p.flags & VJF_FLAG_PRIOR_Q0;
  const float* yf = reinterpret_cast<const float*>(p.y);
  bool first = true;
  for (int b0 = c_lo; b0 < c_hi; b0 += TB) {
    const int nb = min(TB, c_hi - b0);
    // ---- S0: previous posterior, control input, noise (vjf/util.py:11-13, 38-49) ----
    for (int i = tid; i < TB * d; i += VJF_NT) {
      const int b = i / d, k = i - b * d;
      float ms = 0.f, ls = 0.f;
      if (b < nb) {
        if (prior) { ms = st[p.lay.prior_mean + k]; ls = st[p.lay.prior_logvar + k]; }
        else { ms = p.q0m[(size_t)(b0 + b) * d + k]; ls = p.q0l[(size_t)(b0 + b) * d + k]; }
        if (p.eps) { eps_s[b * 2 * d + k] = p.eps[(size_t)(b0 + b) * d + k]; eps_s[b * 2 * d + d + k] = p.eps[((size_t)B + b0 + b) * d + k]; }
      } else { eps_s[b * 2 * d + k] = 0.f; eps_s[b * 2 * d + d + k] = 0.f; }
      ex[b * EP + u + k] = ms; ex[b * EP + u + d + k] = ls;
    }
    if (!p.eps) {
      const int nblk = (d + 3) >> 2;
      for (int i = tid; i < nb * 2 * nblk; i += VJF_NT) {
        const int b = i / (2 * nblk), r = i - b * 2 * nblk, which = r / nblk, blk = r - which * nblk;
        float z[4];
        philox_normal4(p.seed, p.step0, p.trial_offset + b0 + b, which, blk, z);
        for (int k = 0; k < 4; ++k)
          if (blk * 4 + k < d) eps_s[b * 2 * d + which * d + blk * 4 + k] = z[k];
      }
    }
    for (int i = tid; i < TB * u; i += VJF_NT) {
      const int b = i / u, e = i - b * u;
      ex[b * EP + e] = (b < nb) ? p.u_in[(size_t)(b0 + b) * u + e] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < TB * du; i += VJF_NT) {
      const int b = i / du, k = i - b * du;
      xu[i] = (k < d) ? ex[b * EP + u + k] + eps_s[b * 2 * d + k] * expf(0.5f * ex[b * EP + u + d + k]) : ex[b * EP + (k - d)];
    }
    __syncthreads();
    // ---- S1: RBF features (vjf/functional.py:11-22) ; S2: h = tanh(in W1 + b1) (vjf/recognition.py:38-40) with the observation part
    //      of the contraction from the forward GEMM ----
    for (int i = tid; i < TB * R; i += VJF_NT) {
      const int b = i / R, r = i - b * R;
      float v = 0.f;
      if (b < nb) {
        float d2 = 0.f;
        for (int c = 0; c < du; ++c) { const float df = xu[b * du + c] - cen[r * du + c]; d2 = fmaf(df, df, d2); }
        v = expf(d2 * iw[r]);
      }
      phi[b * RP + r] = v;
    }
    for (int i = tid; i < TB * H; i += VJF_NT) {
      const int b = i / H, n = i - b * H;
      float v = 0.f;
      if (b < nb) {
        float a = W1e[E * H + n];
        for (int z = 0; z < w.ZF; ++z) a += w.pre[((size_t)z * B + b0 + b) * H + n];
        for (int e = 0; e < E; ++e) a = fmaf(ex[b * EP + e], W1e[e * H + n], a);
        v = tanhf(a);
      }
      hs[b * HP + n] = v;
    }
    __syncthreads();
    // ---- S3: heads (recognition.py:41-42), xt, dx, posterior out ----
    for (int i0 = 0; i0 < TB * d * 8; i0 += VJF_NT) {
      const int i = i0 + tid, nq = i & 7, bk = i >> 3;
      const bool ok = bk < TB * d;
      const int b = ok ? bk / d : 0, k = ok ? bk - b * d : 0;
      float m = 0.f, lv = 0.f;
      if (ok && b < nb) {
        const float* hr = hs + b * HP;
        for (int n = nq; n < H; n += 8) { const float h = hr[n]; m = fmaf(h, hm[n * d + k], m); lv = fmaf(h, hv[n * d + k], lv); }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) { m += __shfl_xor_sync(0xffffffffu, m, o); lv += __shfl_xor_sync(0xffffffffu, lv, o); }
      if (ok && nq == 0) {
        float x = 0.f, dxv = 0.f;
        if (b < nb) {
          lv += hv[H * d + k];
          x = m + eps_s[b * 2 * d + d + k] * expf(0.5f * lv);
          dxv = x - xu[b * du + k];
          sc[SC_SDX] = fmaf(dxv, dxv, sc[SC_SDX]);
          p.mu[(size_t)(b0 + b) * d + k] = m;
          p.logvar[(size_t)(b0 + b) * d + k] = lv;
        } else { m = 0.f; lv = 0.f; }
        mt_s[bk] = m; lt_s[bk] = lv; xt_s[bk] = x; dx_s[bk] = dxv; xt8[b * 8 + k] = x;
      }
    }
    __syncthreads();  // (xt8, mt_s ... of S3 visible)
    // ---- S5: decoder eta = D xt + bias (model.py:29-30), likelihood, d loss / d eta (times B), decoder gradients, g_xt.
    //      A thread owns the observation columns tid + 512 c and keeps their decoder gradients in registers across all of the
    //      CTA's trials; two trials are in flight, their g_xt partial sums (16 values) meet in one transposed warp reduction ----
    {
      const float* ybase = yf + (size_t)b0 * D;
      float ycur[2][WM_NC];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int c = 0; c < WM_NC; ++c) { const int j = tid + VJF_NT * c; ycur[q][c] = (q < nb && j < D) ? ybase[(size_t)q * D + j] : 0.f; }
#pragma unroll 1
      for (int bb = 0; bb < nb; bb += 2) {
        float ynext[2][WM_NC];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int c = 0; c < WM_NC; ++c) { const int j = tid + VJF_NT * c; ynext[q][c] = (bb + 2 + q < nb && j < D) ? ybase[(size_t)(bb + 2 + q) * D + j] : 0.f; }
        float xt[2][8], gx[2][8], onf[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 xa = *reinterpret_cast<const float4*>(xt8 + min(bb + q, TB - 1) * 8), xb = *reinterpret_cast<const float4*>(xt8 + min(bb + q, TB - 1) * 8 + 4);
          xt[q][0] = xa.x; xt[q][1] = xa.y; xt[q][2] = xa.z; xt[q][3] = xa.w; xt[q][4] = xb.x; xt[q][5] = xb.y; xt[q][6] = xb.z; xt[q][7] = xb.w;
          onf[q] = (bb + q < nb) ? 1.0f : 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) gx[q][k] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < WM_NC; ++c) {
          const int j = tid + VJF_NT * c;
          if (j < D) {
            const float4 wa = *reinterpret_cast<const float4*>(dec + j * 12), wb = *reinterpret_cast<const float4*>(dec + j * 12 + 4);
            const float wk[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            const float bj = dec[j * 12 + 8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float eta = bj;
#pragma unroll
              for (int k = 0; k < 8; ++k) eta = fmaf(wk[k], xt[q][k], eta);
              const float yv = ycur[q][c];
              float gv;
              if (LIK == VJF_LIK_GAUSSIAN) {
                // gaussian_loss(y, eta, lambda), functional.py:55-75 ; update's mse, likelihood.py:36-37
                const float r = yv - eta;
                const float rsd = yv * p_lam - eta * p_lam;
                const float mse = rsd * rsd;
                sc[SC_BADMSE] += (onf[q] != 0.f && !isfinite(mse)) ? 1.f : 0.f;
                sc[SC_RECON] = fmaf(onf[q], 0.5f * (mse + lam), sc[SC_RECON]);
                sc[SC_SSE] = fmaf(onf[q] * r, r, sc[SC_SSE]);
                gv = -r * e_nlam * onf[q];
                sc[6] = fmaf(onf[q], 0.5f * (1.0f - r * r * e_nlam), sc[6]);
              } else {
                // poisson_nll_loss(clamp(eta, max=10), y, log_input=True), likelihood.py:60-62; a NaN eta stays NaN (torch.clamp):
                // the comparison is false for it, so it flows through exp into the loss and the gradient
                const bool big = eta > 10.0f;
                const float ec = big ? 10.0f : eta;
                const float exv = __expf(ec);
                sc[SC_RECON] = fmaf(onf[q], fmaf(-yv, ec, exv), sc[SC_RECON]);
                gv = (big ? 0.f : exv - yv) * onf[q];
              }
              acc_b[c] += gv;
#pragma unroll
              for (int k = 0; k < 8; ++k) { acc_w[c][k] = fmaf(gv, xt[q][k], acc_w[c][k]); gx[q][k] = fmaf(gv, wk[k], gx[q][k]); }
            }
          }
        }
        // transposed reduction: 16 values over 32 lanes in 8 + 4 + 2 + 1 + 1 shuffles; lane L ends with the sum of value L >> 1
        float v8[8], v4[4], v2[2], v1;
        {
          const bool up = lane & 16;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float keep = up ? gx[1][i] : gx[0][i], send = up ? gx[0][i] : gx[1][i];
            v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool up = lane & 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float keep = up ? v8[i + 4] : v8[i], send = up ? v8[i] : v8[i + 4];
            v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        {
          const bool up = lane & 4;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float keep = up ? v4[i + 2] : v4[i], send = up ? v4[i] : v4[i + 2];
            v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
        {
          const bool up = lane & 2;
          const float keep = up ? v2[1] : v2[0], send = up ? v2[0] : v2[1];
          v1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        if ((lane & 1) == 0) {
          const int idx = lane >> 1, q = idx >> 3, k = idx & 7;  // value index: bit 3 = trial of the pair, bits 0-2 = state dimension
          if (bb + q < TB) gxp[(warp * TB + bb + q) * 8 + k] = v1;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int c = 0; c < WM_NC; ++c) ycur[q][c] = ynext[q][c];
      }
    }
    __syncthreads();
    for (int i = tid; i < TB * d; i += VJF_NT) {  // g_xt: the warps' partial sums in warp order
      const int b = i / d, k = i - b * d;
      float gxv = 0.f;
      for (int wv = 0; wv < VJF_NWARP; ++wv) gxv += gxp[(wv * TB + b) * 8 + k];
      gxt_s[i] = gxv;
    }
    if (!rls_ready) wait_rls();  // (first sub-tile: the RLS outputs of the previous step)
    else __syncthreads();
    // ---- S4: dynamics read-out (vjf/module.py:75-77, model.py:338): p_mean = xs + phi W, p_logvar = log |phi w_chol|^2 ----
    {  // FL[b][n] = sum_r phi[b][r] w_chol[r][n]: a thread owns one column n for four trials (one load of w_chol per four FMAs)
      const float* Ug = s.U_in ? U_s : st + p.lay.w_chol;
      for (int it = tid; it < R * (TB / 4); it += VJF_NT) {
        const int bg = it / R, n = it - bg * R;
        const float* ph = phi + (4 * bg) * RP;
        float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
          const float uv = Ug[r * R + n];
          f0 = fmaf(ph[r], uv, f0); f1 = fmaf(ph[RP + r], uv, f1); f2 = fmaf(ph[2 * RP + r], uv, f2); f3 = fmaf(ph[3 * RP + r], uv, f3);
        }
        float* o = gxp + (4 * bg) * 128 + n;  // (scratch: the g_xt partial sums are not live yet)
        o[0] = f0 * f0; o[128] = f1 * f1; o[256] = f2 * f2; o[384] = f3 * f3;
      }
    }
    __syncthreads();
    for (int b = warp; b < TB; b += VJF_NWARP) {
      float q = 0.f;
      for (int n = lane; n < R; n += 32) q += gxp[b * 128 + n];
      q = warp_sum(q);
      if (lane == 0) plv_s[b] = (b < nb) ? logf(q) : 0.f;
    }
    __syncthreads();  // (xt_s, mt_s ... of S3 visible)
    for (int i = tid; i < TB * d; i += VJF_NT) {
      const int b = i / d, k = i - b * d;
      float a = 0.f;
      if (b < nb) {
        const float* ph = phi + b * RP;
        float a0 = 0.f, a1 = 0.f;
        int r = 0;
        for (; r + 1 < R; r += 2) { a0 = fmaf(ph[r], Wm[r * d + k], a0); a1 = fmaf(ph[r + 1], Wm[(r + 1) * d + k], a1); }
        if (r < R) a0 = fmaf(ph[r], Wm[r * d + k], a0);
        a = xu[b * du + k] + (a0 + a1);
      }
      pm_s[i] = a;
    }
    __syncthreads();
    // ---- S6: dynamics NLL (functional.py:55-75 via model.py:390-391), entropy (functional.py:25-29), g_mt and g_lt (times B) ----
    for (int i = tid; i < TB * d; i += VJF_NT) {
      const int b = i / d, k = i - b * d;
      float gm = 0.f, gl = 0.f;
      if (b < nb) {
        const float gxv = gxt_s[i];
        const float m = mt_s[i], lv = lt_s[i], pm = pm_s[i], plv = plv_s[b];
        const float e2 = eps_s[b * 2 * d + d + k];
        const float df = pm * p_gam - m * p_gam;
        const float mse = df * df;
        if (!isfinite(mse)) sc[SC_BADMSE] += 1.f;
        const float tr = expf(plv + lv - gam);
        sc[SC_DYN] += 0.5f * (mse + gam) + 0.5f * tr;
        sc[SC_ENT] += 0.5f * lv;
        gm = gxv; gl = 0.5f * gxv * e2 * expf(0.5f * lv);
        if (h_on) gl -= 0.5f;
        if (d_on) { gm += (m - pm) * e_ngam; gl += 0.5f * tr; }
      }
      gmt_s[i] = gm; glt_s[i] = gl; gm8[b * 8 + k] = gm; gl8[b * 8 + k] = gl;
    }
    __syncthreads();
    // ---- S7: g_pre = (g_mt W_m + g_lt W_v)(1 - h^2), kept for the small gradients and handed to the weight-gradient GEMM as
    //      g_pre^T (raw, lo): lanes over the trials, so that the global rows are written 128 bytes at a time ----
    {
      const int b = tid & 31;
      float gm[8], gl[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { gm[k] = (k < d) ? gmt_s[b * d + k] : 0.f; gl[k] = (k < d) ? glt_s[b * d + k] : 0.f; }
      for (int n = tid >> 5; n < H; n += VJF_NWARP) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < d) { a = fmaf(gm[k], hm[n * d + k], a); a = fmaf(gl[k], hv[n * d + k], a); }
        const float h = hs[b * HP + n];
        const float v = (b < nb) ? a * (1.0f - h * h) : 0.f;
        gp[b * HP + n] = v;
        if (b < nb) {
          w.GT[(size_t)n * w.Bp + b0 + b] = v;
          w.GTlo[(size_t)n * w.Bp + b0 + b] = v - tf32_trunc_f(v);
        }
      }
    }
    __syncthreads();
    // ---- S8: the small gradients and the RLS statistics of this sub-tile into the CTA's slot ----
    if (tid < d) {
      float a = 0.f;
      for (int b = 0; b < nb; ++b) a += glt_s[b * d + tid];
      wm_acc(slot + p.lay.head_v_b + tid, a, first);
    }
    {
      // one list of register-blocked work items over all threads: phi^T phi in 4 x 2 blocks | layer-1 weight gradient rows of
      // [u | m | l] (four inputs per item) and the bias | head weight gradients (a hidden unit x all state dimensions)
      const int ncb = (R + 1) >> 1, nA = ((R + 3) >> 2) * ncb;
      const int nch = EP >> 2, nW = H * (nch + 1), nHd = 2 * H;
      for (int it = tid; it < nA + nW + nHd; it += VJF_NT) {
        if (it < nA) {
          const int r0 = (it / ncb) * 4, c0 = (it % ncb) * 2;
          const int r1 = min(r0 + 1, R - 1), r2 = min(r0 + 2, R - 1), r3 = min(r0 + 3, R - 1), c1 = min(c0 + 1, R - 1);
          float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, a20 = 0.f, a21 = 0.f, a30 = 0.f, a31 = 0.f;
#pragma unroll 4
          for (int b = 0; b < nb; ++b) {
            const float* ph = phi + b * RP;
            const float x0 = ph[r0], x1 = ph[r1], x2 = ph[r2], x3 = ph[r3], y0 = ph[c0], y1 = ph[c1];
            a00 = fmaf(x0, y0, a00); a01 = fmaf(x0, y1, a01); a10 = fmaf(x1, y0, a10); a11 = fmaf(x1, y1, a11);
            a20 = fmaf(x2, y0, a20); a21 = fmaf(x2, y1, a21); a30 = fmaf(x3, y0, a30); a31 = fmaf(x3, y1, a31);
          }
          float* o = slot + p.pa;
          const bool c1ok = c0 + 1 < R;
          wm_acc(o + r0 * R + c0, a00, first); if (c1ok) wm_acc(o + r0 * R + c0 + 1, a01, first);
          if (r0 + 1 < R) { wm_acc(o + (r0 + 1) * R + c0, a10, first); if (c1ok) wm_acc(o + (r0 + 1) * R + c0 + 1, a11, first); }
          if (r0 + 2 < R) { wm_acc(o + (r0 + 2) * R + c0, a20, first); if (c1ok) wm_acc(o + (r0 + 2) * R + c0 + 1, a21, first); }
          if (r0 + 3 < R) { wm_acc(o + (r0 + 3) * R + c0, a30, first); if (c1ok) wm_acc(o + (r0 + 3) * R + c0 + 1, a31, first); }
        } else if (it < nA + nW) {
          const int i2 = it - nA, ch = i2 / H, n = i2 - ch * H;
          if (ch < nch) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
            for (int b = 0; b < nb; ++b) {
              const float g = gp[b * HP + n];
              const float4 x = *reinterpret_cast<const float4*>(ex + b * EP + 4 * ch);
              a0 = fmaf(g, x.x, a0); a1 = fmaf(g, x.y, a1); a2 = fmaf(g, x.z, a2); a3 = fmaf(g, x.w, a3);
            }
            float* o = slot + p.lay.mlp_w[0] + (size_t)(D + 4 * ch) * H + n;
            if (4 * ch < E) wm_acc(o, a0, first);
            if (4 * ch + 1 < E) wm_acc(o + H, a1, first);
            if (4 * ch + 2 < E) wm_acc(o + 2 * H, a2, first);
            if (4 * ch + 3 < E) wm_acc(o + 3 * H, a3, first);
          } else {
            float a = 0.f;
            for (int b = 0; b < nb; ++b) a += gp[b * HP + n];
            wm_acc(slot + p.lay.mlp_b[0] + n, a, first);
          }
        } else {
          const int i2 = it - nA - nW, which = i2 / H, n = i2 - which * H;
          const float* g8 = which ? gl8 : gm8;
          float a[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) a[k] = 0.f;
#pragma unroll 2
          for (int b = 0; b < nb; ++b) {
            const float h = hs[b * HP + n];
            const float4 ga = *reinterpret_cast<const float4*>(g8 + b * 8), gb = *reinterpret_cast<const float4*>(g8 + b * 8 + 4);
            a[0] = fmaf(h, ga.x, a[0]); a[1] = fmaf(h, ga.y, a[1]); a[2] = fmaf(h, ga.z, a[2]); a[3] = fmaf(h, ga.w, a[3]);
            a[4] = fmaf(h, gb.x, a[4]); a[5] = fmaf(h, gb.y, a[5]); a[6] = fmaf(h, gb.z, a[6]); a[7] = fmaf(h, gb.w, a[7]);
          }
          float* o = slot + (which ? p.lay.head_v_w : p.lay.head_m_w) + n * d;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < d) wm_acc(o + k, a[k], first);
        }
      }
    }
    for (int i = tid; i < R * d; i += VJF_NT) {
      const int r = i / d, k = i - r * d;
      float a = 0.f;
      for (int b = 0; b < nb; ++b) a = fmaf(phi[b * RP + r], dx_s[b * d + k], a);
      wm_acc(slot + p.pb + i, a, first);
    }
    first = false;
    __syncthreads();
  }

  if (w.stamps && blockIdx.x == 0 && tid == 0) w.stamps[(p.step0 & 7) * 8 + 5] = (unsigned long long)gtime_ns();
  if (w.stamps && tid == 0) atomicMax(w.stamps + (p.step0 & 7) * 8 + 7, (unsigned long long)gtime_ns());  // last CTA to finish
  // ---- flush: decoder gradients (registers) and the scalar sums ----
#pragma unroll
  for (int c = 0; c < WM_NC; ++c) {
    const int j = tid + VJF_NT * c;
    if (j < D) {
      slot[p.lay.dec_b + j] = acc_b[c];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < d) slot[p.lay.dec_w + k * D + j] = acc_w[c][k];
    }
  }
#pragma unroll
  for (int i = 0; i < VJF_NSCAL; ++i) {
    const float v = warp_sum(sc[i]);
    if (lane == 0) red_s[warp * VJF_NSCAL + i] = v;
  }
  __syncthreads();
  if (tid < VJF_NSCAL) {
    float v = 0.f;
    for (int wv = 0; wv < VJF_NWARP; ++wv) v += red_s[wv * VJF_NSCAL + tid];
    if (tid == 6) slot[p.lay.lik_logvar] = v;  // Gaussian d loss / d lambda (times B)
    else slot[p.ps + tid] = v;
  }
}

// completion counter of the side-stream launches (the kernel boundary orders the RLS kernel's stores before it)
__global__ void wide_flag_kernel(unsigned* flag, unsigned value, unsigned long long* stamp) {
  if (stamp) *stamp = (unsigned long long)gtime_ns();
  __threadfence();
  st_release_gpu_u32(flag, value);
}
__global__ void wide_stamp_kernel(unsigned long long* stamp) { *stamp = (unsigned long long)gtime_ns(); }

// The reduced vector in the standard layout, one launch of 1024-thread blocks, every sum in a fixed order:
//   blocks [0, nW1): reduced[y-rows of dW1][j][n] = sum_z dWT[z][n][j] -- a 32 x 32 tile per block, transposed through shared memory;
//                    a warp owns a row n of the tile and has all split-K partial sums of its 128-byte segment in flight
//   the others:      everything else = sums over the mid kernel's slots; a block owns 32 consecutive elements, warp w adds the
//                    slots w, w + 32, ... (independent row loads), the 32 partial sums meet in warp order
// Sharded run (push > 0): the sums are stored straight into the inbox slot (this rank, parity) of EVERY rank's peer-mapped exchange
// buffer -- posted NVLink writes that leave while the other ranks are still computing; `out` is then unused.
__global__ void __launch_bounds__(1024) wide_reduce_kernel(const StepParams p, const Wide w, float* __restrict__ out, int ntj, int nW1, int push, int par) {
  __shared__ float t[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, H = p.H[0], D = p.D;
  if ((int)blockIdx.x < nW1) {
    const int j0 = ((int)blockIdx.x % ntj) * 32, n0 = ((int)blockIdx.x / ntj) * 32;
    const int n = n0 + warp, j = j0 + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (n < H && j < D) {
      const float* src = w.dWT + (size_t)n * w.Dp + j;
      const size_t zs = (size_t)H * w.Dp;
      int z = 0;
      for (; z + 3 < w.ZD; z += 4) { a0 += src[z * zs]; a1 += src[(z + 1) * zs]; a2 += src[(z + 2) * zs]; a3 += src[(z + 3) * zs]; }
      for (; z < w.ZD; ++z) a0 += src[z * zs];
    }
    t[warp][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    const int jo = j0 + warp, no = n0 + lane;
    if (jo < D && no < H) {
      const size_t e = p.lay.mlp_w[0] + (size_t)jo * H + no;
      const float v = t[lane][warp];
      if (push) { for (int r = 0; r < p.world; ++r) p.peer[r][(size_t)(p.rank * 2 + par) * p.PSx + e] = v; }
      else out[e] = v;
    }
    return;
  }
  const int w0 = p.lay.mlp_w[0], w1 = w0 + D * H;  // (the y-rows of dW1 come from the GEMM)
  const int n_rest = p.PS - (w1 - w0);
  const int i = ((int)blockIdx.x - nW1) * 32 + lane;
  const int ic = min(i, n_rest - 1);
  const int e = (ic < w0) ? ic : ic + (w1 - w0);
  const float* src = p.partials + e;
  float a0 = 0.f, a1 = 0.f;
  int c = warp;
  for (; c + 32 < w.nslots; c += 64) { a0 += src[(size_t)c * p.PS]; a1 += src[(size_t)(c + 32) * p.PS]; }
  if (c < w.nslots) a0 += src[(size_t)c * p.PS];
  t[warp][lane] = a0 + a1;
  __syncthreads();
  if (warp == 0 && i < n_rest) {
    float v = t[0][lane];
#pragma unroll
    for (int k = 1; k < 32; ++k) v += t[k][lane];
    if (push) { for (int r = 0; r < p.world; ++r) p.peer[r][(size_t)(p.rank * 2 + par) * p.PSx + e] = v; }
    else out[e] = v;
  }
}

// Sharded run: all-reduce of the reduced vector over NVLink peer memory, push model.  The reduce launch of every rank has
// stored its sums into slot (rank, parity) of every rank's exchange buffer (the kernel boundary completes those writes at
// system scope); block 0 raises this rank's flag on every peer, every block waits until all ranks have raised theirs here,
// then the vector is the sum of the local inbox slots in rank order -- identical on every rank.
__global__ void __launch_bounds__(256) wide_exchange_kernel(const StepParams p, unsigned epoch) {
  const int par = epoch & 1, nchx = p.PSx >> 7;
  const size_t flag_off = (size_t)p.world * 2 * p.PSx;
  __shared__ int dead;
  if (threadIdx.x == 0) dead = 0;
  __syncthreads();
  if (threadIdx.x < (unsigned)p.world) {
    const int r = threadIdx.x;
    if (blockIdx.x == 0) st_release_sys_u32(reinterpret_cast<unsigned*>(p.peer[r] + flag_off) + p.rank * nchx, epoch << 3);
    const unsigned* wf = reinterpret_cast<const unsigned*>(p.peer[p.rank] + flag_off) + r * nchx;
    const long long t0 = clock64();
    while ((ld_acquire_sys_u32(wf) >> 3) < epoch) {
      if (clock64() - t0 > 6000000000ll) { atomicOr(p.status, (unsigned)VJF_ST_COMM_TIMEOUT); dead = 1; break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (dead) return;
  const int n4 = (p.PS + 3) >> 2;
  const float* inbox = p.peer[p.rank];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 v[VJF_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < VJF_MAX_RANKS; ++r)
      v[r] = (r < p.world) ? ld_volatile_f4(inbox + (size_t)(r * 2 + par) * p.PSx + 4 * (size_t)i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a = v[0];
#pragma unroll
    for (int r = 1; r < VJF_MAX_RANKS; ++r) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
    *reinterpret_cast<float4*>(p.reduced + 4 * (size_t)i) = a;
  }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wide_encode_fn() {
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

// L2 promotion of the TMA requests (a box row is one 128-byte request; its neighbour in the row is the next K chunk's)
static CUtensorMapL2promotion wide_l2_promotion() {
  static const int v = getenv("VJF_WIDE_L2PROMO") ? atoi(getenv("VJF_WIDE_L2PROMO")) : 128;
  return v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
}

// [rows][cols] fp32 row-major, row stride ld floats: boxes of {32 columns, box_rows rows}
static int wmap2d(CUtensorMap* m, const float* ptr, int rows, int cols, int ld, int box_rows, bool atom32) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = wide_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, wide_l2_promotion(),
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { vjf_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -2; }
  return 0;
}

static bool wide_shapes_ok(const StepParams& p) {
  return !p.ext && p.L == 1 && p.D > 480 && p.D <= WM_NC * VJF_NT && p.D % 4 == 0 && p.H[0] % 4 == 0 && p.H[0] <= wg::BN && p.d <= 8 && p.R <= 128;
}

static size_t wide_gemm_smem(int nimg, int* stages) {
  const int st = std::min(wg::MAXST, (int)((200 * 1024) / (nimg * wg::IMG)));
  *stages = st;
  return (size_t)st * nimg * wg::IMG + 256 + 1024;
}

int vjf_wide_create(vjf_handle* h) {
  const StepParams& p = h->base;
  h->wide = nullptr;
  if (!wide_shapes_ok(p) || !wide_encode_fn()) return 0;
  WideSm s = wide_plan(p, true);
  if ((size_t)s.total * 4 > h->smem_limit) s = wide_plan(p, false);
  if ((size_t)s.total * 4 > h->smem_limit) return 0;
  Wide* w = (Wide*)calloc(1, sizeof(Wide));
  const size_t B = (size_t)h->cfg.max_trials, H = (size_t)p.H[0], D = (size_t)p.D;
  w->Bmax = (int)B;
  w->Bp = (int)((B + 3) & ~(size_t)3); if ((w->Bp * 4) % 4096 == 0) w->Bp += 32;
  w->Dp = (int)((D + 3) & ~(size_t)3); if ((w->Dp * 4) % 4096 == 0) w->Dp += 32;
  const int ZFmax = 4, ZDmax = 20;
  auto alloc = [&](float** ptr, size_t n) { if (cudaMalloc(ptr, n * sizeof(float)) != cudaSuccess) return -1; return cudaMemset(*ptr, 0, n * sizeof(float)) == cudaSuccess ? 0 : -1; };
  int bad = 0;
  bad |= alloc(&w->W1T, H * w->Dp); bad |= alloc(&w->W1Tlo, H * w->Dp); bad |= alloc(&w->pre, (size_t)ZFmax * B * H);
  bad |= alloc(&w->GT, H * (size_t)w->Bp); bad |= alloc(&w->GTlo, H * (size_t)w->Bp); bad |= alloc(&w->dWT, (size_t)ZDmax * H * w->Dp);
  if (bad) { vjf_set_error("ydim=%d, max_trials=%d: out of device memory for the wide-observation workspace", p.D, h->cfg.max_trials); free(w); return -2; }
  h->wide = w;
  VJF_CUDA_OK(cudaMalloc(&w->flag, 1024));
  VJF_CUDA_OK(cudaMemset(w->flag, 0, 1024));
  if (getenv("VJF_WIDE_STAMPS")) w->stamps = reinterpret_cast<unsigned long long*>(w->flag + 64);
  VJF_CUDA_OK(cudaStreamCreateWithFlags(&w->side, cudaStreamNonBlocking));
  VJF_CUDA_OK(cudaEventCreateWithFlags(&w->ev_sgd, cudaEventDisableTiming));
  VJF_CUDA_OK(cudaEventCreateWithFlags(&w->ev_rls, cudaEventDisableTiming));
  VJF_CUDA_OK(cudaFuncSetAttribute(vjf_rls_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  VJF_CUDA_OK(cudaFuncSetAttribute(wide_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024 + 256 + 1024)));
  VJF_CUDA_OK(cudaFuncSetAttribute(wide_mid_kernel<VJF_LIK_POISSON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  VJF_CUDA_OK(cudaFuncSetAttribute(wide_mid_kernel<VJF_LIK_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  return 0;
}

void vjf_wide_destroy(vjf_handle* h) {
  Wide* w = h->wide;
  if (!w) return;
  if (w->side) cudaStreamDestroy(w->side);
  if (w->ev_sgd) cudaEventDestroy(w->ev_sgd);
  if (w->ev_rls) cudaEventDestroy(w->ev_rls);
  cudaFree(w->flag); cudaFree(w->W1T); cudaFree(w->W1Tlo); cudaFree(w->pre); cudaFree(w->GT); cudaFree(w->GTlo); cudaFree(w->dWT); cudaFree(w->ylo);
  free(w);
  h->wide = nullptr;
}

// development aid: VJF_WIDE_TIMING=1 prints the CUDA-event time of every launch of the last time step to stderr
struct WideTimer {
  bool on; cudaStream_t s; cudaEvent_t ev[16]; const char* name[16]; int n;
  void mark(const char* nm) { if (!on || n >= 16) return; if (!ev[n]) cudaEventCreate(&ev[n]); cudaEventRecord(ev[n], s); name[n++] = nm; }
  void report() {
    if (!on || n < 2) return;
    cudaEventSynchronize(ev[n - 1]);
    for (int i = 0; i + 1 < n; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); fprintf(stderr, "  wide %-10s %8.1f us\n", name[i], ms * 1e3f); }
  }
};

// Does this launch run the wide-observation path?  (fp32 observations, 16-byte aligned; shapes checked at create)
bool vjf_wide_applies(vjf_handle* h, const StepParams& p, int B) {
  static const bool disabled = getenv("VJF_B200_NO_WIDE") != nullptr;
  if (disabled || vjf_tile_mode_get() == 1 || !h->wide || p.y_dtype != VJF_Y_F32 || (reinterpret_cast<uintptr_t>(p.y) & 15) != 0 || B > h->wide->Bmax) return false;
  if (p.world > 1 && !h->xbuf) return false;
  return true;
}

int vjf_wide_time_loop(vjf_handle* h, StepParams& p0, int T, int B, cudaStream_t s) {
  static WideTimer tm = {getenv("VJF_WIDE_TIMING") != nullptr, nullptr, {}, {}, 0};
  tm.s = s;
  Wide& w = *h->wide;
  StepParams pl = p0;
  if (vjf_plan_tiles_public(h, pl, B)) return -1;  // shared-memory plan of phase B (k_split.cu)
  const int D = pl.D, H = pl.H[0], d = pl.d;
  pl.B = B; pl.T = 1;
  if (pl.world == 1) pl.Bglobal = B;
  // spike counts are exact in tf32 (no lo image of the observations); anything else gets one per step
  bool exact = false;
  if (vjf_observations_exact(h, pl.y, (size_t)T * B * D, s, &exact)) return -2;
  if (!exact && !w.ylo) {
    if (cudaMalloc(&w.ylo, (size_t)w.Bmax * D * sizeof(float)) != cudaSuccess) { vjf_set_error("out of device memory for the lo image of the observations"); return -2; }
  }
  WideSm sm = wide_plan(pl, true);
  if ((size_t)sm.total * 4 > h->smem_limit) sm = wide_plan(pl, false);
  // trials per CTA of the mid kernel: balanced, a multiple of 4; every CTA has work
  const int G0 = std::min(h->max_slots - 1, (B + 3) / 4);  // one SM stays free for the side-stream RLS launch (see wide_mid_kernel)
  w.per = (((B + G0 - 1) / G0) + 3) & ~3;
  w.nslots = (B + w.per - 1) / w.per;
  // split-K factors: fill the SMs
  // two output tiles per CTA share the weight / g_pre images of a stage (fewer bytes from L2 per tensor-core instruction:
  // the GEMMs are bound by the per-SM load bandwidth, not by the tensor pipe)
  static const bool no_dual = getenv("VJF_WIDE_NO_DUAL") != nullptr;
  // (pays once a GEMM is more than one wave of CTAs; below that the doubled split-K partial sums cost more than they save)
  const int dual_f = (!no_dual && B > 16384) ? 1 : 0, dual_d = (!no_dual && B > 16384 && D > wg::BN) ? 2 : 0;
  const int mt = (B + wg::BM * (dual_f ? 2 : 1) - 1) / (wg::BM * (dual_f ? 2 : 1)), nkf = (D + 31) / 32;
  w.ZF = std::max(1, std::min(std::min(4, nkf), h->num_sms / std::max(1, mt)));
  const int ntd = (D + wg::BN * (dual_d ? 2 : 1) - 1) / (wg::BN * (dual_d ? 2 : 1)), nkd = (B + 31) / 32;
  w.ZD = std::max(1, std::min(std::min(20, nkd), h->num_sms / std::max(1, ntd)));
  int st_f, st_d;
  const size_t smem_f = wide_gemm_smem(3 + (exact ? 0 : 1) + (dual_f ? (exact ? 1 : 2) : 0), &st_f);
  const size_t smem_d = wide_gemm_smem(3 + (exact ? 0 : 1) + (dual_d ? (exact ? 1 : 2) : 0), &st_d);
  CUtensorMap mW, mWlo, mG, mGlo;
  if (wmap2d(&mW, w.W1T, H, D, w.Dp, wg::BN, false) || wmap2d(&mWlo, w.W1Tlo, H, D, w.Dp, wg::BN, false) ||
      wmap2d(&mG, w.GT, H, B, w.Bp, wg::BM, false) || wmap2d(&mGlo, w.GTlo, H, B, w.Bp, wg::BM, false)) return -2;
  const size_t PSx = (size_t)((pl.PS + 127) & ~127);
  const int nb_grid = std::max(1, std::min(h->num_sms, (pl.lay.n_train + VJF_NT - 1) / VJF_NT));
  static const bool no_overlap = getenv("VJF_WIDE_NO_OVERLAP") != nullptr;
  const bool overlap = !no_overlap && !tm.on;
  static const bool inkernel_wait = getenv("VJF_WIDE_EVENT_WAIT") == nullptr;
  for (int t = 0; t < T; ++t) {
    StepParams p = pl;
    const float* yt = reinterpret_cast<const float*>(pl.y) + (size_t)t * B * D;
    p.y = yt;
    p.u_in = pl.u_in ? pl.u_in + (size_t)t * B * pl.u : nullptr;
    p.eps = pl.eps ? pl.eps + (size_t)t * 2 * B * d : nullptr;
    p.mu = pl.mu + (size_t)t * B * d; p.logvar = pl.logvar + (size_t)t * B * d;
    p.losses = pl.losses ? pl.losses + (size_t)t * 4 : nullptr;
    if (t > 0) { p.q0m = pl.mu + (size_t)(t - 1) * B * d; p.q0l = pl.logvar + (size_t)(t - 1) * B * d; p.flags = pl.flags & ~(uint32_t)VJF_FLAG_PRIOR_Q0; }
    p.step0 = pl.step0 + t;
    CUtensorMap mYk, mYlok, mYm, mYlom;
    if (wmap2d(&mYk, yt, B, D, D, wg::BM, false) || wmap2d(&mYm, yt, B, D, D, 32, true)) return -2;
    mYlok = mYk; mYlom = mYm;
    tm.n = 0;
    tm.mark("w1t");
    wide_w1t_kernel<<<dim3((w.Dp + 31) / 32, (H + 31) / 32), dim3(32, 8), 0, s>>>(h->state + pl.lay.mlp_w[0], w.W1T, w.W1Tlo, D, H, w.Dp);
    if (!exact) {
      if (wmap2d(&mYlok, w.ylo, B, D, D, wg::BM, false) || wmap2d(&mYlom, w.ylo, B, D, D, 32, true)) return -2;
      wide_ylo_kernel<<<h->num_sms * 4, 256, 0, s>>>(reinterpret_cast<const float4*>(yt), reinterpret_cast<float4*>(w.ylo), (size_t)B * D / 4);
      ++g_vjf_launches;
    }
    tm.mark("gemm_fwd");
    {
      wg::Args g = {B, H, D, 0, exact ? 0 : 1, 1, dual_f, st_f, w.pre, H};
      wide_gemm_kernel<<<dim3(mt, 1, w.ZF), wg::NT, smem_f, s>>>(mYk, mYlok, mW, mWlo, g);
    }
    tm.mark("mid");
    // (the RLS of step t-1 runs on the side stream: the mid kernel waits for its completion counter in-kernel; with the
    // event-based variant the wait sits in front of the whole kernel)
    if (t > 0 && overlap && !inkernel_wait) VJF_CUDA_OK(cudaStreamWaitEvent(s, w.ev_rls, 0));
    const unsigned wait_seq = (overlap && inkernel_wait) ? w.seq : 0u;
    if (pl.lik == VJF_LIK_GAUSSIAN) wide_mid_kernel<VJF_LIK_GAUSSIAN><<<w.nslots, VJF_NT, (size_t)sm.total * 4, s>>>(p, w, sm, wait_seq);
    else wide_mid_kernel<VJF_LIK_POISSON><<<w.nslots, VJF_NT, (size_t)sm.total * 4, s>>>(p, w, sm, wait_seq);
    tm.mark("gemm_dw");
    {
      wg::Args g = {H, D, B, 1, 1, exact ? 0 : 1, dual_d, st_d, w.dWT, w.Dp};
      wide_gemm_kernel<<<dim3(1, ntd, w.ZD), wg::NT, smem_d, s>>>(mG, mGlo, mYm, mYlom, g);
    }
    tm.mark("reduce");
    // sharded: the sums are pushed into every rank's exchange buffer (inbox slot of this rank, parity of the epoch)
    const unsigned epoch = pl.epoch0 + (unsigned)t + 1;
    float* red_out = p.reduced;
    {
      const int ntj = (D + 31) / 32, nW1 = ntj * ((H + 31) / 32);
      wide_reduce_kernel<<<nW1 + (pl.PS - D * H + 31) / 32, 1024, 0, s>>>(p, w, red_out, ntj, nW1, pl.world > 1 ? 1 : 0, (int)(epoch & 1));
    }
    if (pl.world > 1) {
      tm.mark("exchange");
      wide_exchange_kernel<<<std::min(4 * h->num_sms, (pl.PS / 4 + 255) / 256), 256, 0, s>>>(p, epoch);
      ++g_vjf_launches;
    }
    tm.mark("phase_b");
    if (overlap) {
      vjf_sgd_kernel<<<nb_grid, VJF_NT, 0, s>>>(p);
      if (w.stamps) wide_stamp_kernel<<<1, 1, 0, s>>>(w.stamps + (p.step0 & 7) * 8 + 6);
      VJF_CUDA_OK(cudaEventRecord(w.ev_sgd, s));
      VJF_CUDA_OK(cudaStreamWaitEvent(w.side, w.ev_sgd, 0));
      if (w.stamps) wide_stamp_kernel<<<1, 1, 0, w.side>>>(w.stamps + (p.step0 & 7) * 8 + 0);
      vjf_rls_kernel<<<1, VJF_NT, (size_t)p.s_total * sizeof(float), w.side>>>(p);
      wide_flag_kernel<<<1, 1, 0, w.side>>>(w.flag, ++w.seq, w.stamps ? w.stamps + (p.step0 & 7) * 8 + 1 : nullptr);
      VJF_CUDA_OK(cudaEventRecord(w.ev_rls, w.side));
      if (t == T - 1) VJF_CUDA_OK(cudaStreamWaitEvent(s, w.ev_rls, 0));
      ++g_vjf_launches;
    } else {
      vjf_phase_b_kernel<<<nb_grid, VJF_NT, (size_t)p.s_total * sizeof(float), s>>>(p);
    }
    g_vjf_launches += 6;
    tm.mark("end");
  }
  tm.report();
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}

// development aid: device pointer of the stamp words (NULL unless VJF_WIDE_STAMPS was set at create)
extern "C" unsigned long long* vjf_wide_stamps(vjf_handle* h) { return (h && h->wide) ? h->wide->stamps : nullptr; }
