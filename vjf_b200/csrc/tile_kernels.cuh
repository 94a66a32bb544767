// Throughput tile pipeline of the VJF filter + learning step (sm_100a): the trial-parallel phase of a time step
// (vjf/model.py:97-154 forward + ELBO, :209 backward, sufficient statistics of vjf/module.py:94-96) for tiles of TBR trials
// with EVERY contraction on the 5th-generation tensor cores:
//
//   FWD   D1[trial][n]   = in[trial][k] W1[k][n]              A = input image, K-major SW128 (written by TMA tensor copies)
//   QUAD  FL[trial][n']  = phi[trial][r] [w_chol | w_mean]    A = phi image, K-major SW128         (vjf/module.py:75-77)
//   GRAM  A[r][r'], b    = phi^T [phi | dx]                   A, B = phi image, MN-major BASE32B   (vjf/module.py:94-96)
//   DW    dW1[k][n]      = in^T g_pre                         A = input image, B = g_pre image, both MN-major BASE32B
//
// all as tcgen05.mma.cta_group::1.kind::tf32 (M = 128) with fp32-grade accuracy from three products per contraction: the
// tensor core truncates fp32 operands to tf32, so an image holds the raw value x ("hi") and a second image x - trunc(x)
// ("lo"), and hi*hi + hi*lo + lo*hi is accumulated in tensor memory.  One [trials][columns] tile serves both operand roles:
// K-major for the forward GEMM, then -- after an in-place permutation of the 16-byte pieces of every 128-byte row -- MN-major
// for the GEMMs whose contraction runs over the trials (profiles/micro_r02.txt has the validated descriptor forms).
// GRAM and DW accumulate in tensor memory across all tiles of a CTA and are flushed to its slot ONCE per time step.
//
// Warp roles: warp 15 is the control warp (TMA issue, weight ring, tcgen05.mma issue, commits); warps 0-14 compute.  They meet
// only through mbarriers.  The shared phases B1 (slot reduction, SGD, NVLink exchange) and B2 (RLS) are those of
// step_kernels.cuh; the schedule around them is the overlapped one (CTA 0 = RLS CTA).
#pragma once
#include <cstdio>
#include <cuda.h>
#include "step_kernels.cuh"

#define TK_NCT (TK_NCW * 32)
#define TK_CTRL 15
#define TK_MAXNS 4

// development aid (debug builds: VJF_B200_DEBUG=1 python -m vjf_b200.build): globaltimer stamps of one trial CTA, first tile of every step
#ifdef VJF_DEBUG_STAMPS
#define TK_STAMP(p, t, j, who, idx) do { if ((p).dbg && (int)blockIdx.x == (p).dbg_cta && (j) == 0 && threadIdx.x == (who)) (p).dbg[(t) * 64 + (idx)] = gtime_ns(); } while (0)
#else
#define TK_STAMP(p, t, j, who, idx) do {} while (0)
#endif

enum {
  BK_YFULL0 = 0, BK_YFULL1, BK_WFULL0, BK_WEMPTY0 = BK_WFULL0 + TK_MAXNS, BK_UK = BK_WEMPTY0 + TK_MAXNS, BK_D1, BK_FL, BK_GRAM, BK_DW,
  BK_CX, BK_CPHI, BK_CPHIT, BK_CG, BK_N
};

__device__ __forceinline__ void cb_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TK_NCT) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait with a watchdog: a protocol error traps (the launch fails with an error) instead of hanging the device
__device__ __forceinline__ void tk_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spins > (1u << 26)) { printf("vjf_tile_kernel: mbarrier @%u time-out (cta %d thread %d parity %u)\n", smem_u32(bar), (int)blockIdx.x, (int)threadIdx.x, parity); __trap(); }
  }
}
__device__ __forceinline__ uint64_t tk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
         (1ull << 46) | ((uint64_t)type << 61);
}
__device__ __forceinline__ uint64_t tk_kmaj(uint32_t a) { return tk_desc(a, 16, 1024, 2); }                       // K-major, SWIZZLE_128B
__device__ __forceinline__ uint64_t tk_mnmaj(uint32_t a, uint32_t chunk) { return tk_desc(a, chunk, 512, 1); }    // MN-major, SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint32_t tk_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row, col) of a [rows][32 * chunks] fp32 image stored as 32-column chunks of [rows][128 B]
__device__ __forceinline__ int sw128_off(int row, int col, int rows) {   // 16-byte pieces XOR (row % 8): what TMA SWIZZLE_128B writes
  const int c = col & 31;
  return ((col >> 5) * rows + row) * 128 + ((((c >> 2) ^ (row & 7)) << 4) | ((c & 3) << 2));
}
__device__ __forceinline__ int b32_off(int row, int col, int rows) {     // 32-byte pieces XOR (row % 4): SWIZZLE_128B_BASE32B
  return (((col >> 5) * rows + row) * 128 + ((col & 31) << 2)) ^ ((row & 3) << 5);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tk_signal(uint64_t* bar) {  // one arrival per compute warp; orders this warp's shared-memory writes before the MMA
  fence_async_smem();
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// compute warps: wait until a monotonically increasing device counter reaches `want`
__device__ __forceinline__ void tk_wait_counter(const unsigned* ctr, unsigned want) {
  if (threadIdx.x == 0) {
    while (ld_acquire_u32(ctr) < want) __nanosleep(32);
    __threadfence();
  }
  cb_sync();
}
// in-place permutation of one 128-byte row: SWIZZLE_128B image -> SWIZZLE_128B_BASE32B image (same logical [row][32 floats])
__device__ __forceinline__ void tk_permute_row(unsigned char* rowp, int row) {
  float4* q = reinterpret_cast<float4*>(rowp);
  float4 v[8];
#pragma unroll
  for (int pc = 0; pc < 8; ++pc) v[pc] = q[pc ^ (row & 7)];
#pragma unroll
  for (int pc = 0; pc < 8; ++pc) q[(((pc >> 1) ^ (row & 3)) << 1) | (pc & 1)] = v[pc];
}

// tile index of the j-th tile of this CTA (trial CTAs are blocks 1 .. gridDim.x - 1)
__device__ __forceinline__ int tk_tile_of(int j) { return (int)blockIdx.x - 1 + j * ((int)gridDim.x - 1); }

// ------------------------------------------------------------------------------------------------------------------
// control warp: one time step's worth of TMA / tcgen05 issue for the tiles of this CTA
// ------------------------------------------------------------------------------------------------------------------
struct TkCtl { uint32_t rp, rc, ukn; };  // weight-ring items produced / consumed, UK loads (kernel lifetime)

static __device__ __forceinline__ void tk_issue_y(const StepParams& p, const CUtensorMap* ymap, unsigned char* sb, uint64_t* bars, int t, int tile, int buf) {
  const TilePlan& pl = p.tp;
  uint64_t* bar = &bars[BK_YFULL0 + buf];
  mbar_expect_tx(bar, (uint32_t)pl.NCY * pl.TBR * 128u);
  for (int c = 0; c < pl.NCY; ++c) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(sb + pl.o_in[buf] + c * pl.TBR * 128)), "l"(ymap), "r"(c * 32), "r"(tile * pl.TBR), "r"(t), "r"(smem_u32(bar)) : "memory");
  }
}
static __device__ __forceinline__ void tk_produce(const StepParams& p, unsigned char* sb, uint64_t* bars, uint32_t item, int chunk) {
  const TilePlan& pl = p.tp;
  const int s = item % pl.NS;
  const uint32_t u = item / pl.NS;
  if (u >= 1) tk_wait(&bars[BK_WEMPTY0 + s], (u - 1) & 1);
  const uint32_t bytes = 2u * p.H[0] * 128u;
  mbar_expect_tx(&bars[BK_WFULL0 + s], bytes);
  tma_bulk_g2s(sb + pl.o_ring + s * pl.SS, p.w1k + (size_t)chunk * (2 * p.H[0] * 32), bytes, &bars[BK_WFULL0 + s]);
}

static __device__ void tk_control_step(const StepParams& p, const CUtensorMap* ymap, unsigned char* sb, uint64_t* bars, uint32_t tmem, int t,
                                       int ntl, uint32_t it0, TkCtl& cs) {
  const TilePlan& pl = p.tp;
  const int lane = threadIdx.x & 31;
  const int H = p.H[0], TBR = pl.TBR;
  const uint32_t chunkB = (uint32_t)TBR * 128u;
  const uint32_t sbase = smem_u32(sb);
  const uint32_t id_fwd = tk_idesc(128, H, 0, 0), id_quad = tk_idesc(128, pl.NQ, 0, 0), id_gram = tk_idesc(128, pl.NQ, 1, 1), id_dw = tk_idesc(128, H, 1, 1);
  const uint32_t rend = cs.rc + (uint32_t)ntl * pl.NCH;
  if (lane == 0) {
    // the weight images were written by other CTAs through the generic proxy (ordered by the grid barrier / counters)
    asm volatile("fence.proxy.async.global;" ::: "memory");
    for (int b = 0; b < pl.NBUF && b < ntl; ++b) tk_issue_y(p, ymap, sb, bars, t, tk_tile_of(b), (int)((it0 + b) % pl.NBUF));
    while (cs.rp < rend && cs.rp < cs.rc + pl.NS) { tk_produce(p, sb, bars, cs.rp, (int)((cs.rp - (rend - (uint32_t)ntl * pl.NCH)) % pl.NCH)); ++cs.rp; }
  }
  const uint32_t rbase = rend - (uint32_t)ntl * pl.NCH;
  TK_STAMP(p, t, 0, TK_CTRL * 32, 32);
  for (int j = 0; j < ntl; ++j) {
    const uint32_t it = it0 + j;
    const int buf = (int)(it % pl.NBUF);
    const uint32_t par = it & 1;
    // ---------------- FWD ----------------
    tk_wait(&bars[BK_CX], par);
    tc_fence_after();
    TK_STAMP(p, t, j, TK_CTRL * 32, 33);
    if (lane == 0) {
      for (int c = 0; c < pl.NCH; ++c) {
        const int s = cs.rc % pl.NS;
        tk_wait(&bars[BK_WFULL0 + s], (cs.rc / pl.NS) & 1);
        tc_fence_after();
        const uint32_t a_hi = sbase + pl.o_in[buf] + c * chunkB, a_lo = sbase + pl.o_inlo + c * chunkB;
        const uint32_t b_hi = sbase + pl.o_ring + s * pl.SS, b_lo = b_hi + H * 128;
        const int nks = min(4, (pl.K1b - 32 * c + 7) >> 3);
        for (int ks = 0; ks < nks; ++ks) {
          umma_tf32_ss(tmem + pl.c_d1, tk_kmaj(a_lo + 32 * ks), tk_kmaj(b_hi + 32 * ks), id_fwd, (c | ks) ? 1u : 0u);
          umma_tf32_ss(tmem + pl.c_d1, tk_kmaj(a_hi + 32 * ks), tk_kmaj(b_lo + 32 * ks), id_fwd, 1u);
          umma_tf32_ss(tmem + pl.c_d1, tk_kmaj(a_hi + 32 * ks), tk_kmaj(b_hi + 32 * ks), id_fwd, 1u);
        }
        umma_commit(&bars[BK_WEMPTY0 + s]);
        ++cs.rc;
        if (cs.rp < rend) { tk_produce(p, sb, bars, cs.rp, (int)((cs.rp - rbase) % pl.NCH)); ++cs.rp; }
      }
      umma_commit(&bars[BK_D1]);
    }
    __syncwarp();
    TK_STAMP(p, t, j, TK_CTRL * 32, 34);
    // ---------------- QUAD ----------------
    tk_wait(&bars[BK_CPHI], par);
    tc_fence_after();
    TK_STAMP(p, t, j, TK_CTRL * 32, 35);
    if (lane == 0) {
      if (j == 0) {
        // [w_chol^T ; w_mean^T] of the previous step: final once the RLS CTA has published it
        if (t > 0) { while (ld_acquire_u32(p.ctrl + 5) < (unsigned)t) __nanosleep(32); }
        asm volatile("fence.proxy.async.global;" ::: "memory");
        mbar_expect_tx(&bars[BK_UK], 2u * pl.ukimg);
        tma_bulk_g2s(sb + pl.o_uk, p.uk, 2u * pl.ukimg, &bars[BK_UK]);
        tk_wait(&bars[BK_UK], cs.ukn & 1);
        ++cs.ukn;
      }
      TK_STAMP(p, t, j, TK_CTRL * 32, 36);
      const uint32_t p_hi = sbase + pl.o_pg, p_lo = p_hi + pl.PWC * chunkB;
      const uint32_t u_hi = sbase + pl.o_uk, u_lo = u_hi + pl.ukimg;
      const int nch = (pl.Rk + 31) >> 5;
      for (int ch = 0; ch < nch; ++ch) {
        const int nks = min(4, (pl.Rk - 32 * ch) >> 3);
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t ao = ch * chunkB + 32 * ks, bo = ch * pl.NQ * 128 + 32 * ks;
          umma_tf32_ss(tmem + pl.c_fl, tk_kmaj(p_lo + ao), tk_kmaj(u_hi + bo), id_quad, (ch | ks) ? 1u : 0u);
          umma_tf32_ss(tmem + pl.c_fl, tk_kmaj(p_hi + ao), tk_kmaj(u_lo + bo), id_quad, 1u);
          umma_tf32_ss(tmem + pl.c_fl, tk_kmaj(p_hi + ao), tk_kmaj(u_hi + bo), id_quad, 1u);
        }
      }
      umma_commit(&bars[BK_FL]);
    }
    __syncwarp();
    TK_STAMP(p, t, j, TK_CTRL * 32, 37);
    // ---------------- GRAM ----------------
    tk_wait(&bars[BK_CPHIT], par);
    tc_fence_after();
    TK_STAMP(p, t, j, TK_CTRL * 32, 38);
    if (lane == 0) {
      const uint32_t p_hi = sbase + pl.o_pg, p_lo = p_hi + pl.PWC * chunkB;
      for (int ks = 0; ks < (TBR >> 3); ++ks) {
        const uint64_t dh = tk_mnmaj(p_hi + ks * 1024, chunkB), dl = tk_mnmaj(p_lo + ks * 1024, chunkB);
        umma_tf32_ss(tmem + pl.c_gram, dl, dh, id_gram, (j | ks) ? 1u : 0u);
        umma_tf32_ss(tmem + pl.c_gram, dh, dl, id_gram, 1u);
        umma_tf32_ss(tmem + pl.c_gram, dh, dh, id_gram, 1u);
      }
      umma_commit(&bars[BK_GRAM]);
    }
    __syncwarp();
    TK_STAMP(p, t, j, TK_CTRL * 32, 39);
    // ---------------- DW ----------------
    tk_wait(&bars[BK_CG], par);
    tc_fence_after();
    TK_STAMP(p, t, j, TK_CTRL * 32, 40);
    if (lane == 0) {
      const uint32_t g_hi = sbase + pl.o_pg, g_lo = g_hi + pl.HC * chunkB;
      for (int mb = 0; mb < pl.NBLK; ++mb) {
        const uint32_t a_hi = sbase + pl.o_in[buf] + 4 * mb * chunkB, a_lo = sbase + pl.o_inlo + 4 * mb * chunkB;
        const uint32_t d = tmem + pl.c_dw + mb * H;
        for (int ks = 0; ks < (TBR >> 3); ++ks) {
          const uint64_t ah = tk_mnmaj(a_hi + ks * 1024, chunkB), al = tk_mnmaj(a_lo + ks * 1024, chunkB);
          const uint64_t bh = tk_mnmaj(g_hi + ks * 1024, chunkB), bl = tk_mnmaj(g_lo + ks * 1024, chunkB);
          umma_tf32_ss(d, al, bh, id_dw, (j | ks) ? 1u : 0u);
          umma_tf32_ss(d, ah, bl, id_dw, 1u);
          umma_tf32_ss(d, ah, bh, id_dw, 1u);
        }
      }
      umma_commit(&bars[BK_DW]);
      TK_STAMP(p, t, j, TK_CTRL * 32, 41);
      // the input buffer is free once these MMAs are done: observations of the tile after next
      if (j + pl.NBUF < ntl) {
        tk_wait(&bars[BK_DW], par);
        tk_issue_y(p, ymap, sb, bars, t, tk_tile_of(j + pl.NBUF), buf);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// compute warps: per-step accumulators that live in registers across the tiles of a CTA
// ------------------------------------------------------------------------------------------------------------------
template <int DX>
struct TkAcc {
  float gdw[DX], gdb;          // decoder gradient of this lane's observation column (likelihood stage)
  float ghm[4][DX], ghv[4][DX];  // head weight gradients of this lane's hidden units (up to H = 128)
  float ghvb;                  // logvar-head bias gradient (threads ctid < d)
  float sc[VJF_NSCAL];
};

template <int DX>
static __device__ void tk_compute_tile(const StepParams& p, unsigned char* sb, uint64_t* bars, uint32_t tmem, int t, int tile, uint32_t it,
                                       unsigned masks, TkAcc<DX>& acc, int j) {
  const TilePlan& pl = p.tp;
  const int ctid = threadIdx.x, lane = ctid & 31, cw = ctid >> 5;
  const int D = p.D, d = p.d, u = p.u, R = p.R, du = p.du, E = p.E, H = p.H[0], TBR = pl.TBR;
  const int b0 = tile * TBR, nb = min(TBR, p.B - b0);
  const int buf = (int)(it % pl.NBUF);
  const uint32_t par = it & 1;
  float* sf = reinterpret_cast<float*>(sb + pl.o_f);
  float* hs = sf + pl.f_hs;       const float* dec = sf + pl.f_dec;  const float* hm_s = sf + pl.f_hm;  const float* hv_s = sf + pl.f_hv;
  const float* c_s = sf + pl.f_cen; const float* iw_s = sf + pl.f_iw;
  float* ex_s = sf + pl.f_ex;     float* eps_s = sf + pl.f_eps;      float* xu_s = sf + pl.f_xu;        float* xt_s = sf + pl.f_xt;
  float* mt_s = sf + pl.f_mt;     float* lt_s = sf + pl.f_lt;        float* pm_s = sf + pl.f_pm;        float* dx_s = sf + pl.f_dx;
  float* gxt_s = sf + pl.f_gxt;   float* gmt_s = sf + pl.f_gmt;      float* glt_s = sf + pl.f_glt;      float* plv_s = sf + pl.f_plv;
  float* gxp_s = sf + pl.f_gxp;
  unsigned char* in_b = sb + pl.o_in[buf];
  unsigned char* inlo_b = sb + pl.o_inlo;
  unsigned char* pg = sb + pl.o_pg;
  float* st = p.state;
  const bool r_on = masks & 1u, d_on = masks & 2u, h_on = masks & 4u;
  const int ldhs = pl.ldhs;

  // ---- A1: previous posterior, control input, noise; xs = m_s + eps1 exp(l_s / 2) (vjf/util.py:11-13); RBF features ----
  {
  TK_STAMP(p, t, j, 0, 0);
    const size_t row0 = (size_t)t * p.B + b0;
    const bool prior = (t == 0) && (p.flags & VJF_FLAG_PRIOR_Q0);
    const float* qm = (t == 0) ? p.q0m : p.mu + (size_t)(t - 1) * p.B * d;
    const float* ql = (t == 0) ? p.q0l : p.logvar + (size_t)(t - 1) * p.B * d;
    for (int i = ctid; i < TBR * E; i += TK_NCT) {
      const int b = i / E, e = i - b * E;
      float v = 0.f;
      if (b < nb) {
        if (e < u) v = p.u_in[(row0 + b) * u + e];
        else if (e < u + d) v = prior ? st[p.lay.prior_mean + e - u] : qm[(size_t)(b0 + b) * d + e - u];
        else v = prior ? st[p.lay.prior_logvar + e - u - d] : ql[(size_t)(b0 + b) * d + e - u - d];
      }
      ex_s[i] = v;
    }
    if (p.eps) {
      for (int i = ctid; i < TBR * 2 * d; i += TK_NCT) {
        const int b = i / (2 * d), k = i - b * 2 * d;
        float v = 0.f;
        if (b < nb) { const float* e0 = p.eps + ((size_t)t * 2 * p.B + b0 + b) * d; v = (k < d) ? e0[k] : e0[(size_t)p.B * d + k - d]; }
        eps_s[i] = v;
      }
    } else {
      const int nblk = (d + 3) >> 2;
      for (int i = ctid; i < TBR * 2 * nblk; i += TK_NCT) {
        const int b = i / (2 * nblk), r = i - b * 2 * nblk, which = r / nblk, blk = r - which * nblk;
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (b < nb) philox_normal4(p.seed, p.step0 + t, p.trial_offset + b0 + b, which, blk, z);
        for (int k = 0; k < 4; ++k)
          if (blk * 4 + k < d) eps_s[b * 2 * d + which * d + blk * 4 + k] = z[k];
      }
    }
    cb_sync();
    TK_STAMP(p, t, j, 0, 1);
    for (int i = ctid; i < TBR * du; i += TK_NCT) {
      const int b = i / du, k = i - b * du;
      float v;
      if (k < d) v = ex_s[b * E + u + k] + eps_s[b * 2 * d + k] * expf(0.5f * ex_s[b * E + u + d + k]);
      else v = ex_s[b * E + (k - d)];
      xu_s[i] = v;
    }
    cb_sync();
    // the phi / g_pre region is free once the weight-gradient MMAs of the previous tile are done
    TK_STAMP(p, t, j, 0, 2);
    if (it > 0) tk_wait(&bars[BK_DW], (it - 1) & 1);
    TK_STAMP(p, t, j, 0, 3);
    // phi = exp(-0.5 |xu - c|^2 / w^2) (vjf/functional.py:11-22) as a (hi, lo) pair of K-major SW128 images; pad columns and
    // pad rows are zero
    unsigned char* ph = pg;
    unsigned char* plo = pg + pl.PWC * TBR * 128;
    for (int b = cw; b < TBR; b += TK_NCW) {
      for (int r = lane; r < pl.PW; r += 32) {
        float v = 0.f;
        if (b < nb && r < R) {
          float d2 = 0.f;
          for (int c = 0; c < du; ++c) { const float df = xu_s[b * du + c] - c_s[r * du + c]; d2 = fmaf(df, df, d2); }
          v = expf(d2 * iw_s[r]);
        }
        const int o = sw128_off(b, r, TBR);
        *reinterpret_cast<float*>(ph + o) = v;
        *reinterpret_cast<float*>(plo + o) = v - tf32_trunc_f(v);
      }
    }
    tk_signal(&bars[BK_CPHI]);
    TK_STAMP(p, t, j, 0, 4);
    // observations of this tile (TMA) -> append [u | m_s | l_s | 1] behind them (vjf/recognition.py:32-37; the ones column
    // carries the bias through the MMAs), then the lo image of the whole input tile
    tk_wait(&bars[BK_YFULL0 + buf], (it / pl.NBUF) & 1);
    TK_STAMP(p, t, j, 0, 5);
    if (pl.NCH > pl.NCY) {  // chunks without observation columns are not written by TMA
      float4* z = reinterpret_cast<float4*>(in_b + pl.NCY * TBR * 128);
      for (int i = ctid; i < (pl.NCH - pl.NCY) * TBR * 8; i += TK_NCT) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      cb_sync();
    }
    const int EX1 = E + 1;
    for (int i = ctid; i < TBR * EX1; i += TK_NCT) {
      const int b = i / EX1, e = i - b * EX1;
      const float v = (b < nb) ? (e < E ? ex_s[b * E + e] : 1.0f) : 0.f;
      *reinterpret_cast<float*>(in_b + sw128_off(b, D + e, TBR)) = v;
    }
    cb_sync();
    {
      const float4* src = reinterpret_cast<const float4*>(in_b);
      float4* dst = reinterpret_cast<float4*>(inlo_b);
      for (int i = ctid; i < pl.NCH * TBR * 8; i += TK_NCT) {
        const float4 x = src[i];
        dst[i] = make_float4(x.x - tf32_trunc_f(x.x), x.y - tf32_trunc_f(x.y), x.z - tf32_trunc_f(x.z), x.w - tf32_trunc_f(x.w));
      }
    }
    tk_signal(&bars[BK_CX]);
    TK_STAMP(p, t, j, 0, 6);
  }

  // ---- E1: h = tanh(W1 in + b1) from tensor memory (vjf/recognition.py:38-40); heads (:41-42); xt, dx; posterior out ----
  tk_wait(&bars[BK_D1], par);
  tc_fence_after();
  TK_STAMP(p, t, j, 0, 7);
  {
    const int g = cw & 3, si = cw >> 2, nsw = (g == 3) ? 3 : 4;
    if (g * 32 < TBR) {
      const int row = g * 32 + lane;
      for (int un = si; un < (H >> 3); un += nsw) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(g * 32) << 16) + pl.c_d1 + 8 * un, v);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) hs[row * ldhs + 8 * un + jj] = tanhf(v[jj]);
      }
    }
  }
  tc_fence_before();
  cb_sync();
  TK_STAMP(p, t, j, 0, 8);
  for (int i = ctid; i < TBR * d; i += TK_NCT) {
    const int b = i / d, k = i - b * d;
    float m = 0.f, lv = 0.f, x = 0.f, dxv = 0.f;
    if (b < nb) {
      float m1 = 0.f, l1 = 0.f;
      const float* hr = hs + b * ldhs;
      int n = 0;
      for (; n + 1 < H; n += 2) {
        m = fmaf(hr[n], hm_s[n * d + k], m); lv = fmaf(hr[n], hv_s[n * d + k], lv);
        m1 = fmaf(hr[n + 1], hm_s[(n + 1) * d + k], m1); l1 = fmaf(hr[n + 1], hv_s[(n + 1) * d + k], l1);
      }
      m += m1; lv += l1 + hv_s[H * d + k];
      x = m + eps_s[b * 2 * d + d + k] * expf(0.5f * lv);
      dxv = x - xu_s[b * du + k];
      acc.sc[SC_SDX] = fmaf(dxv, dxv, acc.sc[SC_SDX]);
      p.mu[((size_t)t * p.B + b0 + b) * d + k] = m;       // posterior of this step (model.py:218-221, :305-307)
      p.logvar[((size_t)t * p.B + b0 + b) * d + k] = lv;
    }
    mt_s[i] = m; lt_s[i] = lv; xt_s[i] = x; dx_s[i] = dxv;
  }
  TK_STAMP(p, t, j, 0, 9);
  // ---- the input tile becomes the MN-major operand of the weight gradient: permute every 128-byte row in place ----
  for (int i = ctid; i < 2 * pl.NCH * TBR; i += TK_NCT) {
    const int im = i / (pl.NCH * TBR), rr = i - im * (pl.NCH * TBR);
    tk_permute_row((im ? inlo_b : in_b) + rr * 128, rr % TBR);
  }
  cb_sync();

  TK_STAMP(p, t, j, 0, 10);
  // ---- LK: decoder eta = D xt + bias (model.py:29-30), likelihood terms, d loss / d eta (times B), decoder gradients, g_xt ----
  {
    float lam = 0.f;
    if (p.lik == VJF_LIK_GAUSSIAN) {
      if (t > 0) tk_wait_counter(p.ctrl + 4, (unsigned)t);
      lam = __ldcg(st + p.lay.lik_logvar);
    }
    const float e_nlam = expf(-lam), p_lam = expf(-0.5f * lam);
    if (cw < pl.NCY * pl.RS) {
      const int cy = cw % pl.NCY, rs = cw / pl.NCY;
      const int jcol = 32 * cy + lane;
      const bool jok = jcol < D;
      float w[DX], bj = jok ? dec[d * D + jcol] : 0.f;
#pragma unroll
      for (int k = 0; k < DX; ++k) w[k] = (jok && (DX == d || k < d)) ? dec[k * D + jcol] : 0.f;
      for (int b = rs; b < nb; b += pl.RS) {
        float xt[DX], eta = bj;
#pragma unroll
        for (int k = 0; k < DX; ++k) { xt[k] = (DX == d || k < d) ? xt_s[b * d + k] : 0.f; eta = fmaf(w[k], xt[k], eta); }
        const float yv = *reinterpret_cast<const float*>(in_b + b32_off(b, jcol, TBR));
        float g = 0.f;
        if (jok) {
          if (p.lik == VJF_LIK_GAUSSIAN) {
            // gaussian_loss(y, eta, lambda), functional.py:55-75 ; update's mse, likelihood.py:36-37
            const float r = yv - eta;
            const float rsd = yv * p_lam - eta * p_lam;
            const float mse = rsd * rsd;
            if (!isfinite(mse)) acc.sc[SC_BADMSE] += 1.f;
            acc.sc[SC_RECON] += 0.5f * (mse + lam);
            acc.sc[SC_SSE] = fmaf(r, r, acc.sc[SC_SSE]);
            g = -r * e_nlam;
            acc.sc[6] += r_on ? 0.5f * (1.0f - r * r * e_nlam) : 0.f;
          } else {
            // poisson_nll_loss(clamp(eta, max=10), y, log_input=True), likelihood.py:60-62
            const float ec = fminf(eta, 10.0f);
            const float ex = expf(ec);
            acc.sc[SC_RECON] += ex - yv * ec;
            g = (eta <= 10.0f) ? (ex - yv) : 0.f;
            if (eta != eta) { acc.sc[SC_RECON] = eta; g = eta; }  // NaN propagates like torch.clamp
          }
          g = r_on ? g : 0.f;
        }
        acc.gdb += g;
#pragma unroll
        for (int k = 0; k < DX; ++k) {
          acc.gdw[k] = fmaf(g, xt[k], acc.gdw[k]);
          const float s = warp_sum(g * w[k]);
          if (lane == 0 && (DX == d || k < d)) gxp_s[(cy * TBR + b) * d + k] = s;
        }
      }
    }
    cb_sync();
    for (int i = ctid; i < TBR * d; i += TK_NCT) {
      const int b = i / d, k = i - b * d;
      float s = 0.f;
      if (b < nb)
        for (int cy = 0; cy < pl.NCY; ++cy) s += gxp_s[(cy * TBR + b) * d + k];
      gxt_s[i] = s;
    }
  }

  // ---- phi images: K-major (quadratic form, done) -> MN-major for the Gram matrix; dx goes into the spare columns behind
  //      the features so that phi^T dx comes out of the same MMAs ----
  TK_STAMP(p, t, j, 0, 11);
  tk_wait(&bars[BK_FL], par);
  tc_fence_after();
  TK_STAMP(p, t, j, 0, 12);
  for (int i = ctid; i < 2 * pl.PWC * TBR; i += TK_NCT) tk_permute_row(pg + i * 128, i % TBR);
  cb_sync();
  {
    unsigned char* ph = pg;
    unsigned char* plo = pg + pl.PWC * TBR * 128;
    for (int i = ctid; i < TBR * d; i += TK_NCT) {
      const int b = i / d, k = i - b * d;
      const float v = dx_s[i];
      const int o = b32_off(b, pl.Rk + k, TBR);
      *reinterpret_cast<float*>(ph + o) = v;
      *reinterpret_cast<float*>(plo + o) = v - tf32_trunc_f(v);
    }
  }
  // FL = phi [w_chol | w_mean] from tensor memory: p_logvar = log |phi w_chol|^2, p_mean = xs + phi W (module.py:75-77, model.py:338)
  if (cw * 32 < TBR) {
    const int row = cw * 32 + lane;
    float q = 0.f, pmv[DX];
#pragma unroll
    for (int k = 0; k < DX; ++k) pmv[k] = 0.f;
    for (int un = 0; un < (pl.NQ >> 3); ++un) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(cw * 32) << 16) + pl.c_fl + 8 * un, v);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int col = 8 * un + jj;
        if (col < pl.Rk) q = fmaf(v[jj], v[jj], q);
        else {
#pragma unroll
          for (int k = 0; k < DX; ++k)
            if (col - pl.Rk == k) pmv[k] = v[jj];
        }
      }
    }
    plv_s[row] = logf(q);
#pragma unroll
    for (int k = 0; k < DX; ++k)
      if (DX == d || k < d) pm_s[row * d + k] = xu_s[row * du + k] + pmv[k];
  }
  tk_signal(&bars[BK_CPHIT]);
  TK_STAMP(p, t, j, 0, 13);
  // state-noise logvar of the previous step: published by the RLS CTA after its tail
  if (t > 0) tk_wait_counter(p.ctrl + 3, (unsigned)t);
  else cb_sync();
  TK_STAMP(p, t, j, 0, 14);
  {
    const float gam = __ldcg(st + p.lay.tr_logvar);
    const float e_ngam = expf(-gam), p_gam = expf(-0.5f * gam);
    // dynamics NLL (functional.py:55-75 via model.py:390-391), entropy (functional.py:25-29), g_mt and g_lt (times B)
    for (int i = ctid; i < TBR * d; i += TK_NCT) {
      const int b = i / d, k = i - b * d;
      float gm = 0.f, gl = 0.f;
      if (b < nb) {
        const float m = mt_s[i], lv = lt_s[i], pm = pm_s[i], plv = plv_s[b];
        const float e2 = eps_s[b * 2 * d + d + k], gx = gxt_s[i];
        const float df = pm * p_gam - m * p_gam;
        const float mse = df * df;
        if (!isfinite(mse)) acc.sc[SC_BADMSE] += 1.f;
        const float tr = expf(plv + lv - gam);
        acc.sc[SC_DYN] += 0.5f * (mse + gam) + 0.5f * tr;
        acc.sc[SC_ENT] += 0.5f * lv;
        gm = gx; gl = 0.5f * gx * e2 * expf(0.5f * lv);
        if (h_on) gl -= 0.5f;
        if (d_on) { gm += (m - pm) * e_ngam; gl += 0.5f * tr; }
      }
      gmt_s[i] = gm; glt_s[i] = gl;
    }
  }
  cb_sync();

  // ---- GS: g_pre = (g_mt W_m + g_lt W_v) (1 - h^2) as the MN-major (hi, lo) B operand of the weight gradient; head gradients ----
  TK_STAMP(p, t, j, 0, 15);
  tk_wait(&bars[BK_GRAM], par);
  TK_STAMP(p, t, j, 0, 16);
  {
    unsigned char* gh = pg;
    unsigned char* gl = pg + pl.HC * TBR * 128;
    float wm[4][DX], wv[4][DX];
#pragma unroll
    for (int hc = 0; hc < 4; ++hc)
#pragma unroll
      for (int k = 0; k < DX; ++k) {
        const bool ok = hc < pl.HC && (DX == d || k < d);
        wm[hc][k] = ok ? hm_s[(32 * hc + lane) * d + k] : 0.f;
        wv[hc][k] = ok ? hv_s[(32 * hc + lane) * d + k] : 0.f;
      }
    for (int b = cw; b < TBR; b += TK_NCW) {
      float gm[DX], gv[DX];
#pragma unroll
      for (int k = 0; k < DX; ++k) { const bool ok = (DX == d || k < d); gm[k] = ok ? gmt_s[b * d + k] : 0.f; gv[k] = ok ? glt_s[b * d + k] : 0.f; }
#pragma unroll
      for (int hc = 0; hc < 4; ++hc) {
        if (hc < pl.HC) {
          const int n = 32 * hc + lane;
          const float h = hs[b * ldhs + n];
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < DX; ++k) {
            s = fmaf(gm[k], wm[hc][k], s); s = fmaf(gv[k], wv[hc][k], s);
            acc.ghm[hc][k] = fmaf(h, gm[k], acc.ghm[hc][k]);
            acc.ghv[hc][k] = fmaf(h, gv[k], acc.ghv[hc][k]);
          }
          const float v = (b < nb) ? s * (1.0f - h * h) : 0.f;
          const int o = b32_off(b, n, TBR);
          *reinterpret_cast<float*>(gh + o) = v;
          *reinterpret_cast<float*>(gl + o) = v - tf32_trunc_f(v);
        }
      }
    }
    if (ctid < d) {
      float s = 0.f;
      for (int b = 0; b < nb; ++b) s += glt_s[b * d + ctid];
      acc.ghvb += s;
    }
  }
  tk_signal(&bars[BK_CG]);
  TK_STAMP(p, t, j, 0, 17);
}

// Flush of everything a CTA accumulated over its tiles of one time step into its slot (all 16 warps).
template <int DX>
static __device__ void tk_flush_step(const StepParams& p, unsigned char* sb, uint32_t tmem, TkAcc<DX>& acc, bool have_tiles) {
  const TilePlan& pl = p.tp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D, d = p.d, R = p.R, H = p.H[0];
  float* slot = p.partials + (size_t)blockIdx.x * p.PS;
  float* sf = reinterpret_cast<float*>(sb + pl.o_f);
  float* scr = reinterpret_cast<float*>(sb + pl.o_pg);  // scratch: the phi / g_pre region + weight ring (dead between steps)
  if (!have_tiles) {
    for (int i = tid; i < p.PS; i += VJF_NT) slot[i] = 0.f;
    return;
  }
  tc_fence_after();
  // ---- tensor-memory accumulators: A = phi^T phi, b = phi^T dx ; dW1 (+ bias row) ----
  {
    const int g = warp & 3, si = warp >> 2, row = g * 32 + lane;
    for (int un = si; un < (pl.NQ >> 3); un += 4) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(g * 32) << 16) + pl.c_gram + 8 * un, v);
      if (row < R) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int col = 8 * un + jj;
          if (col < R) slot[p.pa + row * R + col] = v[jj];
          else if (col >= pl.Rk && col < pl.Rk + d) slot[p.pb + row * d + (col - pl.Rk)] = v[jj];
        }
      }
    }
    for (int mb = 0; mb < pl.NBLK; ++mb) {
      const int k1 = 128 * mb + row;
      for (int un = si; un < (H >> 3); un += 4) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(g * 32) << 16) + pl.c_dw + mb * H + 8 * un, v);
        float* o = (k1 < p.K1) ? slot + p.lay.mlp_w[0] + (size_t)k1 * H + 8 * un : ((k1 == p.K1) ? slot + p.lay.mlp_b[0] + 8 * un : nullptr);
        if (o) {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
  }
  tc_fence_before();
  // ---- register accumulators through shared-memory scratch, summed in a fixed order ----
  // decoder: [task warp][(d + 1)][32]
  if (warp < pl.NCY * pl.RS) {
    float* o = scr + warp * (DX + 1) * 32;
#pragma unroll
    for (int k = 0; k < DX; ++k) o[k * 32 + lane] = acc.gdw[k];
    o[DX * 32 + lane] = acc.gdb;
  }
  const int HC = pl.HC, hstride = 2 * HC * DX * 32;
  float* scrh = scr + TK_NCW * (DX + 1) * 32;  // heads: [warp][2][HC][DX][32]
  if (warp < TK_NCW) {
    float* o = scrh + warp * hstride;
#pragma unroll
    for (int hc = 0; hc < 4; ++hc)
      if (hc < HC) {
#pragma unroll
        for (int k = 0; k < DX; ++k) { o[(hc * DX + k) * 32 + lane] = acc.ghm[hc][k]; o[((HC + hc) * DX + k) * 32 + lane] = acc.ghv[hc][k]; }
      }
  }
  float* scrs = scrh + TK_NCW * hstride;  // scalars: [warp][NSCAL], then head_v_b [d]
#pragma unroll
  for (int i = 0; i < VJF_NSCAL; ++i) {
    const float s = (warp < TK_NCW) ? warp_sum(acc.sc[i]) : 0.f;
    if (lane == 0) scrs[warp * VJF_NSCAL + i] = s;
  }
  if (tid < d) scrs[16 * VJF_NSCAL + tid] = acc.ghvb;
  __syncthreads();
  for (int i = tid; i < (d + 1) * D; i += VJF_NT) {
    const int k = i / D, j = i - k * D, cy = j >> 5, l = j & 31;
    float s = 0.f;
    for (int rs = 0; rs < pl.RS; ++rs) s += scr[(rs * pl.NCY + cy) * (DX + 1) * 32 + (k < d ? k : DX) * 32 + l];
    if (k < d) slot[p.lay.dec_w + k * D + j] = s;
    else slot[p.lay.dec_b + j] = s;
  }
  for (int i = tid; i < 2 * H * d; i += VJF_NT) {
    const int which = i / (H * d), r = i - which * H * d, n = r / d, k = r - n * d, hc = n >> 5, l = n & 31;
    float s = 0.f;
    for (int w = 0; w < TK_NCW; ++w) s += scrh[w * hstride + ((which * HC + hc) * DX + k) * 32 + l];
    slot[(which ? p.lay.head_v_w : p.lay.head_m_w) + r] = s;
  }
  if (tid < d) slot[p.lay.head_v_b + tid] = scrs[16 * VJF_NSCAL + tid];
  if (tid < VJF_NSCAL) {
    float s = 0.f;
    for (int w = 0; w < TK_NCW; ++w) s += scrs[w * VJF_NSCAL + tid];
    if (tid == 6) slot[p.lay.lik_logvar] = s;  // Gaussian d loss / d lambda (times B)
    else slot[p.ps + tid] = s;
    // a loss-term partial that is not comfortably finite: tell the grid (through barrier 1) to take the exact check
    if (tid < 3 && !(fabsf(s) < 1e30f)) *reinterpret_cast<volatile int*>(sf + pl.f_misc + 2) = 1;
  }
  __syncthreads();
}
