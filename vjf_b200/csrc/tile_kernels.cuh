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
#include "tile_plan.h"

#define TK_NCT (TK_NCW * 32)
#define TK_CTRL 15

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
    if (spins > (1u << 26)) __trap();
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
// control warp: one time step's worth of TMA / tcgen05 issue for the tiles of this CTA.
// Every lane executes the same instruction stream with warp-uniform operands (made explicit with a broadcast shuffle: the
// values arrive through a call boundary), and the asynchronous instructions themselves are predicated on elect.sync --
// operands in uniform registers, no per-instruction broadcast loop.
// ------------------------------------------------------------------------------------------------------------------
struct TkCtl { uint32_t rp, rc, ukn; };  // weight-ring items produced / consumed, UK loads (kernel lifetime)

#define TK_UNI(x) __shfl_sync(0xffffffffu, (x), 0)
__device__ __forceinline__ void tku_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tku_commit(uint32_t bar) {
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tku_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tku_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}\n"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tku_tensor3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n\t}\n"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tku_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 26)) __trap();
  }
}

struct TkU {  // warp-uniform state of the control warp
  uint32_t sbase, bbase, tmem, rp, rc, rbase, rend;
  int t, ntl;
};

static __device__ __forceinline__ void tku_issue_y(const StepParams& p, const CUtensorMap* ymap, const TkU& u, int tile, int buf) {
  const TilePlan& pl = p.tp;
  const uint32_t bar = u.bbase + 8 * (BK_YFULL0 + buf);
  tku_expect(bar, (uint32_t)pl.NCY * pl.TBR * 128u);
  #pragma unroll 1
  for (int c = 0; c < pl.NCY; ++c) tku_tensor3d(u.sbase + pl.o_in[buf] + c * pl.TBR * 128, ymap, c * 32, tile * pl.TBR, u.t, bar);
}
static __device__ __forceinline__ void tku_produce(const StepParams& p, TkU& u) {
  const TilePlan& pl = p.tp;
  const uint32_t item = u.rp, s = item % pl.NS, use = item / pl.NS;
  if (use >= 1) tku_wait(u.bbase + 8 * (BK_WEMPTY0 + s), (use - 1) & 1);
  const uint32_t bytes = 2u * p.H[0] * 128u, bar = u.bbase + 8 * (BK_WFULL0 + s);
  const int chunk = (int)((item - u.rbase) % pl.NCH);
  tku_expect(bar, bytes);
  tku_bulk(u.sbase + pl.o_ring + s * pl.SS, p.w1k + (size_t)chunk * (2 * p.H[0] * 32), bytes, bar);
  ++u.rp;
}

static __device__ void tk_control_step(const StepParams& p, const CUtensorMap* ymap, unsigned char* sb, uint64_t* bars, uint32_t tmem_in, int t_in,
                                       int ntl_in, uint32_t it0_in, TkCtl& cs) {
  const TilePlan& pl = p.tp;
  TkU u;
  u.sbase = TK_UNI(smem_u32(sb)); u.bbase = TK_UNI(smem_u32(bars)); u.tmem = TK_UNI(tmem_in);
  u.rp = TK_UNI(cs.rp); u.rc = TK_UNI(cs.rc);
  u.t = TK_UNI(t_in); u.ntl = TK_UNI(ntl_in);
  uint32_t ukn = TK_UNI(cs.ukn);
  const uint32_t it0 = TK_UNI(it0_in);
  const int H = p.H[0], TBR = pl.TBR, ntl = u.ntl, t = u.t;
  const uint32_t chunkB = (uint32_t)TBR * 128u;
  const uint32_t id_fwd = tk_idesc(128, H, 0, 0), id_quad = tk_idesc(128, pl.NQ, 0, 0), id_gram = tk_idesc(128, pl.NQ, 1, 1), id_dw = tk_idesc(128, H, 1, 1);
  u.rbase = u.rc;
  u.rend = u.rc + (uint32_t)ntl * pl.NCH;
  // the weight images were written by other CTAs through the generic proxy (ordered by the grid barrier / counters)
  asm volatile("fence.proxy.async.global;" ::: "memory");
  #pragma unroll 1
  for (int b = 0; b < pl.NBUF && b < ntl; ++b) tku_issue_y(p, ymap, u, tk_tile_of(b), (int)((it0 + b) % pl.NBUF));
#pragma unroll 1
  while (u.rp < u.rend && u.rp < u.rc + pl.NS) tku_produce(p, u);
  TK_STAMP(p, t, 0, TK_CTRL * 32, 32);
  #pragma unroll 1
  for (int j = 0; j < ntl; ++j) {
    const uint32_t it = it0 + j;
    const int buf = (int)(it % pl.NBUF);
    const uint32_t par = it & 1;
    // ---------------- FWD ----------------
    tku_wait(u.bbase + 8 * BK_CX, par);
    tc_fence_after();
    TK_STAMP(p, t, j, TK_CTRL * 32, 33);
    #pragma unroll 1
    for (int c = 0; c < pl.NCH; ++c) {
      const uint32_t s = u.rc % pl.NS;
      tku_wait(u.bbase + 8 * (BK_WFULL0 + s), (u.rc / pl.NS) & 1);
      tc_fence_after();
      // (lo image: only the chunks from CL0 on exist -- all of them unless the observations are exact in tf32)
      const uint64_t a_hi = tk_kmaj(u.sbase + pl.o_in[buf] + c * chunkB), a_lo = tk_kmaj(u.sbase + pl.o_inlo + (c - pl.CL0) * chunkB);
      const uint64_t b_hi = tk_kmaj(u.sbase + pl.o_ring + s * pl.SS), b_lo = tk_kmaj(u.sbase + pl.o_ring + s * pl.SS + H * 128);
      const int nks = min(4, (pl.K1b - 32 * c + 7) >> 3);
      #pragma unroll 1
      for (int ks = 0; ks < nks; ++ks) {  // a k-step advances 32 bytes inside the 128-byte rows: +2 in the descriptor's address field
        tku_mma(u.tmem + pl.c_d1, a_hi + 2 * ks, b_lo + 2 * ks, id_fwd, (c | ks) ? 1u : 0u);
        if (c >= pl.CL0) tku_mma(u.tmem + pl.c_d1, a_lo + 2 * ks, b_hi + 2 * ks, id_fwd, 1u);
        tku_mma(u.tmem + pl.c_d1, a_hi + 2 * ks, b_hi + 2 * ks, id_fwd, 1u);
      }
      tku_commit(u.bbase + 8 * (BK_WEMPTY0 + s));
      ++u.rc;
      if (u.rp < u.rend) tku_produce(p, u);
    }
    tku_commit(u.bbase + 8 * BK_D1);
    TK_STAMP(p, t, j, TK_CTRL * 32, 34);
    // ---------------- QUAD ----------------
    tku_wait(u.bbase + 8 * BK_CPHI, par);
    tc_fence_after();
    TK_STAMP(p, t, j, TK_CTRL * 32, 35);
    if (j == 0) {
      // [w_chol^T ; w_mean^T] of the previous step: final once the RLS CTA has published it
      if (t > 0) { while (ld_acquire_u32(p.ctrl + 5) < (unsigned)t) __nanosleep(32); }
      asm volatile("fence.proxy.async.global;" ::: "memory");
      tku_expect(u.bbase + 8 * BK_UK, 2u * pl.ukimg);
      tku_bulk(u.sbase + pl.o_uk, p.uk, 2u * pl.ukimg, u.bbase + 8 * BK_UK);
      tku_wait(u.bbase + 8 * BK_UK, ukn & 1);
      ++ukn;
    }
    TK_STAMP(p, t, j, TK_CTRL * 32, 36);
    {
      const uint32_t p_hi = u.sbase + pl.o_pg, p_lo = p_hi + pl.PWC * chunkB;
      const uint32_t u_hi = u.sbase + pl.o_uk, u_lo = u_hi + pl.ukimg;
      const int nch = (pl.Rk + 31) >> 5;
      #pragma unroll 1
      for (int ch = 0; ch < nch; ++ch) {
        const int nks = min(4, (pl.Rk - 32 * ch) >> 3);
        const uint64_t ah = tk_kmaj(p_hi + ch * chunkB), al = tk_kmaj(p_lo + ch * chunkB);
        const uint64_t bh = tk_kmaj(u_hi + ch * pl.NQ * 128), bl = tk_kmaj(u_lo + ch * pl.NQ * 128);
        #pragma unroll 1
        for (int ks = 0; ks < nks; ++ks) {
          tku_mma(u.tmem + pl.c_fl, al + 2 * ks, bh + 2 * ks, id_quad, (ch | ks) ? 1u : 0u);
          tku_mma(u.tmem + pl.c_fl, ah + 2 * ks, bl + 2 * ks, id_quad, 1u);
          tku_mma(u.tmem + pl.c_fl, ah + 2 * ks, bh + 2 * ks, id_quad, 1u);
        }
      }
      tku_commit(u.bbase + 8 * BK_FL);
    }
    TK_STAMP(p, t, j, TK_CTRL * 32, 37);
    // ---------------- GRAM ----------------
    tku_wait(u.bbase + 8 * BK_CPHIT, par);
    tc_fence_after();
    TK_STAMP(p, t, j, TK_CTRL * 32, 38);
    {
      const uint64_t dh = tk_mnmaj(u.sbase + pl.o_pg, chunkB), dl = tk_mnmaj(u.sbase + pl.o_pg + pl.PWC * chunkB, chunkB);
      #pragma unroll 1
      for (int ks = 0; ks < (TBR >> 3); ++ks) {  // a k-step = 8 trials = 1024 bytes: +64 in the address field
        tku_mma(u.tmem + pl.c_gram, dl + 64 * ks, dh + 64 * ks, id_gram, (j | ks) ? 1u : 0u);
        tku_mma(u.tmem + pl.c_gram, dh + 64 * ks, dl + 64 * ks, id_gram, 1u);
        tku_mma(u.tmem + pl.c_gram, dh + 64 * ks, dh + 64 * ks, id_gram, 1u);
      }
      tku_commit(u.bbase + 8 * BK_GRAM);
    }
    TK_STAMP(p, t, j, TK_CTRL * 32, 39);
    // ---------------- DW ----------------
    tku_wait(u.bbase + 8 * BK_CG, par);
    tc_fence_after();
    TK_STAMP(p, t, j, TK_CTRL * 32, 40);
    {
      const uint64_t bh = tk_mnmaj(u.sbase + pl.o_pg, chunkB), bl = tk_mnmaj(u.sbase + pl.o_pg + pl.HC * chunkB, chunkB);
      #pragma unroll 1
      for (int mb = 0; mb < pl.NBLK; ++mb) {
        const uint64_t ah = tk_mnmaj(u.sbase + pl.o_in[buf] + 4 * mb * chunkB, chunkB);
        const uint32_t d = u.tmem + pl.c_dw + mb * H;
#pragma unroll 1
        for (int ks = 0; ks < (TBR >> 3); ++ks) {
          tku_mma(d, ah + 64 * ks, bl + 64 * ks, id_dw, (j | ks) ? 1u : 0u);
          tku_mma(d, ah + 64 * ks, bh + 64 * ks, id_dw, 1u);
        }
      }
      // lo part of the input image (chunks CL0 ..): its own accumulator blocks, aligned with the blocks CL0 / 4 .. of the hi part
      // (the descriptor may start before the lo image; rows of chunks without a lo image multiply whatever lies there and
      // are never read back)
#pragma unroll 1
      for (int mb = 0; mb < pl.NBLKLO; ++mb) {
        const uint64_t al = tk_mnmaj(u.sbase + pl.o_inlo + (4 * ((pl.CL0 >> 2) + mb) - pl.CL0) * chunkB, chunkB);
        const uint32_t d = u.tmem + pl.c_dwlo + mb * H;
#pragma unroll 1
        for (int ks = 0; ks < (TBR >> 3); ++ks) tku_mma(d, al + 64 * ks, bh + 64 * ks, id_dw, (j | ks) ? 1u : 0u);
      }
      tku_commit(u.bbase + 8 * BK_DW);
      TK_STAMP(p, t, j, TK_CTRL * 32, 41);
      // the input buffer is free once these MMAs are done: observations of the tile after next
      if (j + pl.NBUF < ntl) {
        tku_wait(u.bbase + 8 * BK_DW, par);
        tku_issue_y(p, ymap, u, tk_tile_of(j + pl.NBUF), buf);
      }
    }
  }
  cs.rp = u.rp; cs.rc = u.rc; cs.ukn = ukn;
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

// tanh from one exponential: |abs err| < 1.5e-7 (the cancellation in t - 1 for small x costs relative, not absolute accuracy)
__device__ __forceinline__ float tk_tanh(float x) {
  const float xc = fminf(fmaxf(x, -15.f), 15.f);
  const float t = __expf(2.0f * xc);
  const float r = __fdividef(t - 1.0f, t + 1.0f);
  return (x != x) ? x : r;
}

#define TK_LK_RB 4  // rows per batch of the likelihood stage (independent chains in flight per warp)

// Likelihood stage of one warp: observation chunk cy (a lane owns one column), trials rs, rs + RS, ...  TK_LK_RB trials are in
// flight together; the row code is branch-free (selects) so that their chains interleave.
template <int DX, int LIK>
static __device__ __forceinline__ void tk_lk_rows(const StepParams& p, TkAcc<DX>& acc, const unsigned char* in_b, const float* xt_s, float* gxp_s,
                                                   const float* dec, int cy, int rs, int nb, bool r_on, float lam, float e_nlam, float p_lam) {
  const TilePlan& pl = p.tp;
  const int lane = threadIdx.x & 31, d = p.d, D = p.D, TBR = pl.TBR;
  const int jcol = 32 * cy + lane;
  const bool jok = jcol < D;
  float w[DX];
  const float bj = jok ? dec[d * D + jcol] : 0.f;
#pragma unroll
  for (int k = 0; k < DX; ++k) w[k] = (jok && (DX == d || k < d)) ? dec[k * D + jcol] : 0.f;
#pragma unroll 1
  for (int bb = rs; bb < nb; bb += TK_LK_RB * pl.RS) {
    float gx[TK_LK_RB][DX];
#pragma unroll
    for (int q = 0; q < TK_LK_RB; ++q) {
      const int b = bb + q * pl.RS;
      const bool on = jok && b < nb;
      const int bc = min(b, nb - 1);  // loads of a masked row stay in range
      float xt[DX], eta = bj;
#pragma unroll
      for (int k = 0; k < DX; ++k) { xt[k] = (DX == d || k < d) ? xt_s[bc * d + k] : 0.f; eta = fmaf(w[k], xt[k], eta); }
      const float yv = *reinterpret_cast<const float*>(in_b + b32_off(bc, jcol, TBR));
      float g;
      if (LIK == VJF_LIK_GAUSSIAN) {
        // gaussian_loss(y, eta, lambda), functional.py:55-75 ; update's mse, likelihood.py:36-37
        const float r = yv - eta;
        const float rsd = yv * p_lam - eta * p_lam;
        const float mse = rsd * rsd;
        acc.sc[SC_BADMSE] += (on && !isfinite(mse)) ? 1.f : 0.f;
        acc.sc[SC_RECON] += on ? 0.5f * (mse + lam) : 0.f;
        acc.sc[SC_SSE] += on ? r * r : 0.f;
        g = -r * e_nlam;
        acc.sc[6] += (on && r_on) ? 0.5f * (1.0f - r * r * e_nlam) : 0.f;
      } else {
        // poisson_nll_loss(clamp(eta, max=10), y, log_input=True), likelihood.py:60-62; NaN propagates like torch.clamp
        const float ec = fminf(eta, 10.0f);
        const float ex = __expf(ec);
        const bool isn = eta != eta;
        acc.sc[SC_RECON] += on ? (isn ? eta : ex - yv * ec) : 0.f;
        g = isn ? eta : ((eta <= 10.0f) ? (ex - yv) : 0.f);
      }
      g = (on && r_on) ? g : 0.f;
      acc.gdb += g;
#pragma unroll
      for (int k = 0; k < DX; ++k) { acc.gdw[k] = fmaf(g, xt[k], acc.gdw[k]); gx[q][k] = g * w[k]; }
    }
    // transposed warp reduction of the TK_LK_RB x DX partial sums: at every step a lane keeps one half of its values and hands
    // the other half to its partner, so N values over 32 lanes cost N - 1 + (5 - log2 N) shuffles instead of 5 N; lane L ends
    // with the full sum of value L >> (5 - log2 N) and stores it itself
    {
      constexpr int DXP = DX <= 2 ? 2 : (DX <= 4 ? 4 : 8), N = TK_LK_RB * DXP, LOGN = DXP == 2 ? 3 : (DXP == 4 ? 4 : 5);
      float v[N];
#pragma unroll
      for (int q = 0; q < TK_LK_RB; ++q)
#pragma unroll
        for (int k = 0; k < DXP; ++k) v[q * DXP + k] = (k < DX) ? gx[q][k < DX ? k : 0] : 0.f;
      int o = 16;
#pragma unroll
      for (int h = N / 2; h >= 1; h >>= 1, o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float keep = up ? v[i + h] : v[i], send = up ? v[i] : v[i + h];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
#pragma unroll
      for (; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
      if ((lane & ((32 >> LOGN) - 1)) == 0) {
        const int idx = lane >> (5 - LOGN), q = idx / DXP, k = idx - q * DXP;
        const int b = bb + q * pl.RS;
        if (b < nb && k < d) gxp_s[(cy * TBR + b) * d + k] = v[0];
      }
    }
  }
}


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

  TK_STAMP(p, t, j, 0, 0);
  // ---- A1: previous posterior, control input, noise; xs = m_s + eps1 exp(l_s / 2) (vjf/util.py:11-13) ----
  {
    const size_t row0 = (size_t)t * p.B + b0;
    const bool prior = (t == 0) && (p.flags & VJF_FLAG_PRIOR_Q0);
    const float* qm = (t == 0) ? p.q0m : p.mu + (size_t)(t - 1) * p.B * d;
    const float* ql = (t == 0) ? p.q0l : p.logvar + (size_t)(t - 1) * p.B * d;
    // one thread per (trial, state dimension): its four inputs are independent loads
    #pragma unroll 1
    for (int i = ctid; i < TBR * d; i += TK_NCT) {
      const int b = i / d, k = i - b * d;
      float ms = 0.f, ls = 0.f, e1 = 0.f, e2 = 0.f;
      if (b < nb) {
        if (prior) { ms = st[p.lay.prior_mean + k]; ls = st[p.lay.prior_logvar + k]; }
        else { ms = qm[(size_t)(b0 + b) * d + k]; ls = ql[(size_t)(b0 + b) * d + k]; }
        if (p.eps) {
          const float* e0 = p.eps + ((size_t)t * 2 * p.B + b0 + b) * d;
          e1 = e0[k]; e2 = e0[(size_t)p.B * d + k];
        } else if ((k & 3) == 0) {
          float z1[4], z2[4];
          philox_normal4(p.seed, p.step0 + t, p.trial_offset + b0 + b, 0, k >> 2, z1);
          philox_normal4(p.seed, p.step0 + t, p.trial_offset + b0 + b, 1, k >> 2, z2);
          e1 = z1[0]; e2 = z2[0];
          for (int q = 1; q < 4 && k + q < d; ++q) { eps_s[b * 2 * d + k + q] = z1[q]; eps_s[b * 2 * d + d + k + q] = z2[q]; }
        }
      }
      ex_s[b * E + u + k] = ms; ex_s[b * E + u + d + k] = ls;
      if (p.eps || (k & 3) == 0 || b >= nb) { eps_s[b * 2 * d + k] = e1; eps_s[b * 2 * d + d + k] = e2; }
    }
    #pragma unroll 1
    for (int i = ctid; i < TBR * u; i += TK_NCT) {
      const int b = i / u, e = i - b * u;
      ex_s[b * E + e] = (b < nb) ? p.u_in[(row0 + b) * u + e] : 0.f;
    }
    cb_sync();
    TK_STAMP(p, t, j, 0, 1);
    #pragma unroll 1
    for (int i = ctid; i < TBR * du; i += TK_NCT) {
      const int b = i / du, k = i - b * du;
      float v;
      if (k < d) v = ex_s[b * E + u + k] + eps_s[b * 2 * d + k] * __expf(0.5f * ex_s[b * E + u + d + k]);
      else v = ex_s[b * E + (k - d)];
      xu_s[i] = v;
    }
    cb_sync();
    // the phi / g_pre region is free once the weight-gradient MMAs of the previous tile are done
    TK_STAMP(p, t, j, 0, 2);
    if (it > 0) tk_wait(&bars[BK_DW], (it - 1) & 1);
    TK_STAMP(p, t, j, 0, 3);
    // observations of this tile (TMA) -> [u | m_s | l_s | 1] appended behind them (vjf/recognition.py:32-37; the ones column
    // carries the bias through the MMAs) and the lo image of the whole input tile, in one pass over the 16-byte pieces
    tk_wait(&bars[BK_YFULL0 + buf], (it / pl.NBUF) & 1);
    TK_STAMP(p, t, j, 0, 4);
    {
      float4* img = reinterpret_cast<float4*>(in_b);
      float4* dst = reinterpret_cast<float4*>(inlo_b);
      const int c_x = D >> 5;  // first chunk that holds appended columns
      // exact observations (CL0 = c_x): the chunks before c_x need neither the appended columns nor a lo image
#pragma unroll 1
      for (int i = pl.CL0 * TBR * 8 + ctid; i < pl.NCH * TBR * 8; i += TK_NCT) {
        const int c = i / (TBR * 8), rr = (i >> 3) % TBR;
        float4 x;
        if (c < c_x) {
          x = img[i];
        } else {
          const int col0 = 32 * c + (((i & 7) ^ (rr & 7)) << 2);  // logical columns of this physical piece
          float v[4];
          if (c < pl.NCY && col0 + 3 < D) { x = img[i]; v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
          else {
            if (c < pl.NCY) { x = img[i]; v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; } else { v[0] = v[1] = v[2] = v[3] = 0.f; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int e = col0 + q - D;
              if (e >= 0) v[q] = (rr < nb) ? (e < E ? ex_s[rr * E + e] : (e == E ? 1.0f : 0.f)) : 0.f;
            }
            x = make_float4(v[0], v[1], v[2], v[3]);
            img[i] = x;
          }
        }
        dst[i - pl.CL0 * TBR * 8] = make_float4(x.x - tf32_trunc_f(x.x), x.y - tf32_trunc_f(x.y), x.z - tf32_trunc_f(x.z), x.w - tf32_trunc_f(x.w));
      }
    }
    tk_signal(&bars[BK_CX]);
    TK_STAMP(p, t, j, 0, 5);
    // (the features are computed while the forward MMAs of this tile run: the quadratic form needs them only after the RLS
    // outputs of the previous step have arrived)
    // phi = exp(-0.5 |xu - c|^2 / w^2) (vjf/functional.py:11-22) as a (hi, lo) pair of K-major SW128 images; pad columns and
    // pad rows are zero
    unsigned char* ph = pg;
    unsigned char* plo = pg + pl.PWC * TBR * 128;
    #pragma unroll 1
    for (int i = ctid; i < TBR * pl.PW; i += TK_NCT) {
      const int b = i / pl.PW, r = i - b * pl.PW;
      float v = 0.f;
      if (b < nb && r < R) {
        float d2 = 0.f;
        #pragma unroll 1
        for (int c = 0; c < du; ++c) { const float df = xu_s[b * du + c] - c_s[r * du + c]; d2 = fmaf(df, df, d2); }
        v = __expf(d2 * iw_s[r]);
      }
      const int o = sw128_off(b, r, TBR);
      *reinterpret_cast<float*>(ph + o) = v;
      *reinterpret_cast<float*>(plo + o) = v - tf32_trunc_f(v);
    }
    tk_signal(&bars[BK_CPHI]);
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
      #pragma unroll 1
      for (int un = si; un < (H >> 4); un += nsw) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(g * 32) << 16) + pl.c_d1 + 16 * un, v);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) hs[row * ldhs + 16 * un + jj] = tk_tanh(v[jj]);
      }
    }
  }
  tc_fence_before();
  cb_sync();
  TK_STAMP(p, t, j, 0, 8);
  // heads: eight threads per (trial, state dimension), each over every eighth hidden unit; xor-shuffle reduction
  #pragma unroll 1
  for (int i0 = 0; i0 < TBR * d * 8; i0 += TK_NCT) {
    const int i = i0 + ctid, nq = i & 7, bk = i >> 3;
    const bool ok = bk < TBR * d;
    const int b = ok ? bk / d : 0, k = ok ? bk - b * d : 0;
    float m = 0.f, lv = 0.f;
    if (ok && b < nb) {
      const float* hr = hs + b * ldhs;
      #pragma unroll 1
      for (int n = nq; n < H; n += 8) { const float h = hr[n]; m = fmaf(h, hm_s[n * d + k], m); lv = fmaf(h, hv_s[n * d + k], lv); }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { m += __shfl_xor_sync(0xffffffffu, m, o); lv += __shfl_xor_sync(0xffffffffu, lv, o); }
    if (ok && nq == 0) {
      float x = 0.f, dxv = 0.f;
      if (b < nb) {
        lv += hv_s[H * d + k];
        x = m + eps_s[b * 2 * d + d + k] * __expf(0.5f * lv);
        dxv = x - xu_s[b * du + k];
        acc.sc[SC_SDX] = fmaf(dxv, dxv, acc.sc[SC_SDX]);
        p.mu[((size_t)t * p.B + b0 + b) * d + k] = m;       // posterior of this step (model.py:218-221, :305-307)
        p.logvar[((size_t)t * p.B + b0 + b) * d + k] = lv;
      } else { m = 0.f; lv = 0.f; }
      mt_s[bk] = m; lt_s[bk] = lv; xt_s[bk] = x; dx_s[bk] = dxv;
    }
  }
  TK_STAMP(p, t, j, 0, 9);
  // ---- the input tile becomes the MN-major operand of the weight gradient: permute every 128-byte row in place ----
  #pragma unroll 1
  for (int i = ctid; i < (2 * pl.NCH - pl.CL0) * TBR; i += TK_NCT) {
    const int im = i >= pl.NCH * TBR, rr = i - im * (pl.NCH * TBR);
    tk_permute_row((im ? inlo_b : in_b) + rr * 128, rr % TBR);
  }
  cb_sync();
  TK_STAMP(p, t, j, 0, 10);

  // ---- LK: decoder eta = D xt + bias (model.py:29-30), likelihood terms, d loss / d eta (times B), decoder gradients, g_xt.
  //      A warp owns one 32-column chunk of the observations for a subset of the trials, a lane one column; TK_LK_RB trials are
  //      in flight together so that the cross-lane sums of g_xt are independent shuffle chains ----
  {
    float lam = 0.f;
    if (p.lik == VJF_LIK_GAUSSIAN) {
      if (t > 0) tk_wait_counter(p.ctrl + 4, (unsigned)t);
      lam = __ldcg(st + p.lay.lik_logvar);
    }
    const float e_nlam = __expf(-lam), p_lam = __expf(-0.5f * lam);
    if (cw < pl.NCY * pl.RS) {
      if (p.lik == VJF_LIK_GAUSSIAN) tk_lk_rows<DX, VJF_LIK_GAUSSIAN>(p, acc, in_b, xt_s, gxp_s, dec, cw % pl.NCY, cw / pl.NCY, nb, r_on, lam, e_nlam, p_lam);
      else tk_lk_rows<DX, VJF_LIK_POISSON>(p, acc, in_b, xt_s, gxp_s, dec, cw % pl.NCY, cw / pl.NCY, nb, r_on, lam, e_nlam, p_lam);
    }
    cb_sync();
    #pragma unroll 1
    for (int i = ctid; i < TBR * d; i += TK_NCT) {
      const int b = i / d, k = i - b * d;
      float s = 0.f;
      if (b < nb)
        #pragma unroll 1
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
  #pragma unroll 1
  for (int i = ctid; i < 2 * pl.PWC * TBR; i += TK_NCT) tk_permute_row(pg + i * 128, i % TBR);
  // FL = phi [w_chol | w_mean] from tensor memory: p_logvar = log |phi w_chol|^2, p_mean = xs + phi W (module.py:75-77,
  // model.py:338); the warps of a lane group share its columns, the partial sums of squares meet in gxp_s
  {
    const int g = cw & 3, si = cw >> 2, nsw = (g == 3) ? 3 : 4;
    if (g * 32 < TBR) {
      const int row = g * 32 + lane;
      float q = 0.f;
      #pragma unroll 1
      for (int un = si; un < (pl.NQ >> 4); un += nsw) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(g * 32) << 16) + pl.c_fl + 16 * un, v);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int col = 16 * un + jj;
          if (col < pl.Rk) q = fmaf(v[jj], v[jj], q);
          else if (col < pl.Rk + d) pm_s[row * d + col - pl.Rk] = xu_s[row * du + col - pl.Rk] + v[jj];
        }
      }
      gxp_s[pl.NCY * TBR * d + si * TBR + row] = q;
    }
  }
  cb_sync();
  {
    unsigned char* ph = pg;
    unsigned char* plo = pg + pl.PWC * TBR * 128;
    #pragma unroll 1
    for (int i = ctid; i < TBR * d; i += TK_NCT) {
      const int b = i / d, k = i - b * d;
      const float v = dx_s[i];
      const int o = b32_off(b, pl.Rk + k, TBR);
      *reinterpret_cast<float*>(ph + o) = v;
      *reinterpret_cast<float*>(plo + o) = v - tf32_trunc_f(v);
    }
    if (ctid < TBR) {
      const int g = ctid >> 5, nsw = (g == 3) ? 3 : 4;
      float q = 0.f;
      #pragma unroll 1
      for (int si = 0; si < nsw && si < (pl.NQ >> 4); ++si) q += gxp_s[pl.NCY * TBR * d + si * TBR + ctid];
      plv_s[ctid] = __logf(q);
    }
  }
  tk_signal(&bars[BK_CPHIT]);
  TK_STAMP(p, t, j, 0, 13);
  // state-noise logvar of the previous step: published by the RLS CTA after its tail
  if (t > 0) tk_wait_counter(p.ctrl + 3, (unsigned)t);
  else cb_sync();
  TK_STAMP(p, t, j, 0, 14);
  {
    const float gam = __ldcg(st + p.lay.tr_logvar);
    const float e_ngam = __expf(-gam), p_gam = __expf(-0.5f * gam);
    // dynamics NLL (functional.py:55-75 via model.py:390-391), entropy (functional.py:25-29), g_mt and g_lt (times B)
    #pragma unroll 1
    for (int i = ctid; i < TBR * d; i += TK_NCT) {
      const int b = i / d, k = i - b * d;
      float gm = 0.f, gl = 0.f;
      if (b < nb) {
        const float m = mt_s[i], lv = lt_s[i], pm = pm_s[i], plv = plv_s[b];
        const float e2 = eps_s[b * 2 * d + d + k], gx = gxt_s[i];
        const float df = pm * p_gam - m * p_gam;
        const float mse = df * df;
        if (!isfinite(mse)) acc.sc[SC_BADMSE] += 1.f;
        const float tr = __expf(plv + lv - gam);
        acc.sc[SC_DYN] += 0.5f * (mse + gam) + 0.5f * tr;
        acc.sc[SC_ENT] += 0.5f * lv;
        gm = gx; gl = 0.5f * gx * e2 * __expf(0.5f * lv);
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
    #pragma unroll 1
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
      #pragma unroll 1
      for (int b = 0; b < nb; ++b) s += glt_s[b * d + ctid];
      acc.ghvb += s;
    }
  }
  tk_signal(&bars[BK_CG]);
  TK_STAMP(p, t, j, 0, 17);
}

// Flush of everything a CTA accumulated over its tiles of one time step into its slot (all 16 warps).
template <int DX>
static __device__ void tk_flush_step(const StepParams& p, unsigned char* sb, uint32_t tmem, TkAcc<DX>& acc, bool have_tiles, int t = 0) {
  const TilePlan& pl = p.tp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D, d = p.d, R = p.R, H = p.H[0];
  float* slot = p.partials + (size_t)blockIdx.x * p.PS;
  float* sf = reinterpret_cast<float*>(sb + pl.o_f);
  float* scr = reinterpret_cast<float*>(sb + pl.o_in[0]);  // scratch: everything up to the [w_chol ; w_mean] images is dead between steps
  if (!have_tiles) {
    #pragma unroll 1
    for (int i = tid; i < p.PS; i += VJF_NT) slot[i] = 0.f;
    return;
  }
  tc_fence_after();
  // ---- tensor-memory accumulators: A = phi^T phi, b = phi^T dx ; dW1 (+ bias row).  A thread reads a ROW of an accumulator
  //      (tensor-memory lane = row), so storing from the registers would touch 32 different lines per instruction; the blocks go
  //      through a shared-memory stage instead and leave as whole 128-byte lines ----
  if (pl.flush_stage) {
    const int g = warp & 3, si = warp >> 2, row = g * 32 + lane;
    float* stg = scr + pl.flush_floats;            // [128][LD]
    const int LDg = pl.NQ + 4, LDw = H + 4;
    // Gram block
    #pragma unroll 1
    for (int un = si; un < (pl.NQ >> 4); un += 4) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(g * 32) << 16) + pl.c_gram + 16 * un, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(stg + row * LDg + 16 * un + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    __syncthreads();
    #pragma unroll 1
    for (int i = tid; i < R * R; i += VJF_NT) { const int r = i / R, c = i - r * R; slot[p.pa + i] = stg[r * LDg + c]; }
    #pragma unroll 1
    for (int i = tid; i < R * d; i += VJF_NT) { const int r = i / d, k = i - r * d; slot[p.pb + i] = stg[r * LDg + pl.Rk + k]; }
    #pragma unroll 1
    for (int mb = 0; mb < pl.NBLK; ++mb) {
      __syncthreads();
      #pragma unroll 1
      for (int un = si; un < (H >> 4); un += 4) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(g * 32) << 16) + pl.c_dw + mb * H + 16 * un, v);
        if (mb >= (pl.CL0 >> 2)) {  // + the lo part of the input image (same rows of its own accumulator block)
          float vl[16];
          tmem_ld16(tmem + ((uint32_t)(g * 32) << 16) + pl.c_dwlo + (mb - (pl.CL0 >> 2)) * H + 16 * un, vl);
          if (128 * mb + row >= 32 * pl.CL0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] += vl[q];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(stg + row * LDw + 16 * un + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      __syncthreads();
      const int nvalid = min(128, p.K1 - 128 * mb);  // weight rows of this block (the bias row K1 follows them when it is in the block)
      const int h4 = H >> 2;
      float4* dst = reinterpret_cast<float4*>(slot + p.lay.mlp_w[0] + (size_t)(128 * mb) * H);
      #pragma unroll 1
      for (int i = tid; i < nvalid * h4; i += VJF_NT) { const int r = i / h4, c4 = i - r * h4; dst[i] = *reinterpret_cast<const float4*>(stg + r * LDw + 4 * c4); }
      const int rb = p.K1 - 128 * mb;
      if (rb >= 0 && rb < 128 && tid < H) slot[p.lay.mlp_b[0] + tid] = stg[rb * LDw + tid];
    }
  }
  else {  // (the staging block does not fit the scratch of this plan: rows straight from the registers)
    const int g = warp & 3, si = warp >> 2, row = g * 32 + lane;
    #pragma unroll 1
    for (int un = si; un < (pl.NQ >> 4); un += 4) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(g * 32) << 16) + pl.c_gram + 16 * un, v);
      if (row < R) {
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int col = 16 * un + jj;
          if (col < R) slot[p.pa + row * R + col] = v[jj];
          else if (col >= pl.Rk && col < pl.Rk + d) slot[p.pb + row * d + (col - pl.Rk)] = v[jj];
        }
      }
    }
    #pragma unroll 1
    for (int mb = 0; mb < pl.NBLK; ++mb) {
      const int k1 = 128 * mb + row;
      #pragma unroll 1
      for (int un = si; un < (H >> 4); un += 4) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(g * 32) << 16) + pl.c_dw + mb * H + 16 * un, v);
        if (mb >= (pl.CL0 >> 2)) {  // + the lo part of the input image (same rows of its own accumulator block)
          float vl[16];
          tmem_ld16(tmem + ((uint32_t)(g * 32) << 16) + pl.c_dwlo + (mb - (pl.CL0 >> 2)) * H + 16 * un, vl);
          if (k1 >= 32 * pl.CL0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] += vl[q];
          }
        }
        float* o = (k1 < p.K1) ? slot + p.lay.mlp_w[0] + (size_t)k1 * H + 16 * un : ((k1 == p.K1) ? slot + p.lay.mlp_b[0] + 16 * un : nullptr);
        if (o) {
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(o + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
  }
  tc_fence_before();
  TK_STAMP(p, t, 0, 0, 54);
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
  TK_STAMP(p, t, 0, 0, 55);
  #pragma unroll 1
  for (int i = tid; i < (d + 1) * D; i += VJF_NT) {
    const int k = i / D, j = i - k * D, cy = j >> 5, l = j & 31;
    float s = 0.f;
    #pragma unroll 1
    for (int rs = 0; rs < pl.RS; ++rs) s += scr[(rs * pl.NCY + cy) * (DX + 1) * 32 + (k < d ? k : DX) * 32 + l];
    if (k < d) slot[p.lay.dec_w + k * D + j] = s;
    else slot[p.lay.dec_b + j] = s;
  }
  #pragma unroll 1
  for (int i = tid; i < 2 * H * d; i += VJF_NT) {
    const int which = i / (H * d), r = i - which * H * d, n = r / d, k = r - n * d, hc = n >> 5, l = n & 31;
    float s = 0.f;
    #pragma unroll 1
    for (int w = 0; w < TK_NCW; ++w) s += scrh[w * hstride + ((which * HC + hc) * DX + k) * 32 + l];
    slot[(which ? p.lay.head_v_w : p.lay.head_m_w) + r] = s;
  }
  if (tid < d) slot[p.lay.head_v_b + tid] = scrs[16 * VJF_NSCAL + tid];
  if (tid < VJF_NSCAL) {
    float s = 0.f;
    #pragma unroll 1
    for (int w = 0; w < TK_NCW; ++w) s += scrs[w * VJF_NSCAL + tid];
    if (tid == 6) slot[p.lay.lik_logvar] = s;  // Gaussian d loss / d lambda (times B)
    else slot[p.ps + tid] = s;
    // a loss-term partial that is not comfortably finite: tell the grid (through barrier 1) to take the exact check
    if (tid < 3 && !(fabsf(s) < 1e30f)) *reinterpret_cast<volatile int*>(sf + pl.f_misc + 2) = 1;
  }
  __syncthreads();
}
