// Shared device/host definitions for the vjf_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/vjf_b200.h"

#define VJF_NT 512               // threads per CTA of the step kernels
#define VJF_NWARP (VJF_NT / 32)
#define VJF_BIGR_MIN 160           // n_rbf above this runs the large-n_rbf path (bigr.cu)
#define VJF_TB_MAX 32            // trials per tile (rows of the per-CTA tile)
#define VJF_NSCAL 8              // scalar sums carried in the partial vector
#define TK_NCW 15                // compute warps of the tile pipeline (warp 15 is its control warp)

// scalar slots
#define SC_RECON 0   // sum_b sum_j recon terms (without the 0.5*lambda*D constant)
#define SC_DYN 1     // sum_b sum_i dyn terms (without 0.5*gamma*d)
#define SC_ENT 2     // sum_b sum_i 0.5*l_t
#define SC_SSE 3     // sum (y-eta)^2            (Gaussian likelihood.update)
#define SC_SDX 4     // sum dx^2                 (state-noise residual algebra)
#define SC_BADMSE 5  // count of non-finite squared errors (functional.py:60 assert)

struct Lay {  // int copies of vjf_layout (state buffers are < 2^31 floats)
  int lik_logvar, dec_w, dec_b, mlp_w[VJF_MAX_LAYERS], mlp_b[VJF_MAX_LAYERS], head_m_w, head_v_w, head_v_b, n_train;
  int prior_mean, prior_logvar, tr_logvar, centroid, logwidth, w_mean, w_chol, w_precision, w_pchol, lik_n, tr_n, total;
};

// Plan of the throughput tile pipeline (tile_kernels.cuh): tiles of TBR trials, every contraction on tcgen05 from
// shared-memory images, observations streamed by TMA tensor copies.  Filled by the host (tile_host.cu).
struct TilePlan {
  int on;                    // this launch runs vjf_tile_kernel
  int TBR, NBUF;             // trials per tile, observation buffers (double buffering)
  int K1b, NCH, NCY;         // K1 + 1 (ones column = bias), 32-column chunks of the input image, chunks with observation columns
  int Rk, NQ, PW, PWC;       // K of the quadratic form (roundup(R, 8)), N of QUAD / GRAM, phi image width (columns, chunks)
  int HC, NBLK;              // H / 32, 128-row blocks of the layer-1 weight-gradient accumulator
  int CL0, NBLKLO, c_dwlo;   // first chunk of the input image with a lo image (0 unless the observations are exact in tf32), blocks and
                             // tensor-memory column of the lo part of the weight gradient
  int NS, SS;                // weight ring: stages, bytes per stage
  int ukimg;                 // bytes of one image (hi or lo) of the [w_chol^T ; w_mean^T] operand
  int RS;                    // row splits of the likelihood stage (warps per observation chunk)
  // shared memory: byte offsets from the 1024-byte aligned base
  int o_in[2], o_inlo, o_pg, o_ring, o_uk, o_f, scratch_bytes;
  int flush_floats, flush_stage;  // floats of the per-step flush scratch; 1: a [128][N + 4] staging block behind it (accumulator flush by whole lines)
  // float area: float offsets from o_f
  int f_hs, ldhs, f_dec, f_hm, f_hv, f_cen, f_iw, f_ex, f_eps, f_xu, f_xt, f_mt, f_lt, f_pm, f_dx, f_gxt, f_gmt, f_glt, f_plv,
      f_gxp, f_red, f_misc, f_bar, f_total;
  int smem_bytes;
  int c_d1, c_fl, c_gram, c_dw;  // tensor-memory columns of the accumulators
};

struct StepParams {
  // ---- throughput tile pipeline ----
  TilePlan tp;
  float* w1k;   // recognition layer-1 weight as tcgen05 B-operand images (per 32-row chunk: hi | lo, K-major SW128), kept by the SGD step
  float* uk;    // [w_chol^T ; w_mean^T] as B-operand images (hi | lo), kept by the RLS commit
  // ---- dimensions ----
  int B, Bglobal, D, d, u, R, L, H[VJF_MAX_LAYERS];
  int K1, K1p, E, du, Dp, Rp, Hp[VJF_MAX_LAYERS], Hpmax;
  int Gp, ldu;         // row strides of the g_pre buffers and of the staged w_chol
  int lik;
  Lay lay;
  // ---- partial-sum vector layout: [0,G) grads | A (R*R) | b (R*d) | scalars ----
  int G, pa, pb, ps, PS;
  // ---- tiling ----
  int TB, ntiles, nslots;
  // ---- shared-memory plan (float offsets) ----
  int s_in, s_g, s_phi, s_act[VJF_MAX_LAYERS], s_gpa, s_gpb, s_eps, s_xu, s_xt, s_mt, s_lt, s_pm, s_dx, s_gxt, s_gmt,
      s_glt, s_plv, s_U, s_W, s_c, s_iw, s_red, s_dec, s_qp, s_W1, s_hm, s_hv, s_flag, s_scf, s_b1, s_inl, s_phil, s_gpal, s_gpbl, s_total;
  int U_in_smem, dec_in_smem, W1_in_smem, ldw1;
  int in_split;  // the staged input matrix is kept as a presplit (hi, lo) pair (needs a second rows x K1p array)
  int dbg_cta;   // development aid: which trial CTA drops the phase-A stamps
  float* u_mirror;   // [roundup(R,8)][ldu] row-padded copy of w_chol (workspace), the TMA source of the back half
  float* w1_mirror;  // [K1][ldw1] row-padded copy of the recognition layer-1 weight (workspace), the TMA source
  int use_tma;   // TMA bulk staging of the layer-1 weight / decoder and prefetch of the next observation tile (overlapped schedule)
  int use_umma;  // layer-1 weight gradient on tcgen05 (umma.cuh): overlapped schedule, one hidden layer of <= 64 units
  int umma_nk;   // roundup(K1, 8): the N extent of that MMA
  int ldm;  // row stride of the factorisation workspace in phase B2
  // large n_rbf (> 128, bigr.cu): RBF features, dynamics read-out and RLS statistics live outside the tile kernels
  int ext; const float* ext_xs; const float* ext_pm; const float* ext_plv; float* ext_dx;
  int rls64;    // RLS in double precision (long runs): w_precision shadowed in P64, sweep in double
  double* P64;  // [R][R]
  // ---- pointers ----
  float* state;
  float* partials;      // [nslots][PS]
  float* reduced;       // [PS]
  unsigned* barrier;    // grid barrier counter
  unsigned* status;     // status word
  unsigned* ctrl;       // [0] term masks of the current step (persistent redo)
  const void* y;
  int y_dtype;
  const float* u_in;
  const float* q0m;
  const float* q0l;
  const float* eps;
  float* mu;
  float* logvar;
  float* losses;
  unsigned long long seed, step0, trial_offset;
  unsigned flags;
  float lr;
  int T;
  int red_begin;   // first element of the partial vector the reduction touches
  long long* dbg;  // optional [T][8] per-phase globaltimer stamps of CTA 0 (development aid)
  // ---- sharded run: exchange over peer memory ----
  int world, rank;               // world == 1: single GPU
  float* peer[VJF_MAX_RANKS];    // exchange buffer of every rank (peer-mapped); [world][2][PSx] floats then flags
  int PSx;                       // PS rounded up to a multiple of 128
  unsigned epoch0;               // epoch of time step 0 of this launch is epoch0 + 1
  int overlap;     // persistent schedule: CTA 0 is the dedicated RLS CTA, trial CTAs run the front half of step t+1 beside it
  int init_mode;   // phase B2 runs as RBFDS.initialize (vjf/model.py:379-388) instead of a filter step
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid-wide barrier for the persistent (cooperatively launched) kernel.  `target` is a per-thread
// running count of expected arrivals; the counter is zeroed by the host before each launch.
// Grid barrier that also carries one bit of information per arriving CTA: the counter is 64 bits wide, arrivals count in the
// low word, "my loss partials are not comfortably finite" arrivals in the high word (cumulative over the launch).  The
// flag count is read with the same acquire load that observes the barrier, so the decision "take the exact finite
// check this step" costs no extra L2 round trip.  Returns the cumulative flag count (identical on every CTA: the next
// flagged arrival can only happen after every CTA has left this barrier and a later one).
__device__ __forceinline__ unsigned grid_barrier_flag(unsigned long long* counter, unsigned& target, bool my_flag, unsigned* bcast) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    const unsigned long long inc = 1ull + (my_flag ? (1ull << 32) : 0ull);
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(inc) : "memory");
    unsigned long long v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
      if ((unsigned)v >= target) break;
      __nanosleep(32);
    }
    *bcast = (unsigned)(v >> 32);
    __threadfence();  // gpu-scope fence: also drops stale L1 lines before the CTA reads peers' data
  }
  __syncthreads();
  return *bcast;
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, unsigned nparticipants = 0) {
  __syncthreads();
  target += nparticipants ? nparticipants : gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    red_release_add_u32(counter, 1u);
    while (ld_acquire_u32(counter) < target) __nanosleep(32);  // back off: ~130 pollers share one L2 line with the arriving atomics
    __threadfence();  // gpu-scope fence: also drops stale L1 lines before the CTA reads peers' data
  }
  __syncthreads();
}

// system-scope (cross-GPU) flag accessors for the in-kernel exchange
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// Block-wide wait until a monotonically increasing device counter reaches `want` (producer side: fence + atomicAdd).
__device__ __forceinline__ void wait_counter(const unsigned* counter, unsigned want) {
  if (threadIdx.x == 0) {
    while (ld_acquire_u32(counter) < want) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

// Philox4x32-10 (Salmon et al. 2011) -> 4 x N(0,1) by Box-Muller.  Keyed by (seed, step); counter =
// (trial index, which draw, 4-block of the state dimension).  Used identically by the step kernels and
// by vjf_philox_normal() so that tape mode and in-kernel mode produce the same numbers.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}
__device__ __forceinline__ void philox_normal4(unsigned long long seed, unsigned long long step, unsigned long long trial,
                                               uint32_t which, uint32_t blk, float out[4]) {
  uint32_t c[4] = {(uint32_t)trial, (uint32_t)(trial >> 32), (uint32_t)step, (which << 16) | blk};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32);
  philox4x32_10(c, k0, k1);
  const float TWO_PI = 6.283185307179586f;
  // (0,1] uniforms so that log() is finite
  float u0 = ((float)(c[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);
  float u1 = (float)(c[1] >> 8) * (1.0f / 16777216.0f);
  float u2 = ((float)(c[2] >> 8) + 1.0f) * (1.0f / 16777216.0f);
  float u3 = (float)(c[3] >> 8) * (1.0f / 16777216.0f);
  // hardware transcendentals (MUFU): the draw sits on the critical path of every step and noise needs no last-bit accuracy;
  // the tape generator (vjf_philox_normal) runs this same code, so in-kernel and tape mode still agree bit for bit
  float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(TWO_PI * (u1 - 0.5f), &s0, &c0);
  __sincosf(TWO_PI * (u3 - 0.5f), &s1, &c1);
  out[0] = r0 * c0; out[1] = r0 * s0; out[2] = r1 * c1; out[3] = r1 * s1;
}

// ---- Ampere-style async global->shared copies (LDGSTS): no register staging, overlap with compute ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// stage a [rows][cols] global matrix (row stride lds) into shared memory with row stride ldd
__device__ __forceinline__ void stage_async(float* dst, int ldd, const float* src, int lds, int rows, int cols, int tid, int nthr) {
  if (((cols | lds | ldd) & 3) == 0 && ((((size_t)src) | ((size_t)dst)) & 15) == 0) {
    const int c4 = cols >> 2;
    for (int i = tid; i < rows * c4; i += nthr) { const int r = i / c4, c = (i - r * c4) << 2; cp_async16(dst + r * ldd + c, src + r * lds + c); }
  } else {
    for (int i = tid; i < rows * cols; i += nthr) { const int r = i / cols, c = i - r * cols; cp_async4(dst + r * ldd + c, src + r * lds + c); }
  }
}

__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// development aid: thread 0 of CTA 0 (phase-A stage stamps 7..23: of the first trial CTA) drops a globaltimer
// stamp into slot idx of step t (64 slots per step)
#ifndef VJF_DEBUG_STAMPS
#define VJF_STAMP(p, t, idx) do {} while (0)
#else
#define VJF_STAMP(p, t, idx)                                                              \
  do {                                                                                    \
    if ((p).dbg && threadIdx.x == 0 && blockIdx.x == (((p).overlap && (((idx) >= 7 && (idx) <= 23) || ((idx) >= 41 && (idx) <= 55))) ? (p).dbg_cta : 0)) \
      (p).dbg[(t) * 64 + (idx)] = gtime_ns();                                             \
  } while (0)
#endif

// ---- operand images of the tile pipeline that the SGD step / RLS commit keep current (see tile_kernels.cuh) ----
// The tensor core TRUNCATES fp32 operands to tf32 (profiles/micro_r02.txt), so an image holds the raw fp32 value as the
// "hi" part and x - trunc(x) as the "lo" part.
__device__ __forceinline__ float tf32_trunc_f(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// layer-1 weight element (input k, hidden unit n); bias row k = K1.  Chunk k / 32: [hi image | lo image], each [H rows][32 floats],
// 16-byte pieces XOR (row % 8)
__device__ __forceinline__ void w1k_store(const StepParams& p, int k, int n, float v) {
  const int H = p.H[0], c = k >> 5, kk = k & 31;
  float* base = p.w1k + (size_t)c * (2 * H * 32);
  const int off = n * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));
  base[off] = v;
  base[H * 32 + off] = v - tf32_trunc_f(v);
}
// element (row nq, contraction index r) of the [NQ][Rk] operand: rows < Rk hold w_chol^T, rows Rk.. hold w_mean^T
__device__ __forceinline__ void uk_store(const StepParams& p, int nq, int r, float v) {
  const int c = r >> 5, rr = r & 31;
  const int off = (c * p.tp.NQ + nq) * 32 + ((((rr >> 2) ^ (nq & 7)) << 2) | (rr & 3));
  p.uk[off] = v;
  p.uk[(p.tp.ukimg >> 2) + off] = v - tf32_trunc_f(v);
}

// clamp that propagates NaN the way torch.clamp does (fminf/fmaxf would drop it)
__device__ __forceinline__ float clip1(float g) { return g < -1.0f ? -1.0f : (g > 1.0f ? 1.0f : g); }

// internal host API shared by the translation units
struct vjf_handle {
  // sharded run
  float* xbuf; size_t xbuf_bytes; int comm_rank, comm_world; float* peer[VJF_MAX_RANKS]; unsigned comm_epoch;
  vjf_config cfg;
  vjf_layout lay64;
  StepParams base;       // dims, layout, pointers to workspace; per-call fields filled by the entry points
  float* state;
  float* partials;
  float* reduced;
  unsigned* sync_words;  // [0] barrier, [1] status, [2..] ctrl
  int device;
  int num_sms;
  int max_slots;
  size_t smem_limit;
  // staging for vjf_run_host
  void* stage_y[2];
  float* stage_u[2];
  float* stage_eps[2];
  float* stage_mu;
  float* stage_lv;
  float* stage_loss;
  size_t stage_y_sz[2], stage_u_sz[2], stage_eps_sz[2], stage_mu_sz, stage_lv_sz, stage_loss_sz;
  cudaStream_t copy_stream, compute_stream;
  cudaEvent_t ev_copied[2], ev_done[2];
  int aux_attr_set;          // aux.cu kernels' shared-memory attribute set on this handle's device
  float* fc_w; size_t fc_w_sz;  // forecast: sampled weights of every step (grow-only)
  double* P64;                   // double shadow of w_precision (vjf_set_rls_precision)
  struct BigR* bigr;             // workspace of the large-n_rbf path (bigr.cu)
  struct Wide* wide;             // workspace of the wide-observation path (wide.cu)
  double* wk_ws; size_t wk_ws_sz;  // weight-space Kalman update: fp64 R x R workspace (grow-only)
  float* w1k; float* uk;         // operand images of the tile pipeline
};

// floats of shared memory the double-precision RLS (rls_factor_f64) needs: [(2R+d)][ldm] + [R] doubles
static inline size_t vjf_rls64_floats(const StepParams& p) { return 2 * ((size_t)(2 * p.R + p.d) * p.ldm + 2 * p.R + 4) + 2 * VJF_NWARP + 16; }
void vjf_set_error(const char* fmt, ...);
extern long long g_vjf_launches;
#define VJF_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      vjf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                                  \
    }                                                                                             \
  } while (0)
