// vjf_tile_kernel: the whole time loop of VJF.fit / VJF.filter (vjf/model.py:252-261, :179-221) in one cooperative launch,
// with the trial-parallel phase on the throughput tile pipeline of tile_kernels.cuh.
//
// Schedule (CTA 0 = RLS CTA, CTAs 1.. = trial CTAs; the overlapped schedule of k_persistent.cu):
//     tiles(0) ; T x { barrier 1 | B1(t): slot reduction + SGD (+ NVLink exchange) | { CTA 0: B2(t) = RLS || trial CTAs: tiles(t+1) } }
// The tiles of step t+1 need the RLS outputs of step t only from the quadratic form on (the control warp waits for them there,
// ctrl[5]), and the state-noise variance only at the dynamics NLL (ctrl[3]): everything before -- observation loads, recognition
// forward, decoder + likelihood -- runs beside the serial factorisation.
#include "tile_kernels.cuh"
#include "kernels.cuh"

// 1024-byte aligned base of the dynamic shared memory, derived by pointer arithmetic on the __shared__ array itself so that
// the compiler keeps the address space (LDS / STS instead of generic loads and stores, and no false aliasing with global
// memory); a round trip through an integer loses it -- so does a pointer argument of a non-inlined function, hence every
// non-inlined function re-derives the base here instead of taking it as a parameter.
extern __shared__ __align__(1024) unsigned char tk_smraw[];
static __device__ __forceinline__ unsigned char* tk_smem_base() { return tk_smraw + ((1024u - (smem_u32(tk_smraw) & 1023u)) & 1023u); }

static __device__ __forceinline__ unsigned tk_base_masks(const StepParams& p) {
  return 1u | ((p.flags & VJF_FLAG_WARMUP) ? 0u : 2u) | 4u;
}

template <int DX>
static __device__ __forceinline__ void tk_run_tiles(const StepParams& p, const CUtensorMap* ymap, uint32_t tmem, int t, unsigned masks, uint32_t& it0,
                                                    TkCtl& cs) {
  const TilePlan& pl = p.tp;
  unsigned char* sb = tk_smem_base();
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(sb + pl.o_f) + pl.f_bar);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ntc = (int)gridDim.x - 1, tile0 = (int)blockIdx.x - 1;
  const int ntl = (tile0 < p.ntiles) ? (p.ntiles - tile0 + ntc - 1) / ntc : 0;
  __syncthreads();  // the shared-memory workspace of the previous phase is dead
  TK_STAMP(p, t, 0, 0, 48);
  TkAcc<DX> acc;
#pragma unroll
  for (int k = 0; k < DX; ++k) acc.gdw[k] = 0.f;
  acc.gdb = 0.f; acc.ghvb = 0.f;
#pragma unroll
  for (int hc = 0; hc < 4; ++hc)
#pragma unroll
    for (int k = 0; k < DX; ++k) { acc.ghm[hc][k] = 0.f; acc.ghv[hc][k] = 0.f; }
#pragma unroll
  for (int i = 0; i < VJF_NSCAL; ++i) acc.sc[i] = 0.f;
  if (ntl > 0) {
    if (warp == TK_CTRL) {
      tk_control_step(p, ymap, sb, bars, tmem, t, ntl, it0, cs);
    } else {
      // parameters every tile of this step shares: decoder, head weights (final after the SGD step of the previous time step);
      // read past L1 (other CTAs wrote them)
      float* sf = reinterpret_cast<float*>(sb + pl.o_f);
      const float* st = p.state;
      const int D = p.D, d = p.d, H = p.H[0];
      const int n0 = d * D, n1 = n0 + D, n2 = n1 + H * d, n3 = n2 + H * d, n4 = n3 + d;
#pragma unroll 1
      for (int i = tid; i < n4; i += TK_NCT) {  // one pass of independent loads
        float v; int o;
        if (i < n0) { v = __ldcg(st + p.lay.dec_w + i); o = pl.f_dec + i; }
        else if (i < n1) { v = __ldcg(st + p.lay.dec_b + (i - n0)); o = pl.f_dec + i; }
        else if (i < n2) { v = __ldcg(st + p.lay.head_m_w + (i - n1)); o = pl.f_hm + (i - n1); }
        else if (i < n3) { v = __ldcg(st + p.lay.head_v_w + (i - n2)); o = pl.f_hv + (i - n2); }
        else { v = __ldcg(st + p.lay.head_v_b + (i - n3)); o = pl.f_hv + H * d + (i - n3); }
        sf[o] = v;
      }
      cb_sync();
      for (int j = 0; j < ntl; ++j) tk_compute_tile<DX>(p, sb, bars, tmem, t, tk_tile_of(j), it0 + j, masks, acc, j);
    }
    tk_wait(&bars[BK_DW], (it0 + ntl - 1) & 1);  // the accumulators in tensor memory are complete
  }
  __syncthreads();
  TK_STAMP(p, t, 0, 0, 49);
  tk_flush_step<DX>(p, sb, tmem, acc, ntl > 0, t);
  TK_STAMP(p, t, 0, 0, 50);
  it0 += (uint32_t)ntl;
}

static __device__ __forceinline__ void tk_run_tiles_d(const StepParams& p, const CUtensorMap* ymap, uint32_t tmem, int t, unsigned masks, uint32_t& it0,
                                                      TkCtl& cs) {
  switch (p.d) {
    case 1: case 2: tk_run_tiles<2>(p, ymap, tmem, t, masks, it0, cs); break;
    case 3: tk_run_tiles<3>(p, ymap, tmem, t, masks, it0, cs); break;
    case 4: tk_run_tiles<4>(p, ymap, tmem, t, masks, it0, cs); break;
    default: tk_run_tiles<8>(p, ymap, tmem, t, masks, it0, cs); break;
  }
}

__global__ void __launch_bounds__(VJF_NT, 1) vjf_tile_kernel(const __grid_constant__ StepParams p, const __grid_constant__ CUtensorMap ymap) {
  unsigned char* sb = tk_smem_base();
  const TilePlan& pl = p.tp;
  float* sf = reinterpret_cast<float*>(sb + pl.o_f);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sf + pl.f_bar);
  uint32_t* misc = reinterpret_cast<uint32_t*>(sf + pl.f_misc);  // [0] tensor-memory base, [1] barrier broadcast, [2] non-finite partial seen
  float* smB = reinterpret_cast<float*>(sb + pl.o_pg);           // workspace of the shared phases B1 / B2 (phi / g_pre region + weight ring)
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool trial_cta = blockIdx.x > 0;
  const bool early_rls = p.lik == VJF_LIK_POISSON;
  const unsigned n_stat_chunks = (unsigned)(((p.PS + 127) >> 7) - (p.pa >> 7));
  unsigned target = 0, target1 = 0, target2 = 0, nflag_seen = 0;

  // ---- one-time set-up: tensor memory, mbarriers, operand images of the weights ----
  if (warp == TK_CTRL) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&misc[0])) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < BK_N; ++i) mbar_init(&bars[i], (i >= BK_CX) ? TK_NCW : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    misc[1] = 0; misc[2] = 0;
  }
  {
    const int H = p.H[0];
    for (int i = blockIdx.x * VJF_NT + tid; i < (p.K1 + 1) * H; i += gridDim.x * VJF_NT) {
      const int k = i / H, n = i - k * H;
      w1k_store(p, k, n, k < p.K1 ? p.state[p.lay.mlp_w[0] + i] : p.state[p.lay.mlp_b[0] + n]);
    }
    for (int i = blockIdx.x * VJF_NT + tid; i < p.R * p.R; i += gridDim.x * VJF_NT) uk_store(p, i % p.R, i / p.R, p.state[p.lay.w_chol + i]);
    for (int i = blockIdx.x * VJF_NT + tid; i < p.R * p.d; i += gridDim.x * VJF_NT) uk_store(p, pl.Rk + i % p.d, i / p.d, p.state[p.lay.w_mean + i]);
  }
  // RBF centres and widths are not trained (vjf/module.py:20-21: requires_grad=False): staged once per launch
  for (int i = tid; i < p.R * p.du; i += VJF_NT) sf[pl.f_cen + i] = p.state[p.lay.centroid + i];
  for (int i = tid; i < p.R; i += VJF_NT) { const float w = expf(p.state[p.lay.logwidth + i]); sf[pl.f_iw + i] = -0.5f / (w * w); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  grid_barrier(p.barrier, target);

  uint32_t it0 = 0;
  TkCtl cs{0u, 0u, 0u};
  if (!trial_cta) for (int i = tid; i < p.PS; i += VJF_NT) p.partials[i] = 0.f;  // CTA 0 owns no trials: its slot stays zero

  // One iteration = the shared phases of step t (none before the first step), then the tiles of step t + 1 -- or of step t
  // again when a non-finite ELBO term has to be switched off (vjf/model.py:138-145).  The tile pipeline has a single call site
  // so that it is inlined into the kernel and reads the launch parameters from the constant bank.
  unsigned masks = tk_base_masks(p), fin = 7u;
  int attempt = 0;
  for (int t = -1;;) {
    int tt = 0;  // step whose tiles run at the end of this iteration
    if (t >= 0) {
      const unsigned epoch = (p.world > 1) ? p.epoch0 + 1u + (unsigned)t : 0u;
      // barrier 1 also tells every CTA whether any CTA saw a loss partial that is not comfortably finite
      const unsigned nflag = grid_barrier_flag(reinterpret_cast<unsigned long long*>(p.ctrl + 6), target1, misc[2] != 0, &misc[1]);
      if (tid == 0) misc[2] = 0;
      TK_STAMP(p, t, 0, 0, 51);
      const bool suspicious = nflag != nflag_seen;
      nflag_seen = nflag;
      fin = !suspicious ? 7u : term_finite_mask(p, p.partials, gridDim.x, smB);
      // a non-finite term becomes the constant 0 and carries no gradient: redo the tiles of this step without it
      const unsigned nm = masks & (fin | ~7u);
      if (p.world == 1 && attempt == 0 && nm != masks && (p.flags & VJF_FLAG_SGD)) {
        masks = nm;
        attempt = 1;
        grid_barrier(p.barrier, target);
        tt = t;
      } else {
        if (early_rls) {
          // Poisson likelihood: nothing in the RLS depends on the SGD step -- CTA 0 starts the factorisation as soon as the
          // statistics chunks are reduced, the trial CTAs finish the gradient reduction + SGD and go on to the tiles of step t+1
          if (trial_cta) {
            phase_b1(p, smB, p.partials, gridDim.x, true, blockIdx.x - 1, gridDim.x - 1, p.ctrl + 1, epoch, fin, tk_base_masks(p));
            __syncthreads();
            TK_STAMP(p, t, 0, 0, 52);
            target2 += gridDim.x - 1;
            if (tid == 0) {
              __threadfence();
              red_release_add_u32(p.ctrl + 2, 1u);
              while (ld_acquire_u32(p.ctrl + 2) < target2) __nanosleep(32);
              __threadfence();
            }
            __syncthreads();
            TK_STAMP(p, t, 0, 0, 53);
          } else {
            phase_b2(p, smB, t, fin, p.ctrl + 1, n_stat_chunks * (unsigned)(t + 1));
            __syncthreads();
            if (tid == 0) { __threadfence(); st_release_gpu_u32(p.ctrl + 3, (unsigned)(t + 1)); }
          }
        } else {
          phase_b1(p, smB, p.partials, gridDim.x, true, blockIdx.x, gridDim.x, nullptr, epoch, fin, tk_base_masks(p));
          grid_barrier(p.barrier, target);
          if (!trial_cta) {
            phase_b2(p, smB, t, fin);
            __syncthreads();
            if (tid == 0) { __threadfence(); st_release_gpu_u32(p.ctrl + 3, (unsigned)(t + 1)); }
          }
        }
        attempt = 0;
        masks = tk_base_masks(p);
        tt = t + 1;
      }
    }
    if (trial_cta && tt < p.T) tk_run_tiles_d(p, &ymap, tmem, tt, masks, it0, cs);
    t = tt;
    if (t >= p.T) break;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TK_CTRL) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
