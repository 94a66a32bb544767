// The persistent cooperative kernel: the whole time loop of VJF.fit / VJF.filter in one launch (vjf/model.py:252-261, :179-221).
#include "step_kernels.cuh"
#include "kernels.cuh"

__device__ __forceinline__ unsigned base_masks(const StepParams& p) {
  return 1u | ((p.flags & VJF_FLAG_WARMUP) ? 0u : 2u) | 4u;
}

// The whole time loop in one cooperative launch.
//
// Plain schedule (several tiles per CTA):   T x { A | bar | B1 | bar | B2 (CTA 0) | bar }
// Overlapped schedule (one tile per trial CTA, CTA 0 dedicated to the RLS):
//     front(0) ; T x { back(t) | bar | B1(t) | bar | { CTA 0: B2(t)  ||  trial CTAs: front(t+1) } | bar }
// front(t+1) needs only the SGD-updated parameters (B1), not the RLS outputs, so it hides behind the
// serial factorisation of step t.
// RLS CTA, overlapped schedule: everything phase B2 of step t writes (state-noise logvar last) is final
static __device__ __forceinline__ void publish_step_done(const StepParams& p, int t) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); st_release_gpu_u32(p.ctrl + 3, (unsigned)(t + 1)); }
}

__global__ void __launch_bounds__(VJF_NT, 1) vjf_persistent_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) float sm[];
  unsigned target = 0, target2 = 0, target1 = 0, nflag_seen = 0;
  bool back_staged = false;  // w_chol / w_mean of the coming back half already staged (issued behind the RLS tail)
  const bool trial_cta = blockIdx.x > 0;
  const bool early_rls = p.overlap && p.lik == VJF_LIK_POISSON;
  const unsigned n_stat_chunks = (unsigned)(((p.PS + 127) >> 7) - (p.pa >> 7));
  // Blackwell asynchronous machinery of the overlapped schedule (step_kernels.cuh, TileCtx): every trial CTA owns
  // 256 TMEM columns (tcgen05 weight gradient) and three mbarriers (tcgen05 commit, TMA weights, TMA observations)
  TileCtx ctx{0u, 0u, 0u, 0u, -1, 0, 0, -1, 0};
  TileCtx* cx = nullptr;
  if (p.overlap && trial_cta && (p.use_umma || p.use_tma)) {
    cx = &ctx;
    ctx.flags = (p.use_umma ? CX_UMMA : 0) | (p.use_tma ? CX_TMA : 0);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + p.s_flag + 1);
    if (p.use_umma && threadIdx.x < 32) tmem_alloc256(tslot);
    if (threadIdx.x == 0) {
      for (int i = 0; i < 3; ++i) mbar_init(reinterpret_cast<uint64_t*>(sm + p.s_flag + 2 + 2 * i), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.use_umma) ctx.tmem = *tslot;
  }
  if (threadIdx.x == 0) *reinterpret_cast<int*>(sm + p.s_flag + 8) = 0;
  if (p.use_tma) {  // (re)build the row-padded mirror of the layer-1 weight; first read after the first grid barrier
    const int n = p.K1 * p.H[0];
    for (int i = blockIdx.x * VJF_NT + threadIdx.x; i < n; i += gridDim.x * VJF_NT) {
      const int r = i / p.H[0];
      p.w1_mirror[r * p.ldw1 + (i - r * p.H[0])] = p.state[p.lay.mlp_w[0] + i];
    }
    for (int i = blockIdx.x * VJF_NT + threadIdx.x; i < p.R * p.R; i += gridDim.x * VJF_NT) {
      const int r = i / p.R;
      p.u_mirror[r * p.ldu + (i - r * p.R)] = p.state[p.lay.w_chol + i];
    }
    grid_barrier(p.barrier, target);  // the front half of step 0 reads the mirror (TMA) right away
  }
  if (p.overlap) {
    if (!trial_cta) {
      float* slot = p.partials;  // CTA 0 owns no trials: its slot stays zero
      for (int i = threadIdx.x; i < p.PS; i += VJF_NT) slot[i] = 0.f;
    } else {
      phase_a_prologue(p, sm, STAGE_FRONT, cx);
      phase_a_tile(p, sm, 0, blockIdx.x - 1, true, base_masks(p), PART_FRONT, cx);
    }
  }
  for (int t = 0; t < p.T; ++t) {
    unsigned masks = base_masks(p), fin;
    const unsigned epoch = (p.world > 1) ? p.epoch0 + 1u + (unsigned)t : 0u;
    VJF_STAMP(p, t, 0);
#ifdef VJF_DEBUG_STAMPS
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[t * 64 + 40] = clock64();
#endif
    for (int attempt = 0;; ++attempt) {
      if (!p.overlap) {
        phase_a(p, sm, t, masks);
      } else if (trial_cta) {
        if (!back_staged) phase_a_prologue(p, sm, STAGE_BACK, cx);
        back_staged = false;
        phase_a_tile(p, sm, t, blockIdx.x - 1, true, masks, PART_BACK, cx);
      }
      VJF_STAMP(p, t, 1);
      // barrier 1 also tells every CTA whether any CTA saw a loss partial that is not comfortably finite this attempt
      int* nf = reinterpret_cast<int*>(sm + p.s_flag + 8);
      const unsigned nflag = grid_barrier_flag(reinterpret_cast<unsigned long long*>(p.ctrl + 6), target1, *nf != 0,
                                               reinterpret_cast<unsigned*>(sm + p.s_flag + 9));
      if (threadIdx.x == 0) *nf = 0;
      const bool suspicious = nflag != nflag_seen;
      nflag_seen = nflag;
      VJF_STAMP(p, t, 2);
      fin = !suspicious ? 7u : term_finite_mask(p, p.partials, gridDim.x, sm);  // sharded: the LOCAL mask; the ranks exchange it with the chunk flags (phase_b1)
      // vjf/model.py:138-145: a non-finite term becomes the constant 0 => it must not contribute a
      // gradient either.  Rare; redo the trial-parallel phase with that term switched off.
      const unsigned nm = masks & (fin | ~7u);
      if (p.world == 1 && attempt == 0 && nm != masks && (p.flags & VJF_FLAG_SGD)) {
        masks = nm;
        grid_barrier(p.barrier, target);
        if (p.overlap && trial_cta) {
          phase_a_prologue(p, sm, STAGE_FRONT, cx);
          phase_a_tile(p, sm, t, blockIdx.x - 1, true, masks, PART_FRONT, cx);
        }
        continue;
      }
      break;
    }
    if (early_rls) {
      // Poisson likelihood + overlapped schedule: nothing in the RLS depends on the SGD step, so CTA 0 starts the
      // factorisation as soon as the statistics chunks are reduced, while the trial CTAs finish the gradient
      // reduction + SGD, synchronise among themselves and go on to the front half of step t+1.
      if (trial_cta) {
        phase_b1(p, sm, p.partials, gridDim.x, true, blockIdx.x - 1, gridDim.x - 1, p.ctrl + 1, epoch, fin, base_masks(p));
        VJF_STAMP(p, t, 3);
        VJF_STAMP(p, t, 23);
        // trial-only barrier; the Philox draw of step t+1 for this tile is computed by otherwise idle warps while
        // thread 0 polls (noise depends on nothing but (seed, step, trial))
        {
          __syncthreads();
          target2 += gridDim.x - 1;
          if (threadIdx.x == 0) { __threadfence(); red_release_add_u32(p.ctrl + 2, 1u); }
          if (cx && (cx->flags & CX_TMA) && !p.eps && t + 1 < p.T) {
            const int b0 = (blockIdx.x - 1) * p.TB, nb = min(p.TB, p.B - b0), nblk = (p.d + 3) >> 2;
            float* eps_s = sm + p.s_eps;
            const int i = (int)threadIdx.x - 32;
            if (i >= 0 && i < nb * 2 * nblk) {
              const int b = i / (2 * nblk), r = i - b * 2 * nblk, which = r / nblk, blk = r - which * nblk;
              float z[4];
              philox_normal4(p.seed, p.step0 + t + 1, p.trial_offset + b0 + b, which, blk, z);
              for (int k = 0; k < 4; ++k)
                if (blk * 4 + k < p.d) eps_s[b * 2 * p.d + which * p.d + blk * 4 + k] = z[k];
            }
            cx->eps_ready_t = t + 1;
          }
          if (threadIdx.x == 0) {
            while (ld_acquire_u32(p.ctrl + 2) < target2) __nanosleep(32);
            __threadfence();
          }
          __syncthreads();
        }
        VJF_STAMP(p, t, 21);
        if (t + 1 < p.T) {
          phase_a_prologue(p, sm, STAGE_FRONT, cx);
          VJF_STAMP(p, t, 7);
          if (cx) cx->early_ok = 1;
          phase_a_tile(p, sm, t + 1, blockIdx.x - 1, true, base_masks(p), PART_FRONT, cx);
          if (cx) cx->early_ok = 0;
          // w_chol / w_mean of step t are published before the RLS CTA finishes the step: stage them now, behind its tail
          // (unless the front half found them published already and issued the copies itself)
          if (!(cx && (cx->flags & CX_TMA) && *reinterpret_cast<volatile int*>(sm + p.s_flag + 10))) {
            wait_counter(p.ctrl + 5, (unsigned)(t + 1));
            phase_a_prologue(p, sm, STAGE_BACK, cx);
          }
          back_staged = true;
        }
      } else {
        VJF_STAMP(p, t, 3);
        VJF_STAMP(p, t, 4);
        phase_b2(p, sm, t, fin, p.ctrl + 1, n_stat_chunks * (unsigned)(t + 1));
        publish_step_done(p, t);
      }
    } else {
      phase_b1(p, sm, p.partials, gridDim.x, true, blockIdx.x, gridDim.x, nullptr, epoch, fin, base_masks(p));
      VJF_STAMP(p, t, 3);
      grid_barrier(p.barrier, target);
      VJF_STAMP(p, t, 4);
      if (blockIdx.x == 0) {
        phase_b2(p, sm, t, fin);
        if (p.overlap) publish_step_done(p, t);
      } else if (p.overlap && t + 1 < p.T) {
        phase_a_prologue(p, sm, STAGE_FRONT, cx);
        if (cx) cx->early_ok = 1;
        phase_a_tile(p, sm, t + 1, blockIdx.x - 1, true, base_masks(p), PART_FRONT, cx);
        if (cx) cx->early_ok = 0;
        if (!(cx && (cx->flags & CX_TMA) && *reinterpret_cast<volatile int*>(sm + p.s_flag + 10))) {
          wait_counter(p.ctrl + 5, (unsigned)(t + 1));
          phase_a_prologue(p, sm, STAGE_BACK, cx);
        }
        back_staged = true;
      }
    }
    VJF_STAMP(p, t, 5);
    // Plain schedule: barrier 3 ends the step.  Overlapped schedule: nothing a trial CTA reads next is ordered by it any more
    // (w_chol / w_mean: ctrl[5]; state-noise logvar: ctrl[3]; Gaussian likelihood logvar: ctrl[4]; the RLS CTA takes part in
    // barrier 1 of the next step only after its tail, which orders the reuse of the reduced vector), so it is dropped and the
    // back half of step t+1 starts while the RLS CTA is still in the residual / noise-variance tail of step t.
    if (!p.overlap) grid_barrier(p.barrier, target);
    VJF_STAMP(p, t, 6);
  }
  if (cx && p.use_umma) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_free256(ctx.tmem);
  }
}


#ifdef VJF_DEBUG_STAMPS
// development aid: clock64 at every column of the last RLS sweep of the persistent kernel
extern "C" int vjf_debug_read_sweep(long long* host_out) { return (int)cudaMemcpyFromSymbol(host_out, g_sweep_ticks, sizeof(long long) * 160); }
#endif
