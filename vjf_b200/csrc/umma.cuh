// Hand-written Blackwell tensor-core path (sm_100a): tcgen05.mma kind::tf32 issued by ONE thread, operands read
// by the tensor core straight from shared memory through matrix descriptors, fp32 accumulators in TMEM, results
// read back with tcgen05.ld.  Validated in isolation by scripts/micro/umma_test.cu (K-major x K-major,
// M = 64, N in {32, 208}, no-swizzle canonical layout; the MN-major no-swizzle descriptors returned zeros on
// this driver, so every operand is laid out K-major).
//
// Canonical K-major no-swizzle layout of an [MN x K] tf32 operand ("core matrix" = 8 rows x 16 bytes):
//     float index of element (r, k) = ((k / 4) * MN + r) * 4 + (k % 4)
// i.e. 16-byte chunks of 4 consecutive k, all MN rows of a chunk contiguous.  Descriptor: LBO = byte distance
// between the two 16-byte K-halves of one MMA (K = 8) = MN * 16, SBO = byte distance between 8-row groups = 128.
#pragma once
#include "common.cuh"

__device__ __forceinline__ int umma_canon(int r, int k, int MN) { return (((k >> 2) * MN + r) << 2) + (k & 3); }

// cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version = 1 [46,48) | layout_type = SWIZZLE_NONE [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 = 1 [4,6) | a_format TF32 = 2 [7,10) | b_format TF32 = 2 [10,13) | a/b K-major | N>>3 [17,23) | M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// ---- TMA bulk copies (global -> shared, completion counted in bytes on an mbarrier) ----
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst, src and bytes must be multiples of 16
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation (one warp; 256 columns): the base address is written to *slot in shared memory
__device__ __forceinline__ void tmem_alloc256(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free256(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base) : "memory");
}
// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread i of warp w reads TMEM lane 32 * (w % 4) + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Recognition layer-1 weight gradient on the 5th-generation tensor cores:
//     dW[k1][n] (+)= sum_b in[b][k1] * G[b][n]      computed as D[M = 64 (n)][N = NK (k1)] = G^T [64 x rows] * in_b [NK x rows]^T
// Operands (3xTF32: hi and lo arrays each) in the canonical K-major layout with K = trials:
//     gt_*  : [64 x rows]  (MN = 64),  element (n, b)      in_b_* : [NK x rows] (MN = NK), element (k1, b),   NK = roundup(K1, 8) <= 224
// (`bar` is an mbarrier initialised once per kernel with count 1; `phase` = parity of its current phase.)
// One thread issues rows/8 k-steps x 3 MMAs (lo*hi, hi*lo, hi*hi) accumulating in TMEM columns [tcol, tcol + NK); 8 warps
// then read the accumulator back (row n lives in TMEM lane n % 16 + 32 * (n / 16)) and store/add it to the slot.
__device__ __forceinline__ void umma_wgrad(const float* gt_hi, const float* gt_lo, const float* inb_hi, const float* inb_lo, int NK, int K1, int H,
                                           int rows, uint32_t tmem_base, uint32_t tcol, uint64_t* bar, uint32_t phase, float* dW, float* dB, float* stage,
                                           bool first, const StepParams* sp = nullptr, int t = 0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // operands were written with ordinary shared-memory stores: make them visible to the async proxy, then hand over
  if (sp) VJF_STAMP(*sp, t, 52);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (sp) VJF_STAMP(*sp, t, 53);
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_tf32(64, NK);
    const uint32_t a_hi = smem_u32(gt_hi), a_lo = smem_u32(gt_lo), b_hi = smem_u32(inb_hi), b_lo = smem_u32(inb_lo);
    const uint32_t d = tmem_base + tcol;
    for (int s = 0; s < (rows >> 3); ++s) {
      const uint32_t ao = s * 2 * 64 * 16, bo = s * 2 * NK * 16;  // two 16-byte K-chunks per MMA
      const uint64_t adh = umma_desc(a_hi + ao, 64 * 16, 128), adl = umma_desc(a_lo + ao, 64 * 16, 128);
      const uint64_t bdh = umma_desc(b_hi + bo, NK * 16, 128), bdl = umma_desc(b_lo + bo, NK * 16, 128);
      umma_tf32_ss(d, adl, bdh, idesc, s > 0);
      umma_tf32_ss(d, adh, bdl, idesc, 1);
      umma_tf32_ss(d, adh, bdh, idesc, 1);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, phase);
  tc_fence_after();
  if (sp) VJF_STAMP(*sp, t, 54);
  // epilogue: a warp reaches the TMEM lanes of sub-partition q = warp % 4, which hold rows n = 16q .. 16q+15 in lanes
  // 0..15 (M = 64 layout).  The four warp groups take 32-column chunks; a chunk goes through shared memory (`stage`, the
  // operand area the MMAs have finished reading) so that the slot is written with 128-bit stores of whole [k1][0..63] rows.
  // Column NK-1 carries the bias gradient when `dB` is given (a row of ones in the B operand).
  {
    const int q = warp & 3, grp = warp >> 2, n = 16 * q + lane, tg = tid & 127;
    float* stg = stage + grp * (32 * 64);
    for (int it = 0; it < 2; ++it) {
      const int c0 = grp * 32 + it * 128;
      if (c0 < NK) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + c0, v);
        if (lane < 16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[j * 64 + n] = v[j];
        }
      }
      __syncthreads();
      if (c0 < NK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = tg + 128 * i, j = f >> 4, n4 = (f & 15) << 2, k1 = c0 + j;
          if (n4 < H) {
            const float4 val = *reinterpret_cast<const float4*>(stg + j * 64 + n4);
            float* o = (k1 < K1) ? dW + (size_t)k1 * H + n4 : ((dB && k1 == NK - 1) ? dB + n4 : nullptr);
            if (o) {
              if (first) *reinterpret_cast<float4*>(o) = val;
              else { atomicAdd(o, val.x); atomicAdd(o + 1, val.y); atomicAdd(o + 2, val.z); atomicAdd(o + 3, val.w); }
            }
          }
        }
      }
      if (it == 0) __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
}
