// Warp-level tile GEMMs on the tensor cores with fp32-grade accuracy (3xTF32 error compensation):
// every fp32 operand is split into hi = tf32(x) and lo = tf32(x - hi) and the product is accumulated as
// lo*hi + hi*lo + hi*hi in fp32 (the dropped lo*lo term is ~2^-22 relative).  The tiles of this path are
// tiny (<= 32 trials per CTA), so the register-fragment mma.sync form is used: operands come straight
// from shared memory / L1 with no layout constraints.
#pragma once
#include "common.cuh"

// Round an fp32 bit pattern to tf32 (10 explicit mantissa bits), ties away from zero: add half an ulp of
// the kept field and clear the 13 dropped bits.  Same result as cvt.rna.tf32.f32 for finite inputs, but it
// issues on the integer ALU instead of the quarter-rate conversion (XU) pipe, which the profile showed to
// be the limiter of the 3xTF32 operand split.
__device__ __forceinline__ uint32_t f2tf32(float x) {
  return (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float c[4], const uint32_t a[4], const uint32_t b[2]) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// c += a * b with a, b given as raw fp32 fragments
__device__ __forceinline__ void mma3(float c[4], const float a[4], const float b[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}
__device__ __forceinline__ void mma3_presplit(float c[4], const uint32_t ah[4], const uint32_t al[4], const float b[2]) {
  uint32_t bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}

// Same product accumulated into three independent chains (lo*hi, hi*lo, hi*hi): the caller sums them once
// at the end.  Shortens the dependent mma chain by 3x (mma.sync latency ~21 cycles on sm_100a).
__device__ __forceinline__ void mma3x(float cl[4], float cm[4], float ch[4], const float a[4], const float b[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
  mma_tf32(cl, al, bh);
  mma_tf32(cm, ah, bl);
  mma_tf32(ch, ah, bh);
}
__device__ __forceinline__ void mma3x_presplit(float cl[4], float cm[4], float ch[4], const uint32_t ah[4], const uint32_t al[4],
                                               const float b[2]) {
  uint32_t bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
  mma_tf32(cl, al, bh);
  mma_tf32(cm, ah, bl);
  mma_tf32(ch, ah, bh);
}

__device__ __forceinline__ void frag_store(float* p, float v, bool first) {
  if (first) *p = v;
  else atomicAdd(p, v);
}

// Fragment coordinates of mma.m16n8k8 (PTX ISA): g = lane / 4, t = lane % 4
//   A (16x8 row):  a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
//   B (8x8 col):   b0 (k=t, n=g)  b1 (k=t+4, n=g)
//   C (16x8):      c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)

// out[b][n] = act(bias[n] + sum_k A[b][k] W[k][n]);  A: smem rows x lda (zero-padded to a multiple of 8
// columns), W: [K][N] with row stride ldw (global, or staged in smem); out: smem rows x ldo (ldo >= roundup(N, 8)); rows is a multiple of 16.
__device__ __forceinline__ void mma_linear_fwd(const float* A, int lda, int K, const float* W, int ldw, const float* bias, int N,
                                               float* out, int ldo, int rows, bool do_tanh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = rows >> 4, nt = (N + 7) >> 3;
  for (int tile = warp; tile < mt * nt; tile += VJF_NWARP) {
    const int m0 = (tile % mt) << 4, n0 = (tile / mt) << 3;
    const int nb = n0 + g;            // column of the B fragment this lane loads
    const bool nb_ok = nb < N;
    const int nc = n0 + 2 * t;        // first column of this lane's C fragment
    float c[4], c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f}, c3[4] = {0.f, 0.f, 0.f, 0.f},
                c4[4] = {0.f, 0.f, 0.f, 0.f}, c5[4] = {0.f, 0.f, 0.f, 0.f};
    c[0] = c[2] = (bias && nc < N) ? bias[nc] : 0.f;
    c[1] = c[3] = (bias && nc + 1 < N) ? bias[nc + 1] : 0.f;
    const float* a_lo = A + (m0 + g) * lda + t;
    const float* a_hi = a_lo + 8 * lda;
    const float* wp = W + nb;
    int k0 = 0;
    for (; k0 + 8 < K; k0 += 16) {  // two k-steps per trip, six independent accumulation chains
      float a[4], b[2], a2[4], b2[2];
      a[0] = a_lo[k0]; a[1] = a_hi[k0]; a[2] = a_lo[k0 + 4]; a[3] = a_hi[k0 + 4];
      a2[0] = a_lo[k0 + 8]; a2[1] = a_hi[k0 + 8]; a2[2] = a_lo[k0 + 12]; a2[3] = a_hi[k0 + 12];
      b[0] = nb_ok ? wp[(k0 + t) * ldw] : 0.f;
      b[1] = nb_ok ? wp[(k0 + t + 4) * ldw] : 0.f;
      b2[0] = (nb_ok && k0 + 8 + t < K) ? wp[(k0 + 8 + t) * ldw] : 0.f;
      b2[1] = (nb_ok && k0 + 12 + t < K) ? wp[(k0 + 12 + t) * ldw] : 0.f;
      mma3x(c, c1, c2, a, b);
      mma3x(c3, c4, c5, a2, b2);
    }
    for (; k0 < K; k0 += 8) {
      float a[4], b[2];
      a[0] = a_lo[k0]; a[1] = a_hi[k0]; a[2] = a_lo[k0 + 4]; a[3] = a_hi[k0 + 4];
      b[0] = (nb_ok && k0 + t < K) ? wp[(k0 + t) * ldw] : 0.f;
      b[1] = (nb_ok && k0 + t + 4 < K) ? wp[(k0 + t + 4) * ldw] : 0.f;
      mma3x(c, c1, c2, a, b);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] += ((c1[i] + c2[i]) + (c3[i] + c4[i])) + c5[i];
    if (do_tanh) { c[0] = tanhf(c[0]); c[1] = tanhf(c[1]); c[2] = tanhf(c[2]); c[3] = tanhf(c[3]); }
    float* o = out + (m0 + g) * ldo + nc;   // columns >= N of `out` receive act(0 [+0]) : finite padding
    if (nc < N) { o[0] = c[0]; o[8 * ldo] = c[2]; } else { o[0] = 0.f; o[8 * ldo] = 0.f; }
    if (nc + 1 < N) { o[1] = c[1]; o[8 * ldo + 1] = c[3]; } else { o[1] = 0.f; o[8 * ldo + 1] = 0.f; }
  }
}

// dW[k][n] (+)= sum_b A[b][k] G[b][n];  A: smem rows x lda (columns zero-padded to a multiple of 16),
// G: smem rows x ldg (columns zero-padded to a multiple of 8); rows in {16, 32}; dW: global [K][N].
__device__ __forceinline__ void mma_wgrad(const float* A, int lda, int K, const float* G, int ldg, int N, int rows,
                                          float* dW, bool first) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = (K + 15) >> 4, nt = (N + 7) >> 3, ks = rows >> 3;  // ks <= 4
  for (int mtile = warp; mtile < mt; mtile += VJF_NWARP) {
    const int m0 = mtile << 4;
    uint32_t ah[4][4], al[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < ks) {
        // A^T fragment: element (m, k) = A[(8s + k) * lda + m0 + m]
        const float* ap = A + (8 * s + t) * lda + m0 + g;
        split_tf32(ap[0], ah[s][0], al[s][0]);
        split_tf32(ap[8], ah[s][1], al[s][1]);
        split_tf32(ap[4 * lda], ah[s][2], al[s][2]);
        split_tf32(ap[4 * lda + 8], ah[s][3], al[s][3]);
      }
    }
    for (int ntile = 0; ntile < nt; ++ntile) {
      const int n0 = ntile << 3;
      float c[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s < ks) {
          float b[2];
          b[0] = G[(8 * s + t) * ldg + n0 + g];
          b[1] = G[(8 * s + t + 4) * ldg + n0 + g];
          mma3x_presplit(c, c1, c2, ah[s], al[s], b);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) c[i] += c1[i] + c2[i];
      const int r0 = m0 + g, nc = n0 + 2 * t;
      float* o = dW + r0 * N + nc;
      if (r0 < K) {
        if (nc < N) frag_store(o, c[0], first);
        if (nc + 1 < N) frag_store(o + 1, c[1], first);
      }
      if (r0 + 8 < K) {
        if (nc < N) frag_store(o + 8 * N, c[2], first);
        if (nc + 1 < N) frag_store(o + 8 * N + 1, c[3], first);
      }
    }
  }
}

// qpart[ntile][b] = sum over the 8 columns of n-tile of (phi U)[b][n]^2 ;  phi: smem rows x ldp,
// U: smem Rk x ldu with Rk = roundup(R, 8) rows, zero padded; rows multiple of 16.
// Only k-steps up to the diagonal block are visited when U is upper triangular (upper != 0).
__device__ __forceinline__ void mma_quadform(const float* phi, int ldp, const float* U, int ldu, int R, int rows,
                                             float* qpart, int upper) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = rows >> 4, nt = (R + 7) >> 3, Rk = nt << 3;
  for (int tile = warp; tile < mt * nt; tile += VJF_NWARP) {
    const int m0 = (tile % mt) << 4, n0 = (tile / mt) << 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
    const int kend = upper ? (n0 + 8) : Rk;
    const float* a_lo = phi + (m0 + g) * ldp + t;
    const float* a_hi = a_lo + 8 * ldp;
    const float* up = U + t * ldu + n0 + g;
    for (int k0 = 0; k0 < kend; k0 += 8) {
      float a[4], b[2];
      a[0] = a_lo[k0]; a[1] = a_hi[k0]; a[2] = a_lo[k0 + 4]; a[3] = a_hi[k0 + 4];
      b[0] = up[k0 * ldu]; b[1] = up[(k0 + 4) * ldu];
      mma3x(c, c1, c2, a, b);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] += c1[i] + c2[i];
    float q0 = c[0] * c[0] + c[1] * c[1], q1 = c[2] * c[2] + c[3] * c[3];
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    if (t == 0) { qpart[(tile / mt) * rows + m0 + g] = q0; qpart[(tile / mt) * rows + m0 + g + 8] = q1; }
  }
}

// Aout[k][k'] (+)= sum_b phi[b][k] phi[b][k'] for k, k' < R;  phi: smem rows x ldp, columns zero padded to a
// multiple of 16; rows in {16, 32}; Aout: global [R][R].
__device__ __forceinline__ void mma_gram(const float* phi, int ldp, int R, int rows, float* Aout, bool first) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = (R + 15) >> 4, nt = (R + 7) >> 3, ks = rows >> 3;
  for (int tile = warp; tile < mt * nt; tile += VJF_NWARP) {
    const int m0 = (tile % mt) << 4, n0 = (tile / mt) << 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < ks) {
        const float* ap = phi + (8 * s + t) * ldp + m0 + g;
        float a[4], b[2];
        a[0] = ap[0]; a[1] = ap[8]; a[2] = ap[4 * ldp]; a[3] = ap[4 * ldp + 8];
        b[0] = phi[(8 * s + t) * ldp + n0 + g];
        b[1] = phi[(8 * s + t + 4) * ldp + n0 + g];
        mma3x(c, c1, c2, a, b);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] += c1[i] + c2[i];
    const int r0 = m0 + g, nc = n0 + 2 * t;
    float* o = Aout + r0 * R + nc;
    if (r0 < R) {
      if (nc < R) frag_store(o, c[0], first);
      if (nc + 1 < R) frag_store(o + 1, c[1], first);
    }
    if (r0 + 8 < R) {
      if (nc < R) frag_store(o + 8 * R, c[2], first);
      if (nc + 1 < R) frag_store(o + 8 * R + 1, c[3], first);
    }
  }
}
