// Warp-level tile GEMMs on the tensor cores with fp32-grade accuracy (3xTF32 error compensation):
// every fp32 operand x is split into hi = tf32(x) (round to nearest) and lo = x - hi, and the product is
// accumulated in fp32 as lo*hi + hi*lo + hi*hi (the dropped lo*lo term is ~2^-22 relative; the tensor
// core reads only the upper 19 bits of lo, another 2^-22).  The tiles of this path are tiny (<= 32 trials per
// CTA), so the register-fragment mma.sync form is used: operands come straight from shared memory with
// no layout constraints.
//
// Measured on B200 (scripts/micro/mma_lat.cu): mma.sync.m16n8k8.tf32 has 21 cycles latency and issues every
// 8 cycles per SM sub-partition.  The first version split the operands in registers at every use
// (5 ALU instructions per element, repeated by every tile that touches it) and was issue-bound on those;
// operands that are reused across tiles are therefore split ONCE into a (hi, lo) pair of shared-memory
// arrays ("presplit"), and the three products go to independent accumulators to keep the mma chain short.
#pragma once
#include "common.cuh"

// Round an fp32 bit pattern to tf32 (10 explicit mantissa bits), ties away from zero.  Same result as
// cvt.rna.tf32.f32 for finite inputs, but it issues on the integer ALU instead of the quarter-rate
// conversion (XU) pipe.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ void mma_tf32(float c[4], const float a[4], const float b[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
}
// three independent accumulation chains: cl += lo_a*hi_b, cm += hi_a*lo_b, ch += hi_a*hi_b
__device__ __forceinline__ void mma3x(float cl[4], float cm[4], float ch[4], const float ah[4], const float al[4], const float bh[2],
                                      const float bl[2]) {
  mma_tf32(cl, al, bh);
  mma_tf32(cm, ah, bl);
  mma_tf32(ch, ah, bh);
}
__device__ __forceinline__ void split2(const float x[2], float h[2], float l[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) { h[i] = tf32_hi(x[i]); l[i] = x[i] - h[i]; }
}
__device__ __forceinline__ void split4(const float x[4], float h[4], float l[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { h[i] = tf32_hi(x[i]); l[i] = x[i] - h[i]; }
}

// X[i] <- tf32(X[i]), Xl[i] <- X[i] - tf32(X[i]) for i < n (all threads of the CTA)
__device__ __forceinline__ void presplit_inplace(float* X, float* Xl, int n) {
  for (int i = threadIdx.x; i < n; i += VJF_NT) {
    const float x = X[i], h = tf32_hi(x);
    X[i] = h;
    Xl[i] = x - h;
  }
}

__device__ __forceinline__ void frag_store(float* p, float v, bool first) {
  if (first) *p = v;
  else atomicAdd(p, v);
}

// Fragment coordinates of mma.m16n8k8 (PTX ISA): g = lane / 4, t = lane % 4
//   A (16x8 row):  a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
//   B (8x8 col):   b0 (k=t, n=g)  b1 (k=t+4, n=g)
//   C (16x8):      c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)

// out[b][n] = act(bias[n] + sum_k A[b][k] W[k][n]).  A: smem rows x lda, zero-padded to a multiple of 8 columns;
// Al: its presplit lo part (same layout) or nullptr when A holds plain fp32 values; W: [K][N] with row stride ldw
// (global, or staged in smem); out: smem rows x ldo (ldo >= roundup(N, 8)); rows is a multiple of 16.
__device__ __forceinline__ void mma_linear_fwd(const float* A, const float* Al, int lda, int K, const float* W, int ldw,
                                                   const float* bias, int N, float* out, int ldo, int rows, bool do_tanh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = rows >> 4, nt = (N + 7) >> 3;
  for (int tile = warp; tile < mt * nt; tile += VJF_NWARP) {
    const int m0 = (tile % mt) << 4, n0 = (tile / mt) << 3;
    const int nb = n0 + g;            // column of the B fragment this lane loads
    const bool nb_ok = nb < N;
    const int nc = n0 + 2 * t;        // first column of this lane's C fragment
    float c[4], c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
    c[0] = c[2] = (bias && nc < N) ? bias[nc] : 0.f;
    c[1] = c[3] = (bias && nc + 1 < N) ? bias[nc + 1] : 0.f;
    const int aoff = (m0 + g) * lda + t;
    const float* wp = W + nb;
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 8) {
      float ah[4], al[4], b[2], bh[2], bl[2];
      ah[0] = A[aoff + k0]; ah[1] = A[aoff + 8 * lda + k0]; ah[2] = A[aoff + k0 + 4]; ah[3] = A[aoff + 8 * lda + k0 + 4];
      if (Al) {
        al[0] = Al[aoff + k0]; al[1] = Al[aoff + 8 * lda + k0]; al[2] = Al[aoff + k0 + 4]; al[3] = Al[aoff + 8 * lda + k0 + 4];
      } else {
        float x[4] = {ah[0], ah[1], ah[2], ah[3]};
        split4(x, ah, al);
      }
      b[0] = (nb_ok && k0 + t < K) ? wp[(k0 + t) * ldw] : 0.f;
      b[1] = (nb_ok && k0 + t + 4 < K) ? wp[(k0 + t + 4) * ldw] : 0.f;
      split2(b, bh, bl);
      mma3x(c1, c2, c, ah, al, bh, bl);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] += c1[i] + c2[i];
    if (do_tanh) { c[0] = tanhf(c[0]); c[1] = tanhf(c[1]); c[2] = tanhf(c[2]); c[3] = tanhf(c[3]); }
    float* o = out + (m0 + g) * ldo + nc;   // columns in [N, roundup(N,8)) are written as zeros
    if (nc < N) { o[0] = c[0]; o[8 * ldo] = c[2]; } else { o[0] = 0.f; o[8 * ldo] = 0.f; }
    if (nc + 1 < N) { o[1] = c[1]; o[8 * ldo + 1] = c[3]; } else { o[1] = 0.f; o[8 * ldo + 1] = 0.f; }
  }
}

// dW[k][n] (+)= sum_b A[b][k] G[b][n].  A (+Al or nullptr): smem rows x lda, columns zero-padded to a multiple of 16;
// G, Gl: presplit pair, smem rows x ldg, columns zero-padded to a multiple of 8; rows in {16, 32}; dW: global [K][N].
__device__ __forceinline__ void mma_wgrad(const float* A, const float* Al, int lda, int K, const float* G, const float* Gl,
                                              int ldg, int N, int rows, float* dW, bool first) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = (K + 15) >> 4, nt = (N + 7) >> 3, ks = rows >> 3;  // ks <= 4
  for (int mtile = warp; mtile < mt; mtile += VJF_NWARP) {
    const int m0 = mtile << 4;
    float ah[4][4], al[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < ks) {
        // A^T fragment: element (m, k) = A[(8s + k) * lda + m0 + m]
        const int o = (8 * s + t) * lda + m0 + g;
        ah[s][0] = A[o]; ah[s][1] = A[o + 8]; ah[s][2] = A[o + 4 * lda]; ah[s][3] = A[o + 4 * lda + 8];
        if (Al) {
          al[s][0] = Al[o]; al[s][1] = Al[o + 8]; al[s][2] = Al[o + 4 * lda]; al[s][3] = Al[o + 4 * lda + 8];
        } else {
          float x[4] = {ah[s][0], ah[s][1], ah[s][2], ah[s][3]};
          split4(x, ah[s], al[s]);
        }
      }
    }
#pragma unroll 2
    for (int ntile = 0; ntile < nt; ++ntile) {
      const int n0 = ntile << 3;
      float c[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s < ks) {
          const int o = (8 * s + t) * ldg + n0 + g;
          const float bh[2] = {G[o], G[o + 4 * ldg]}, bl[2] = {Gl[o], Gl[o + 4 * ldg]};
          mma3x(c1, c2, c, ah[s], al[s], bh, bl);
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

// qpart[ntile][b] = sum over the 8 columns of n-tile of (phi U)[b][n]^2.  phi, phil: presplit pair, smem rows x ldp;
// U: smem Rk x ldu with Rk = roundup(R, 8) rows, zero padded (plain fp32); rows multiple of 16.
// Only k-steps up to the diagonal block are visited when U is upper triangular (upper != 0).
__device__ __forceinline__ void mma_quadform(const float* phi, const float* phil, int ldp, const float* U, int ldu, int R,
                                                 int rows, float* qpart, int upper) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = rows >> 4, nt = (R + 7) >> 3, Rk = nt << 3;
  for (int tile = warp; tile < mt * nt; tile += VJF_NWARP) {
    const int m0 = (tile % mt) << 4, n0 = (tile / mt) << 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
    const int kend = upper ? (n0 + 8) : Rk;
    const int aoff = (m0 + g) * ldp + t;
    const float* up = U + t * ldu + n0 + g;
#pragma unroll 2
    for (int k0 = 0; k0 < kend; k0 += 8) {
      const float ah[4] = {phi[aoff + k0], phi[aoff + 8 * ldp + k0], phi[aoff + k0 + 4], phi[aoff + 8 * ldp + k0 + 4]};
      const float al[4] = {phil[aoff + k0], phil[aoff + 8 * ldp + k0], phil[aoff + k0 + 4], phil[aoff + 8 * ldp + k0 + 4]};
      const float b[2] = {up[k0 * ldu], up[(k0 + 4) * ldu]};
      float bh[2], bl[2];
      split2(b, bh, bl);
      mma3x(c1, c2, c, ah, al, bh, bl);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] += c1[i] + c2[i];
    float q0 = c[0] * c[0] + c[1] * c[1], q1 = c[2] * c[2] + c[3] * c[3];
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    if (t == 0) { qpart[(tile / mt) * rows + m0 + g] = q0; qpart[(tile / mt) * rows + m0 + g + 8] = q1; }
  }
}

// Aout[k][k'] (+)= sum_b phi[b][k] phi[b][k'] for k, k' < R.  phi, phil: presplit pair, smem rows x ldp, columns zero
// padded to a multiple of 16; rows in {16, 32}; Aout: global [R][R].
__device__ __forceinline__ void mma_gram(const float* phi, const float* phil, int ldp, int R, int rows, float* Aout, bool first) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mt = (R + 15) >> 4, nt = (R + 7) >> 3, ks = rows >> 3;
  for (int tile = warp; tile < mt * nt; tile += VJF_NWARP) {
    const int m0 = (tile % mt) << 4, n0 = (tile / mt) << 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < ks) {
        const int oa = (8 * s + t) * ldp + m0 + g, ob = (8 * s + t) * ldp + n0 + g;
        const float ah[4] = {phi[oa], phi[oa + 8], phi[oa + 4 * ldp], phi[oa + 4 * ldp + 8]};
        const float al[4] = {phil[oa], phil[oa + 8], phil[oa + 4 * ldp], phil[oa + 4 * ldp + 8]};
        const float bh[2] = {phi[ob], phi[ob + 4 * ldp]}, bl[2] = {phil[ob], phil[ob + 4 * ldp]};
        mma3x(c1, c2, c, ah, al, bh, bl);
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
