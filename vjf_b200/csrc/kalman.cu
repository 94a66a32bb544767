// Batched small-matrix Kalman operator: vjf/kalman.py:15-145 and vjf/numerical.py:8-19.
// One warp per problem, every matrix of the problem in shared memory (n <= 16, m <= 32, nb <= 32).
// The reference's layout is kept: states are (n, batch) column-batched, covariances travel as their
// lower Cholesky factors (cholesky=True in the reference calls).
#include "common.cuh"

#define KN 16
#define KM 32
#define KB 32

namespace {

// C[M x N] = A[M x K] * B   (B is [K x N], or [N x K] when transB)
__device__ __forceinline__ void w_matmul(float* Cm, const float* A, const float* B, int M, int N, int K, bool transB,
                                         int lane) {
  for (int i = lane; i < M * N; i += 32) {
    const int r = i / N, c = i - r * N;
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(A[r * K + k], transB ? B[c * K + k] : B[k * N + c], s);
    Cm[i] = s;
  }
  __syncwarp();
}

// in-place lower Cholesky of the symmetric n x n matrix a; returns false if a pivot is not positive
__device__ __forceinline__ bool w_cholesky(float* a, int n, int lane) {
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    const float piv = a[k * n + k];
    if (!(piv > 0.f)) { ok = false; break; }
    const float s = sqrtf(piv);
    __syncwarp();
    for (int i = k + lane; i < n; i += 32) a[i * n + k] = (i == k) ? s : a[i * n + k] / s;
    __syncwarp();
    for (int i = lane; i < (n - k - 1) * (n - k - 1); i += 32) {
      const int r = k + 1 + i / (n - k - 1), c = k + 1 + i % (n - k - 1);
      if (c <= r) a[r * n + c] = fmaf(-a[r * n + k], a[c * n + k], a[r * n + c]);
    }
    __syncwarp();
  }
  for (int i = lane; i < n * n; i += 32) { const int r = i / n, c = i - r * n; if (c > r) a[i] = 0.f; }
  __syncwarp();
  return ok;
}

// X <- L^-1 X  (L lower n x n, X is n x c): each lane owns columns
__device__ __forceinline__ void w_trsm_lower(const float* L, float* X, int n, int c, int lane) {
  for (int j = lane; j < c; j += 32)
    for (int k = 0; k < n; ++k) {
      float s = X[k * c + j];
      for (int i = 0; i < k; ++i) s = fmaf(-L[k * n + i], X[i * c + j], s);
      X[k * c + j] = s / L[k * n + k];
    }
  __syncwarp();
}
// X <- L^-T X
__device__ __forceinline__ void w_trsm_lower_t(const float* L, float* X, int n, int c, int lane) {
  for (int j = lane; j < c; j += 32)
    for (int k = n - 1; k >= 0; --k) {
      float s = X[k * c + j];
      for (int i = k + 1; i < n; ++i) s = fmaf(-L[i * n + k], X[i * c + j], s);
      X[k * c + j] = s / L[k * n + k];
    }
  __syncwarp();
}

__device__ __forceinline__ void w_load(float* dst, const float* src, int n, int lane) {
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
  __syncwarp();
}
__device__ __forceinline__ void w_store(float* dst, const float* src, int n, int lane) {
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
  __syncwarp();
}

// kalman.predict (vjf/kalman.py:15-50), cholesky=True
__global__ void kalman_predict_kernel(int P, int n, int m, int nb, const float* x, const float* L, const float* A,
                                      const float* Q, const float* H, float* yhat, float* xhat, float* Lhat, int* info) {
  __shared__ float sA[KN * KN], sL[KN * KN], sT[KN * KN], sV[KN * KN], sx[KN * KB], sxh[KN * KB], sH[KM * KN], sy[KM * KB];
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= P) return;
  w_load(sA, A + (size_t)p * n * n, n * n, lane);
  w_load(sL, L + (size_t)p * n * n, n * n, lane);
  w_load(sx, x + (size_t)p * n * nb, n * nb, lane);
  w_load(sH, H + (size_t)p * m * n, m * n, lane);
  w_matmul(sxh, sA, sx, n, nb, n, false, lane);        // xhat = A x            (:40)
  w_matmul(sT, sA, sL, n, n, n, false, lane);          // AL                    (:45)
  w_matmul(sV, sT, sT, n, n, n, true, lane);           // AL AL^T               (:46)
  for (int i = lane; i < n * n; i += 32) sV[i] += Q[(size_t)p * n * n + i];
  __syncwarp();
  w_matmul(sy, sH, sxh, m, nb, n, false, lane);        // yhat = H xhat         (:47)
  const bool ok = w_cholesky(sV, n, lane);             //                       (:48-49)
  w_store(yhat + (size_t)p * m * nb, sy, m * nb, lane);
  w_store(xhat + (size_t)p * n * nb, sxh, n * nb, lane);
  w_store(Lhat + (size_t)p * n * n, sV, n * n, lane);
  if (lane == 0 && info) info[p] = ok ? 0 : 1;
}

// shared front end of update / joseph_update: e, Vhat, S = H Lhat (H Lhat)^T + R, chol(S)
struct UpdShared {
  float Lh[KN * KN], Vh[KN * KN], H[KM * KN], HL[KM * KN], S[KM * KM], e[KM * KB], xh[KN * KB];
};

__device__ __forceinline__ bool upd_front(UpdShared& s, int p, int n, int m, int nb, const float* y, const float* yhat,
                                          const float* xhat, const float* Lhat, const float* H, const float* R, int lane) {
  w_load(s.Lh, Lhat + (size_t)p * n * n, n * n, lane);
  w_load(s.H, H + (size_t)p * m * n, m * n, lane);
  w_load(s.xh, xhat + (size_t)p * n * nb, n * nb, lane);
  for (int i = lane; i < m * nb; i += 32) s.e[i] = y[(size_t)p * m * nb + i] - yhat[(size_t)p * m * nb + i];
  __syncwarp();
  w_matmul(s.Vh, s.Lh, s.Lh, n, n, n, true, lane);     // Vhat = Lhat Lhat^T
  w_matmul(s.HL, s.H, s.Lh, m, n, n, false, lane);     // H Lhat
  w_matmul(s.S, s.HL, s.HL, m, m, n, true, lane);      // HL HL^T
  for (int i = lane; i < m * m; i += 32) s.S[i] += R[(size_t)p * m * m + i];
  __syncwarp();
  return w_cholesky(s.S, m, lane);
}

// kalman.update (vjf/kalman.py:53-99), cholesky=True
__global__ void kalman_update_kernel(int P, int n, int m, int nb, const float* y, const float* yhat, const float* xhat,
                                     const float* Lhat, const float* H, const float* R, float* x_out, float* L_out,
                                     int* info) {
  __shared__ UpdShared s;
  __shared__ float X[KM * KN], G[KN * KM], V[KN * KN], xo[KN * KB];
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= P) return;
  int st = upd_front(s, p, n, m, nb, y, yhat, xhat, Lhat, H, R, lane) ? 0 : 2;
  w_matmul(X, s.H, s.Vh, m, n, n, false, lane);        // H Vhat
  w_trsm_lower(s.S, X, m, n, lane);                    // L^-1 H Vhat         (:86)
  for (int i = lane; i < n * m; i += 32) { const int r = i / m, c = i - r * m; G[i] = X[c * n + r]; }
  __syncwarp();
  w_trsm_lower(s.S, s.e, m, nb, lane);                 // L^-1 e              (:89)
  w_matmul(xo, G, s.e, n, nb, m, false, lane);
  for (int i = lane; i < n * nb; i += 32) xo[i] += s.xh[i];
  w_matmul(V, G, G, n, n, m, true, lane);              // G G^T
  for (int i = lane; i < n * n; i += 32) V[i] = s.Vh[i] - V[i];   // "minus is dangerous" (:90)
  __syncwarp();
  // keep the raw V in case the factorisation fails: the reference then returns V unfactorised (:94-97)
  for (int i = lane; i < n * n; i += 32) s.HL[i] = V[i];
  __syncwarp();
  if (!w_cholesky(V, n, lane)) { st |= 1; for (int i = lane; i < n * n; i += 32) V[i] = s.HL[i]; __syncwarp(); }
  w_store(x_out + (size_t)p * n * nb, xo, n * nb, lane);
  w_store(L_out + (size_t)p * n * n, V, n * n, lane);
  if (lane == 0 && info) info[p] = st;
}

// kalman.joseph_update (vjf/kalman.py:102-145) AS WRITTEN: G = (S^-1 H Vhat)^T, and S^-1 is applied again
// to e, to H and to the elementwise sqrt(R) (:136-140).
__global__ void kalman_joseph_kernel(int P, int n, int m, int nb, const float* y, const float* yhat, const float* xhat,
                                     const float* Lhat, const float* H, const float* R, float* x_out, float* L_out,
                                     int* info) {
  __shared__ UpdShared s;
  __shared__ float X[KM * KM], G[KN * KM], V[KN * KN], T1[KN * KN], T2[KN * KM], xo[KN * KB];
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= P) return;
  int st = upd_front(s, p, n, m, nb, y, yhat, xhat, Lhat, H, R, lane) ? 0 : 2;
  // G = (S^-1 H Vhat)^T   (:135)
  w_matmul(X, s.H, s.Vh, m, n, n, false, lane);
  w_trsm_lower(s.S, X, m, n, lane);
  w_trsm_lower_t(s.S, X, m, n, lane);
  for (int i = lane; i < n * m; i += 32) { const int r = i / m, c = i - r * m; G[i] = X[c * n + r]; }
  __syncwarp();
  // x = xhat + G S^-1 e   (:136)
  w_trsm_lower(s.S, s.e, m, nb, lane);
  w_trsm_lower_t(s.S, s.e, m, nb, lane);
  w_matmul(xo, G, s.e, n, nb, m, false, lane);
  for (int i = lane; i < n * nb; i += 32) xo[i] += s.xh[i];
  __syncwarp();
  // I - G S^-1 H   (:139)
  for (int i = lane; i < m * n; i += 32) X[i] = s.H[i];
  __syncwarp();
  w_trsm_lower(s.S, X, m, n, lane);
  w_trsm_lower_t(s.S, X, m, n, lane);
  w_matmul(T1, G, X, n, n, m, false, lane);
  for (int i = lane; i < n * n; i += 32) T1[i] = ((i / n == i % n) ? 1.f : 0.f) - T1[i];
  __syncwarp();
  w_matmul(V, T1, s.Lh, n, n, n, false, lane);          // (I-KH) Lhat   (:140)
  w_matmul(T1, V, V, n, n, n, true, lane);
  // K R : G S^-1 sqrt(R) elementwise sqrt   (:141)
  for (int i = lane; i < m * m; i += 32) X[i] = sqrtf(R[(size_t)p * m * m + i]);
  __syncwarp();
  w_trsm_lower(s.S, X, m, m, lane);
  w_trsm_lower_t(s.S, X, m, m, lane);
  w_matmul(T2, G, X, n, m, m, false, lane);
  w_matmul(V, T2, T2, n, n, m, true, lane);
  for (int i = lane; i < n * n; i += 32) V[i] += T1[i];
  __syncwarp();
  if (!w_cholesky(V, n, lane)) st |= 1;
  w_store(x_out + (size_t)p * n * nb, xo, n * nb, lane);
  w_store(L_out + (size_t)p * n * n, V, n * n, lane);
  if (lane == 0 && info) info[p] = st;
}

// numerical.symmetrize (vjf/numerical.py:17-19): a.triu() + a.triu(1)^T
__global__ void symmetrize_kernel(int P, int n, const float* a, float* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P * n * n) return;
  const long long p = i / (n * n);
  const int e = (int)(i - p * n * n), r = e / n, c = e - r * n;
  out[i] = (c >= r) ? a[i] : a[p * n * n + c * n + r];
}

// numerical.positivize (vjf/numerical.py:8-14): eigh (lower triangle), clamp eigenvalues >= eps, rebuild.
// Cyclic Jacobi on the symmetric matrix; one warp per problem.
__global__ void positivize_kernel(int P, int n, const float* a, float eps, float* out) {
  __shared__ float A[KN * KN], V[KN * KN];
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= P) return;
  for (int i = lane; i < n * n; i += 32) {
    const int r = i / n, c = i - r * n;
    A[i] = (c <= r) ? a[(size_t)p * n * n + i] : a[(size_t)p * n * n + c * n + r];  // eigh reads the lower triangle
    V[i] = (r == c) ? 1.f : 0.f;
  }
  __syncwarp();
  for (int sweep = 0; sweep < 30; ++sweep) {
    float off = 0.f;
    for (int i = lane; i < n * n; i += 32) { const int r = i / n, c = i - r * n; if (r != c) off += A[i] * A[i]; }
    off = warp_sum(off);
    if (off < 1e-30f) break;
    for (int pi = 0; pi < n - 1; ++pi)
      for (int qi = pi + 1; qi < n; ++qi) {
        const float apq = A[pi * n + qi];
        if (fabsf(apq) > 1e-30f) {
          const float app = A[pi * n + pi], aqq = A[qi * n + qi];
          const float tau = (aqq - app) / (2.f * apq);
          const float tt = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
          const float c = 1.f / sqrtf(1.f + tt * tt), sn = tt * c;
          __syncwarp();
          for (int k = lane; k < n; k += 32) {  // columns p, q
            const float akp = A[k * n + pi], akq = A[k * n + qi];
            A[k * n + pi] = c * akp - sn * akq; A[k * n + qi] = sn * akp + c * akq;
            const float vkp = V[k * n + pi], vkq = V[k * n + qi];
            V[k * n + pi] = c * vkp - sn * vkq; V[k * n + qi] = sn * vkp + c * vkq;
          }
          __syncwarp();
          for (int k = lane; k < n; k += 32) {  // rows p, q
            const float apk = A[pi * n + k], aqk = A[qi * n + k];
            A[pi * n + k] = c * apk - sn * aqk; A[qi * n + k] = sn * apk + c * aqk;
          }
          __syncwarp();
        }
      }
  }
  for (int i = lane; i < n * n; i += 32) {
    const int r = i / n, c = i - r * n;
    float s = 0.f;
    for (int k = 0; k < n; ++k) s = fmaf(V[r * n + k] * fmaxf(A[k * n + k], eps), V[c * n + k], s);
    out[(size_t)p * n * n + i] = s;
  }
}

int check_dims(int P, int n, int m, int nb) {
  if (P < 1 || n < 1 || n > KN || m < 1 || m > KM || nb < 1 || nb > KB) {
    vjf_set_error("kalman operator limits: 1<=n<=%d, 1<=m<=%d, 1<=nb<=%d, P>=1", KN, KM, KB);
    return -1;
  }
  return 0;
}
}  // namespace

extern "C" int vjf_kalman_predict_batched(int32_t P, int32_t n, int32_t m, int32_t nb, const float* x, const float* L,
                                          const float* A, const float* Q, const float* H, float* yhat, float* xhat,
                                          float* Lhat, int32_t* info, void* stream) {
  if (check_dims(P, n, m, nb)) return -1;
  kalman_predict_kernel<<<P, 32, 0, (cudaStream_t)stream>>>(P, n, m, nb, x, L, A, Q, H, yhat, xhat, Lhat, info);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vjf_kalman_update_batched(int32_t P, int32_t n, int32_t m, int32_t nb, const float* y, const float* yhat,
                                         const float* xhat, const float* Lhat, const float* H, const float* R,
                                         float* x_out, float* L_out, int32_t* info, void* stream) {
  if (check_dims(P, n, m, nb)) return -1;
  kalman_update_kernel<<<P, 32, 0, (cudaStream_t)stream>>>(P, n, m, nb, y, yhat, xhat, Lhat, H, R, x_out, L_out, info);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vjf_kalman_joseph_update_batched(int32_t P, int32_t n, int32_t m, int32_t nb, const float* y,
                                                const float* yhat, const float* xhat, const float* Lhat, const float* H,
                                                const float* R, float* x_out, float* L_out, int32_t* info, void* stream) {
  if (check_dims(P, n, m, nb)) return -1;
  kalman_joseph_kernel<<<P, 32, 0, (cudaStream_t)stream>>>(P, n, m, nb, y, yhat, xhat, Lhat, H, R, x_out, L_out, info);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vjf_symmetrize_batched(int32_t P, int32_t n, const float* a, float* out, void* stream) {
  if (P < 1 || n < 1) { vjf_set_error("bad argument"); return -1; }
  const long long tot = (long long)P * n * n;
  symmetrize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, n, a, out);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vjf_positivize_batched(int32_t P, int32_t n, const float* a, float eps, float* out, void* stream) {
  if (P < 1 || n < 1 || n > KN) { vjf_set_error("positivize: 1<=n<=%d", KN); return -1; }
  positivize_kernel<<<P, 32, 0, (cudaStream_t)stream>>>(P, n, a, eps, out);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}
