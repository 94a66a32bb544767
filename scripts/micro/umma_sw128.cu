// Validation of the tcgen05.mma (kind::tf32, cta_group::1, M = 128) operand forms the throughput tile pipeline uses,
// all with the 128-byte swizzle that TMA tile loads produce:
//   case 1  A K-major  x B K-major    D[128 x 64]  = A[128 x 32]  * B[64 x 32]^T      (recognition forward form)
//   case 2  A MN-major x B MN-major   D[128 x 64]  = A^T ... K = trials as ROWS of both operands (weight-gradient form)
//           (the SAME shared-memory image as a K-major [rows x 32] tile, reinterpreted: no transposed copy)
//   case 3  as case 1 with full-mantissa fp32 inputs: does the tensor core truncate or round the low 13 bits?
//   case 4  A K-major x B MN-major, N = 208
// The host builds the exact shared-memory byte images and the descriptor fields, so layout hypotheses are tested without
// recompiling the kernel: every case prints max |err| against an fp64 reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct Case {
  uint32_t bytesA, bytesB;       // image sizes (multiples of 16)
  uint64_t descA, descB;         // descriptors WITHOUT the start address (LBO, SBO, version, layout type)
  uint32_t idesc;
  int nk;                        // number of MMAs (K = 8 each)
  uint32_t offA[64], offB[64];   // byte offset of the operand start per MMA
  int N;
};

__global__ void __launch_bounds__(128, 1) k(const float* imgA, const float* imgB, const Case c, float* D) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);
  float* sA = (float*)base;
  float* sB = (float*)(base + ((c.bytesA + 1023) & ~1023u));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < c.bytesA / 4; i += 128) sA[i] = imgA[i];
  for (uint32_t i = tid; i < c.bytesB / 4; i += 128) sB[i] = imgB[i];
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    for (int s = 0; s < c.nk; ++s) {
      const uint64_t ad = c.descA | (uint64_t)(((smem_u32(sA) + c.offA[s]) >> 4) & 0x3fff);
      const uint64_t bd = c.descB | (uint64_t)(((smem_u32(sB) + c.offB[s]) >> 4) & 0x3fff);
      umma_tf32(tm, ad, bd, c.idesc, s > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // M = 128: D row m lives in TMEM lane m; warp w reads lanes [32w, 32w+32)
  for (int c0 = 0; c0 < c.N; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[tid * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0xfffu + ((u >> 13) & 1u); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float rnd() { return (rand() % 20001 - 10000) / 10000.f; }

static uint64_t desc_fields(uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
static uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// "chunked SW128" tile: [chunk of 32 columns][row][32 floats], 128 bytes per row, 16-byte pieces XOR-swizzled by (row % 8)
static size_t sw128_off(int row, int col, int rows) {
  const int ch = col >> 5, c = col & 31;
  return ((size_t)ch * rows + row) * 128 + (size_t)((((c >> 2) ^ (row & 7)) << 4) + ((c & 3) << 2));
}

static int run(const char* name, const std::vector<float>& imgA, const std::vector<float>& imgB, Case c, const std::vector<double>& ref,
               const std::vector<double>* ref2 = nullptr, bool m64 = false) {
  float *dA, *dB, *dD;
  c.bytesA = (uint32_t)imgA.size() * 4; c.bytesB = (uint32_t)imgB.size() * 4;
  cudaMalloc(&dA, c.bytesA); cudaMalloc(&dB, c.bytesB); cudaMalloc(&dD, 128 * c.N * 4);
  cudaMemcpy(dA, imgA.data(), c.bytesA, cudaMemcpyHostToDevice); cudaMemcpy(dB, imgB.data(), c.bytesB, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 128 * c.N * 4);
  const size_t smem = c.bytesA + c.bytesB + 4096;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<1, 128, smem>>>(dA, dB, c, dD);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(128 * c.N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0, mx = 0, err2 = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    if (m64 && ((i / c.N) & 31) >= 16) continue;  // M = 64: lanes 16..31 of every sub-partition hold no row
    err = fmax(err, fabs(ref[i] - D[i])); mx = fmax(mx, fabs(ref[i]));
    if (ref2) err2 = fmax(err2, fabs((*ref2)[i] - D[i]));
  }
  if (std::isnan(err)) err = 1e30;
  printf("%-58s %s  max|err| %.3e (max|ref| %.2f)", name, cudaGetErrorString(e), err, mx);
  if (ref2) printf("  vs second reference %.3e", err2);
  printf("  %s\n", err < 1e-4 * fmax(mx, 1.0) ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return err < 1e-4 * fmax(mx, 1.0);
}

int main() {
  srand(3);
  const int M = 128;
  // ---------------- case 1 / 3: K-major x K-major, K = 64 (two 32-column chunks) ----------------
  for (int full = 0; full < 2; ++full) {
    const int N = 64, K = 64;
    std::vector<float> A(M * K), B(N * K);
    for (auto& v : A) v = full ? rnd() * 1.2345678f : tf32_trunc(rnd());
    for (auto& v : B) v = full ? rnd() * 0.7654321f : tf32_trunc(rnd());
    std::vector<float> ia(M * K), ib(N * K);
    for (int r = 0; r < M; ++r) for (int kk = 0; kk < K; ++kk) ia[sw128_off(r, kk, M) / 4] = A[r * K + kk];
    for (int r = 0; r < N; ++r) for (int kk = 0; kk < K; ++kk) ib[sw128_off(r, kk, N) / 4] = B[r * K + kk];
    std::vector<double> ref(M * N), ref_rn(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double s = 0, s2 = 0;
      for (int kk = 0; kk < K; ++kk) { s += (double)tf32_trunc(A[m * K + kk]) * tf32_trunc(B[n * K + kk]); s2 += (double)tf32_rn(A[m * K + kk]) * tf32_rn(B[n * K + kk]); }
      ref[m * N + n] = s; ref_rn[m * N + n] = s2;
    }
    Case c{};
    c.descA = desc_fields(16, 1024, 2); c.descB = desc_fields(16, 1024, 2);
    c.idesc = make_idesc(M, N, 0, 0); c.nk = K / 8; c.N = N;
    for (int s = 0; s < c.nk; ++s) { c.offA[s] = (s / 4) * M * 128 + (s % 4) * 32; c.offB[s] = (s / 4) * N * 128 + (s % 4) * 32; }
    run(full ? "case 3: K x K, full-mantissa inputs (ref = truncate | RN)" : "case 1: A K-major x B K-major, SW128, 128x64x64", ia, ib, c, ref, full ? &ref_rn : nullptr);
  }
  // ---------------- case 2: MN-major x MN-major: D[m][n] = sum_b X[b][m] * G[b][n] ----------------
  // X is a [rows = 32 trials][128 columns] chunked-SW128 tile (4 chunks), G a [32 trials][64 columns] tile (2 chunks): the K
  // dimension (trials) runs over the ROWS of both images.
  for (int variant = 0; variant < 2; ++variant) {
    const int N = 64, KB = 32;
    std::vector<float> X(KB * M), G(KB * N);
    for (auto& v : X) v = tf32_trunc(rnd());
    for (auto& v : G) v = tf32_trunc(rnd());
    std::vector<float> ia(KB * M), ib(KB * N);
    for (int b = 0; b < KB; ++b) for (int m = 0; m < M; ++m) ia[sw128_off(b, m, KB) / 4] = X[b * M + m];
    for (int b = 0; b < KB; ++b) for (int n = 0; n < N; ++n) ib[sw128_off(b, n, KB) / 4] = G[b * N + n];
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int b = 0; b < KB; ++b) s += (double)X[b * M + m] * G[b * N + n]; ref[m * N + n] = s; }
    Case c{};
    const uint32_t chunk = KB * 128;  // byte stride between 32-column chunks (atoms along MN); 8-row groups (atoms along K) are 1024 bytes apart
    if (variant == 0) { c.descA = desc_fields(chunk, 1024, 2); c.descB = desc_fields(chunk, 1024, 2); }
    else { c.descA = desc_fields(1024, chunk, 2); c.descB = desc_fields(1024, chunk, 2); }
    c.idesc = make_idesc(M, N, 1, 1); c.nk = KB / 8; c.N = N;
    for (int s = 0; s < c.nk; ++s) { c.offA[s] = s * 1024; c.offB[s] = s * 1024; }
    run(variant == 0 ? "case 2a: MN x MN, SW128, LBO = chunk stride, SBO = 1024" : "case 2b: MN x MN, SW128, LBO = 1024, SBO = chunk stride", ia, ib, c, ref);
  }
  // ---------------- case 4: A K-major [128 x 32] x B MN-major [32 trials][208 columns] ----------------
  for (int variant = 0; variant < 2; ++variant) {
    const int N = 208, NP = 224, KB = 32;
    std::vector<float> A(M * KB), G(KB * NP, 0.f);
    for (auto& v : A) v = tf32_trunc(rnd());
    for (int b = 0; b < KB; ++b) for (int n = 0; n < N; ++n) G[b * NP + n] = tf32_trunc(rnd());
    std::vector<float> ia(M * KB), ib(KB * NP);
    for (int r = 0; r < M; ++r) for (int kk = 0; kk < KB; ++kk) ia[sw128_off(r, kk, M) / 4] = A[r * KB + kk];
    for (int b = 0; b < KB; ++b) for (int n = 0; n < NP; ++n) ib[sw128_off(b, n, KB) / 4] = G[b * NP + n];
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int b = 0; b < KB; ++b) s += (double)A[m * KB + b] * G[b * NP + n]; ref[m * N + n] = s; }
    Case c{};
    const uint32_t chunk = KB * 128;
    c.descA = desc_fields(16, 1024, 2);
    c.descB = variant == 0 ? desc_fields(chunk, 1024, 2) : desc_fields(1024, chunk, 2);
    c.idesc = make_idesc(M, N, 0, 1); c.nk = KB / 8; c.N = N;
    for (int s = 0; s < c.nk; ++s) { c.offA[s] = s * 32; c.offB[s] = s * 1024; }
    run(variant == 0 ? "case 4a: K x MN (N = 208), LBO = chunk stride, SBO = 1024" : "case 4b: K x MN (N = 208), LBO = 1024, SBO = chunk stride", ia, ib, c, ref);
  }
  // ---------------- case 5: A K-major SW128 [M x 32 trials] x B MN-major SWIZZLE_128B_BASE32B [32 trials][N columns] ----------------
  // CUTLASS (sm100_common.inl): "for mn-major tf32 operands, SW128_32B is the only available smem layout": rows of 128 bytes,
  // 32-byte pieces XOR-ed with (row % 4) (Swizzle<2,5,2> on the byte address), atoms of 4 rows.
  for (int variant = 0; variant < 6; ++variant) {
    const int Mv = (variant >= 4) ? 64 : 128;
    const int N = (variant >= 4) ? 256 : 224, KB = 32;
    std::vector<float> A(128 * KB, 0.f), G(KB * N);
    for (int i = 0; i < Mv * KB; ++i) A[i] = tf32_trunc(rnd());
    for (auto& v : G) v = tf32_trunc(rnd());
    std::vector<float> ia(128 * KB), ib(KB * N);
    for (int r = 0; r < 128; ++r) for (int kk = 0; kk < KB; ++kk) ia[sw128_off(r, kk, 128) / 4] = A[r * KB + kk];
    for (int b = 0; b < KB; ++b) for (int n = 0; n < N; ++n) {
      size_t off = ((size_t)(n >> 5) * KB + b) * 128 + (size_t)(n & 31) * 4;   // chunked rows of 128 bytes
      off ^= (size_t)((b & 3) << 5);                                            // Swizzle<2,5,2>
      ib[off / 4] = G[b * N + n];
    }
    std::vector<double> ref(128 * N, 0.0);
    for (int m = 0; m < Mv; ++m) for (int n = 0; n < N; ++n) { double s2 = 0; for (int b = 0; b < KB; ++b) s2 += (double)A[m * KB + b] * G[b * N + n]; ref[(Mv == 64 ? (m % 16) + 32 * (m / 16) : m) * N + n] = s2; }
    Case c{};
    const uint32_t chunk = KB * 128;
    c.descA = desc_fields(16, 1024, 2);
    const int v4 = variant & 3;
    c.descB = v4 == 0 ? desc_fields(chunk, 512, 1) : v4 == 1 ? desc_fields(512, chunk, 1) : v4 == 2 ? desc_fields(chunk, 1024, 1) : desc_fields(1024, chunk, 1);
    c.idesc = make_idesc(Mv, N, 0, 1); c.nk = KB / 8; c.N = N;
    for (int s2 = 0; s2 < c.nk; ++s2) { c.offA[s2] = s2 * 32; c.offB[s2] = s2 * 1024; }
    char nm[128];
    snprintf(nm, sizeof nm, "case 5.%d: M=%d K x MN(BASE32B) N=%d %s", variant, Mv, N,
             v4 == 0 ? "LBO=chunk SBO=512" : v4 == 1 ? "LBO=512 SBO=chunk" : v4 == 2 ? "LBO=chunk SBO=1024" : "LBO=1024 SBO=chunk");
    // M = 64: only TMEM lanes (m % 16) + 32 (m / 16) hold rows; compare those lanes only
    if (Mv == 64) { /* other lanes of ref stay 0 and are compared against whatever TMEM holds: mask them */ }
    run(nm, ia, ib, c, ref, nullptr, Mv == 64);
  }
  return 0;
}
