// Validation of hand-written tcgen05.mma (kind::tf32, cta_group::1, M=64) with no-swizzle canonical smem layouts:
//   test 1: D[64 x 32]  = A[64 x 208] * B^T,  A K-major, B K-major   (recognition layer-1 forward shape)
//   test 2: D[64 x 208] = A[64 x 32]  * B^T,  A K-major, B MN-major  (weight-gradient shape; B is the SAME array as test 1's B)
// Results are read back from TMEM with tcgen05.ld and compared with a CPU reference (tf32-truncated inputs).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type [61,64)=0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6) | a_format TF32=2 [7,10) | b_format TF32=2 [10,13) | a_major [15] | b_major [16] | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
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

// canonical K-major no-swizzle index: element (r, k) of an [MN x K] operand; uint128 columns at stride MN
__host__ __device__ inline int canon(int r, int k, int MN) { return ((k >> 2) * MN + r) * 4 + (k & 3); }

constexpr int M = 64, NB = 32, K1 = 208;

__global__ void __launch_bounds__(128, 1) k(const float* A1, const float* Bsrc, const float* A2, float* D1, float* D2, int variant) {
  extern __shared__ __align__(1024) float sm[];
  float* sA1 = sm;                         // [64 x 208]  canonical K-major (MN = 64)
  float* sB = sA1 + M * K1;                // [32 x 208]  canonical K-major (MN = 32)  == MN-major view [208 x 32] for test 2
  float* sA2 = sB + NB * K1;               // [64 x 32]   canonical K-major (MN = 64)
  float* sBt = sA2 + M * NB;               // [208 x 32]  canonical K-major (MN = 208): transposed copy of B for variant 3
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * K1; i += 128) { int r = i / K1, kk = i % K1; sA1[canon(r, kk, M)] = A1[i]; }
  for (int i = tid; i < NB * K1; i += 128) { int r = i / K1, kk = i % K1; sB[canon(r, kk, NB)] = Bsrc[i]; }
  for (int i = tid; i < M * NB; i += 128) { int r = i / NB, kk = i % NB; sA2[canon(r, kk, M)] = A2[i]; }
  for (int i = tid; i < NB * K1; i += 128) { int b = i / K1, n = i % K1; sBt[canon(n, b, K1)] = Bsrc[i]; }
  float* sA2mn = sBt + NB * K1;            // [64 x 32] MN-major: uint128 index (m/4)*32 + k  (k in 0..31) -> SBO(chunk)=32 units, LBO(k-group)=8 units
  for (int i = tid; i < M * NB; i += 128) { int m = i / NB, kk = i % NB; sA2mn[((m >> 2) * 32 + kk) * 4 + (m & 3)] = A2[i]; }
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    // ---- test 1: D1 (cols [0,32)) = A1 * B^T : K = 208 -> 26 MMAs of K = 8 ----
    const uint32_t idesc1 = make_idesc(M, NB, 0, 0);
    for (int s = 0; s < K1 / 8; ++s) {
      // K-major: LBO = byte stride between the two 16-byte K-halves of one MMA = MN*16 ; SBO = stride between 8-row groups = 128
      const uint64_t ad = make_desc(smem_u32(sA1) + s * 2 * M * 16, M * 16, 128);
      const uint64_t bd = make_desc(smem_u32(sB) + s * 2 * NB * 16, NB * 16, 128);
      umma_tf32(tm + 0, ad, bd, idesc1, s > 0);
    }
    // ---- test 2: D2 (cols [32, 240)) = A2 * Bmn : M = 64, N = 208, K = 32 -> 4 MMAs; B MN-major over the same array ----
    const int N2 = (variant == 4) ? 64 : K1;
    const uint32_t idesc2 = (variant == 5 || variant == 6) ? make_idesc(M, N2, 1, 0) : make_idesc(M, N2, 0, variant == 3 ? 0 : 1);
    for (int s = 0; s < NB / 8; ++s) {
      const uint64_t ad = (variant == 5) ? make_desc(smem_u32(sA2mn) + s * 8 * 16, 8 * 16, 32 * 16)
                        : (variant == 6) ? make_desc(smem_u32(sA2mn) + s * 8 * 16, 32 * 16, 8 * 16)
                                         : make_desc(smem_u32(sA2) + s * 2 * M * 16, M * 16, 128);
      // MN-major: element (n, k) at uint128 index (n/4)*SBO + (k%8) + (k/8)*LBO ; here SBO = 32 units (=NB), LBO = 8 units
      const uint64_t bd = (variant == 3 || variant == 5 || variant == 6) ? make_desc(smem_u32(sBt) + s * 2 * K1 * 16, K1 * 16, 128) : (variant == 0 || variant == 4) ? make_desc(smem_u32(sB) + s * 8 * 16, 8 * 16, NB * 16)
                         : (variant == 1) ? make_desc(smem_u32(sB) + s * 8 * 16, NB * 16, 8 * 16)
                                          : make_desc(smem_u32(sB) + s * 8 * 16, NB * 16, NB * 16);
      umma_tf32(tm + 32, ad, bd, idesc2, s > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D row m lives in TMEM lane (m % 16) + 32 * (m / 16); warp w reads lanes [32w, 32w+32)
  for (int c0 = 0; c0 < 240; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (lane < 16) {
      const int m = warp * 16 + lane;
      for (int j = 0; j < 16; ++j) {
        const int c = c0 + j;
        if (c < 32) D1[m * 32 + c] = __uint_as_float(r[j]);
        else if (c < 240) D2[m * K1 + (c - 32)] = __uint_as_float(r[j]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
#include <cstring>
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  float *hA1 = new float[M * K1], *hB = new float[NB * K1], *hA2 = new float[M * NB];
  srand(1);
  for (int i = 0; i < M * K1; ++i) hA1[i] = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
  for (int i = 0; i < NB * K1; ++i) hB[i] = tf32_trunc((rand() % 2001 - 1000) / 500.f);
  for (int i = 0; i < M * NB; ++i) hA2[i] = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
  float *dA1, *dB, *dA2, *dD1, *dD2;
  cudaMalloc(&dA1, M * K1 * 4); cudaMalloc(&dB, NB * K1 * 4); cudaMalloc(&dA2, M * NB * 4); cudaMalloc(&dD1, M * 32 * 4); cudaMalloc(&dD2, M * K1 * 4);
  cudaMemcpy(dA1, hA1, M * K1 * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, NB * K1 * 4, cudaMemcpyHostToDevice); cudaMemcpy(dA2, hA2, M * NB * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD1, 0, M * 32 * 4); cudaMemset(dD2, 0, M * K1 * 4);
  const size_t smem = (M * K1 + 2 * NB * K1 + 2 * M * NB) * 4 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<1, 128, smem>>>(dA1, dB, dA2, dD1, dD2, variant);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  float *hD1 = new float[M * 32], *hD2 = new float[M * K1];
  cudaMemcpy(hD1, dD1, M * 32 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hD2, dD2, M * K1 * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, n1 = 0, n2 = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < NB; ++n) { double s = 0; for (int kk = 0; kk < K1; ++kk) s += (double)hA1[m * K1 + kk] * hB[n * K1 + kk]; e1 = fmax(e1, fabs(s - hD1[m * 32 + n])); n1 = fmax(n1, fabs(s)); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < K1; ++n) { double s = 0; for (int b = 0; b < NB; ++b) s += (double)hA2[m * NB + b] * hB[b * K1 + n]; e2 = fmax(e2, fabs(s - hD2[m * K1 + n])); n2 = fmax(n2, fabs(s)); }
  printf("test1 (K-major x K-major, 64x32x208): max abs err %.3e (max |ref| %.2f)\n", e1, n1);
  printf("test2 (K-major x MN-major, 64x208x32): max abs err %.3e (max |ref| %.2f)\n", e2, n2);
  printf("sample D1[0][0..3] = %f %f %f %f\n", hD1[0], hD1[1], hD1[2], hD1[3]);
  printf("D2[0][0..7]   = "); for (int n = 0; n < 8; ++n) printf("%9.4f ", hD2[n]); printf("\nref           = ");
  for (int n = 0; n < 8; ++n) { double s = 0; for (int b = 0; b < NB; ++b) s += (double)hA2[0 * NB + b] * hB[b * K1 + n]; printf("%9.4f ", s); } printf("\n");
  printf("D2[5][100..107]= "); for (int n = 100; n < 108; ++n) printf("%9.4f ", hD2[5 * K1 + n]); printf("\nref           = ");
  for (int n = 100; n < 108; ++n) { double s = 0; for (int b = 0; b < NB; ++b) s += (double)hA2[5 * NB + b] * hB[b * K1 + n]; printf("%9.4f ", s); } printf("\n");
  return 0;
}
