// Second validation pass of tcgen05.mma kind::tf32 operand forms for the throughput tile pipeline (see umma_sw128.cu for the
// first): which shared-memory IMAGES can serve which operand roles, and what images the TMA swizzle modes produce.
//   case 6   A MN-major SWIZZLE_128B_BASE32B  x  B K-major SW128            (weight-gradient form with the trials as K)
//   case 7   A K-major read of a BASE32B image (layout type 1)  x  B K-major SW128
//   case 8   A K-major SW128  x  B K-major read of a BASE32B image
//   case 11  A MN-major BASE32B x B MN-major BASE32B, K = 128 trials (16 k-steps), 128-row images
//   case 9   TMA tensor-map loads: image of CU_TENSOR_MAP_SWIZZLE_128B and of ..._128B_ATOM_32B against the host formulas
// Host builds the byte images and the descriptor fields; every case prints max |err| against an fp64 reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda.h>
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
  uint32_t bytesA, bytesB;
  uint64_t descA, descB;
  uint32_t idesc;
  int nk;
  uint32_t offA[64], offB[64];
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

// TMA: load one box {32 floats, rows} of a row-major [rows_total][ld] matrix at (col0, row0) into shared memory and dump it
__global__ void __launch_bounds__(128, 1) ktma(const __grid_constant__ CUtensorMap map, int col0, int row0, int bytes, float* out) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(base)), "l"(&map), "r"(col0), "r"(row0), "r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes / 4; i += 128) out[i] = ((float*)base)[i];
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
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
// image offsets (bytes) of element (row, col) of a [rows x cols] fp32 tile stored as 32-column chunks of [rows][128 B]
static size_t sw128_off(int row, int col, int rows) {      // 16-byte pieces XOR (row % 8)
  const int ch = col >> 5, c = col & 31;
  return ((size_t)ch * rows + row) * 128 + (size_t)((((c >> 2) ^ (row & 7)) << 4) + ((c & 3) << 2));
}
static size_t b32_off(int row, int col, int rows) {        // 32-byte pieces XOR (row % 4)  (Swizzle<2,5,2>)
  size_t off = ((size_t)(col >> 5) * rows + row) * 128 + (size_t)(col & 31) * 4;
  return off ^ (size_t)((row & 3) << 5);
}

static int run(const char* name, const std::vector<float>& imgA, const std::vector<float>& imgB, Case c, const std::vector<double>& ref) {
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
  double err = 0, mx = 0;
  for (size_t i = 0; i < D.size(); ++i) { err = fmax(err, fabs(ref[i] - D[i])); mx = fmax(mx, fabs(ref[i])); }
  if (std::isnan(err)) err = 1e30;
  printf("%-72s %s  max|err| %.3e (max|ref| %.2f)  %s\n", name, cudaGetErrorString(e), err, mx, err < 1e-4 * fmax(mx, 1.0) ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return err < 1e-4 * fmax(mx, 1.0);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  srand(5);
  const int M = 128;
  const int only = argc > 1 ? atoi(argv[1]) : 0;  // run one case per process: a faulting case poisons the context

  // ---------------- case 6: A MN-major BASE32B [KB trials][128 m] x B K-major SW128 [64 n][KB] ----------------
  if (only == 0 || only == 6)
  for (int variant = 0; variant < 3; ++variant) {
    const int N = 64, KB = 32;
    std::vector<float> X(KB * M), Bm(N * KB);
    for (auto& v : X) v = tf32_trunc(rnd());
    for (auto& v : Bm) v = tf32_trunc(rnd());
    std::vector<float> ia(KB * M), ib(N * KB);
    for (int b = 0; b < KB; ++b) for (int m = 0; m < M; ++m) ia[b32_off(b, m, KB) / 4] = X[b * M + m];
    for (int n = 0; n < N; ++n) for (int kk = 0; kk < KB; ++kk) ib[sw128_off(n, kk, N) / 4] = Bm[n * KB + kk];
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int b = 0; b < KB; ++b) s += (double)X[b * M + m] * Bm[n * KB + b]; ref[m * N + n] = s; }
    Case c{};
    const uint32_t chunk = KB * 128;
    c.descA = variant == 0 ? desc_fields(chunk, 512, 1) : variant == 1 ? desc_fields(512, chunk, 1) : desc_fields(chunk, 1024, 1);
    c.descB = desc_fields(16, 1024, 2);
    c.idesc = make_idesc(M, N, 1, 0); c.nk = KB / 8; c.N = N;
    for (int s = 0; s < c.nk; ++s) { c.offA[s] = s * 1024; c.offB[s] = s * 32; }
    char nm[160];
    snprintf(nm, sizeof nm, "case 6.%d: A MN(BASE32B) x B K(SW128) %s", variant, variant == 0 ? "LBO=chunk SBO=512" : variant == 1 ? "LBO=512 SBO=chunk" : "LBO=chunk SBO=1024");
    run(nm, ia, ib, c, ref);
  }
  // ---------------- case 7 / 8: K-major reads of a BASE32B image ----------------
  if (only == 7)
  for (int which = 0; which < 2; ++which) {
    for (int variant = 0; variant < 2; ++variant) {
      const int N = 64, K = 64;
      std::vector<float> A(M * K), B(N * K);
      for (auto& v : A) v = tf32_trunc(rnd());
      for (auto& v : B) v = tf32_trunc(rnd());
      std::vector<float> ia(M * K), ib(N * K);
      for (int r = 0; r < M; ++r) for (int kk = 0; kk < K; ++kk) ia[(which == 0 ? b32_off(r, kk, M) : sw128_off(r, kk, M)) / 4] = A[r * K + kk];
      for (int r = 0; r < N; ++r) for (int kk = 0; kk < K; ++kk) ib[(which == 1 ? b32_off(r, kk, N) : sw128_off(r, kk, N)) / 4] = B[r * K + kk];
      std::vector<double> ref(M * N);
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int kk = 0; kk < K; ++kk) s += (double)A[m * K + kk] * B[n * K + kk]; ref[m * N + n] = s; }
      Case c{};
      const uint64_t swz = desc_fields(16, 1024, 2);
      const uint64_t b32 = variant == 0 ? desc_fields(16, 1024, 1) : desc_fields(16, 512, 1);
      c.descA = which == 0 ? b32 : swz; c.descB = which == 1 ? b32 : swz;
      c.idesc = make_idesc(M, N, 0, 0); c.nk = K / 8; c.N = N;
      for (int s = 0; s < c.nk; ++s) { c.offA[s] = (s / 4) * M * 128 + (s % 4) * 32; c.offB[s] = (s / 4) * N * 128 + (s % 4) * 32; }
      char nm[160];
      snprintf(nm, sizeof nm, "case %d.%d: %s K-major read of a BASE32B image, SBO=%d", 7 + which, variant, which == 0 ? "A" : "B", variant == 0 ? 1024 : 512);
      run(nm, ia, ib, c, ref);
    }
  }
  // ---------------- case 11: A MN(BASE32B) [128 trials][128 k1] x B MN(BASE32B) [128 trials][64 n], K = 128 trials ----------------
  if (only == 0 || only == 11) {
    const int N = 64, KB = 128;
    std::vector<float> X(KB * M), G(KB * N);
    for (auto& v : X) v = tf32_trunc(rnd());
    for (auto& v : G) v = tf32_trunc(rnd());
    std::vector<float> ia(KB * M), ib(KB * N);
    for (int b = 0; b < KB; ++b) for (int m = 0; m < M; ++m) ia[b32_off(b, m, KB) / 4] = X[b * M + m];
    for (int b = 0; b < KB; ++b) for (int n = 0; n < N; ++n) ib[b32_off(b, n, KB) / 4] = G[b * N + n];
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int b = 0; b < KB; ++b) s += (double)X[b * M + m] * G[b * N + n]; ref[m * N + n] = s; }
    Case c{};
    const uint32_t chunk = KB * 128;
    c.descA = desc_fields(chunk, 512, 1); c.descB = desc_fields(chunk, 512, 1);
    c.idesc = make_idesc(M, N, 1, 1); c.nk = KB / 8; c.N = N;
    for (int s = 0; s < c.nk; ++s) { c.offA[s] = s * 1024; c.offB[s] = s * 1024; }
    run("case 11: A MN(BASE32B) x B MN(BASE32B), K = 128 trials", ia, ib, c, ref);
  }
  // ---------------- case 12: M = 64: A MN(BASE32B) [128 trials][64 k1] x B MN(BASE32B) [128 trials][224 n] ----------------
  if (only == 0 || only == 12) {
    const int Mv = 64, N = 224, KB = 128;
    std::vector<float> X(KB * Mv), G(KB * N);
    for (auto& v : X) v = tf32_trunc(rnd());
    for (auto& v : G) v = tf32_trunc(rnd());
    std::vector<float> ia(KB * Mv), ib(KB * N);
    for (int b = 0; b < KB; ++b) for (int m = 0; m < Mv; ++m) ia[b32_off(b, m, KB) / 4] = X[b * Mv + m];
    for (int b = 0; b < KB; ++b) for (int n = 0; n < N; ++n) ib[b32_off(b, n, KB) / 4] = G[b * N + n];
    std::vector<double> ref(128 * N, 0.0);
    std::vector<float> dummy;
    for (int m = 0; m < Mv; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int b = 0; b < KB; ++b) s += (double)X[b * Mv + m] * G[b * N + n]; ref[((m % 16) + 32 * (m / 16)) * N + n] = s; }
    Case c{};
    const uint32_t chunk = KB * 128;
    c.descA = desc_fields(chunk, 512, 1); c.descB = desc_fields(chunk, 512, 1);
    c.idesc = make_idesc(Mv, N, 1, 1); c.nk = KB / 8; c.N = N;
    for (int s = 0; s < c.nk; ++s) { c.offA[s] = s * 1024; c.offB[s] = s * 1024; }
    // compare only the lanes that hold rows (others: whatever TMEM held)
    float *dA, *dB, *dD;
    c.bytesA = (uint32_t)ia.size() * 4; c.bytesB = (uint32_t)ib.size() * 4;
    cudaMalloc(&dA, c.bytesA); cudaMalloc(&dB, c.bytesB); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, ia.data(), c.bytesA, cudaMemcpyHostToDevice); cudaMemcpy(dB, ib.data(), c.bytesB, cudaMemcpyHostToDevice);
    const size_t smem = c.bytesA + c.bytesB + 4096;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<1, 128, smem>>>(dA, dB, c, dD);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int l = 0; l < 128; ++l) if ((l & 31) < 16) for (int n = 0; n < N; ++n) err = fmax(err, fabs(ref[l * N + n] - D[l * N + n]));
    printf("%-72s %s  max|err| %.3e  %s\n", "case 12: M=64 A MN(BASE32B) x B MN(BASE32B) N=224, K = 128 trials", cudaGetErrorString(e), err, err < 1e-3 ? "PASS" : "FAIL");
  }
  // ---------------- case 9: what images do the TMA swizzle modes produce? ----------------
  if (only == 0 || only == 9) {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
    if (!encode) { printf("case 9: cuTensorMapEncodeTiled not found\n"); return 0; }
    const int rows_total = 300, ld = 224, rows = 128;
    std::vector<float> g(rows_total * ld);
    for (size_t i = 0; i < g.size(); ++i) g[i] = (float)i;
    float *dg, *dout;
    cudaMalloc(&dg, g.size() * 4); cudaMalloc(&dout, rows * 128);
    cudaMemcpy(dg, g.data(), g.size() * 4, cudaMemcpyHostToDevice);
    const CUtensorMapSwizzle modes[2] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B};
    for (int md = 0; md < 2; ++md) {
      CUtensorMap map;
      const cuuint64_t gdim[2] = {(cuuint64_t)200 /* logical columns: 24 of the last chunk are out of bounds */, (cuuint64_t)rows_total};
      const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
      const cuuint32_t box[2] = {32, (cuuint32_t)rows};
      const cuuint32_t estr[2] = {1, 1};
      CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dg, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[md],
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("case 9.%d: encode failed (%d)\n", md, (int)r); continue; }
      for (int col0 = 64; col0 <= 192; col0 += 128) {
        const int row0 = 40;
        cudaFuncSetAttribute(ktma, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 128 + 2048);
        ktma<<<1, 128, rows * 128 + 2048>>>(map, col0, row0, rows * 128, dout);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> o(rows * 32);
        cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad_sw = 0, bad_b32 = 0;
        for (int rr = 0; rr < rows; ++rr) for (int cc = 0; cc < 32; ++cc) {
          const float want = (col0 + cc < 200) ? g[(size_t)(row0 + rr) * ld + col0 + cc] : 0.f;
          if (o[sw128_off(rr, cc, rows) / 4] != want) ++bad_sw;
          if (o[b32_off(rr, cc, rows) / 4] != want) ++bad_b32;
        }
        printf("case 9.%d: TMA %s box {32, %d} at col %d: %s; mismatches vs sw128 formula %d, vs base32b formula %d\n", md,
               md == 0 ? "SWIZZLE_128B" : "SWIZZLE_128B_ATOM_32B", rows, col0, cudaGetErrorString(e), bad_sw, bad_b32);
      }
    }
  }
  return 0;
}
