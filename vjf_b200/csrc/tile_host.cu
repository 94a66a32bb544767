// Host side of the throughput tile pipeline (tile_kernels.cuh / k_tile.cu): shared-memory / tensor-memory plan, TMA tensor map
// of the observations, launch.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.cuh"
#include "tile_plan.h"

static inline int rup(int v, int a) { return (v + a - 1) / a * a; }

// 0: automatic; 1: never use the tile pipeline; 2: tile pipeline without the exact-observation mode; 3: 64-trial tiles whenever
// the observations are exact (tests drive every variant through the same process)
static int g_tile_mode = 0;
int vjf_tile_mode_get() { return g_tile_mode; }
extern "C" int vjf_set_tile_mode(int32_t mode) {
  if (mode < 0 || mode > 3) { vjf_set_error("tile mode must be 0..3"); return -1; }
  g_tile_mode = mode;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Are all observations exactly representable in tf32 (integers below 2048, i.e. spike counts)?  Then the tensor core's
// truncation of the input image loses nothing, the image needs no "lo" companion except for the appended columns, and the
// tiles can be twice as large.  One pass over the buffer at copy speed; the flag is read back (stream synchronisation).
__global__ void vjf_y_exact_kernel(const float4* __restrict__ y, size_t n4, const float* __restrict__ tail, int ntail, unsigned* flag) {
  unsigned bad = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = y[i];
    bad |= (__float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w)) & 0x1fffu;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) bad |= __float_as_uint(tail[threadIdx.x]) & 0x1fffu;
  if (__any_sync(0xffffffffu, bad != 0) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

int vjf_observations_exact(vjf_handle* h, const void* y, size_t n, cudaStream_t s, bool* exact) {
  unsigned* flag = h->sync_words + 48;
  VJF_CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(unsigned), s));
  const size_t n4 = n / 4;
  const int blocks = (int)std::min<size_t>((size_t)h->num_sms * 8, std::max<size_t>(1, (n4 + 255) / 256));
  vjf_y_exact_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(y), n4, reinterpret_cast<const float*>(y) + n4 * 4, (int)(n - n4 * 4), flag);
  ++g_vjf_launches;
  unsigned host = 1;
  VJF_CUDA_OK(cudaMemcpyAsync(&host, flag, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  VJF_CUDA_OK(cudaStreamSynchronize(s));
  *exact = host == 0;
  return 0;
}

// dimensions that do not depend on the batch
static bool tile_static_dims(const StepParams& p, TilePlan& pl) {
  const int H = p.H[0];
  if (p.L != 1 || H % 32 != 0 || H > 128 || p.d > 8 || p.D % 4 != 0) return false;
  pl.K1b = p.K1 + 1;
  pl.NCH = (pl.K1b + 31) / 32;
  pl.NCY = (p.D + 31) / 32;
  pl.Rk = rup(p.R, 8);
  pl.NQ = rup(pl.Rk + p.d, 16);
  pl.PW = rup(pl.Rk + p.d, 32);
  pl.PWC = pl.PW / 32;
  pl.HC = H / 32;
  pl.NBLK = (pl.NCH + 3) / 4;
  pl.ukimg = ((pl.Rk + 31) / 32) * pl.NQ * 128;
  pl.SS = 2 * H * 128;
  if (pl.PWC > 4 || pl.NQ > 256 || pl.NCY > TK_NCW) return false;
  pl.c_d1 = 0; pl.c_fl = H; pl.c_gram = H + pl.NQ; pl.c_dw = H + 2 * pl.NQ;
  pl.c_dwlo = pl.c_dw + pl.NBLK * H;
  if (pl.c_dwlo + pl.NBLK * H > 512) return false;  // (general observations: a lo block for every block)
  return true;
}

int vjf_tile_create(vjf_handle* h) {
  StepParams& p = h->base;
  TilePlan pl;
  memset(&pl, 0, sizeof(pl));
  h->w1k = nullptr; h->uk = nullptr;
  if (!tile_static_dims(p, pl)) return 0;
  const size_t w1k_bytes = (size_t)pl.NCH * 2 * p.H[0] * 32 * sizeof(float), uk_bytes = 2 * (size_t)pl.ukimg;
  VJF_CUDA_OK(cudaMalloc(&h->w1k, w1k_bytes));
  VJF_CUDA_OK(cudaMemset(h->w1k, 0, w1k_bytes));
  VJF_CUDA_OK(cudaMalloc(&h->uk, uk_bytes));
  VJF_CUDA_OK(cudaMemset(h->uk, 0, uk_bytes));
  VJF_CUDA_OK(cudaFuncSetAttribute(vjf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  return 0;
}

// Returns 1 when the tile pipeline runs this launch (p.tp and the tensor map are filled), 0 when the launch stays on the
// persistent kernel of k_persistent.cu (shapes outside the plan), < 0 on error.
int vjf_tile_plan(vjf_handle* h, StepParams& p, const void* y, int y_dtype, int T, int B, CUtensorMap* map, cudaStream_t stream) {
  static const bool disabled = getenv("VJF_B200_NO_TILE") != nullptr;
  p.tp.on = 0;
  if (disabled || g_tile_mode == 1 || !h->w1k || y_dtype != VJF_Y_F32 || (reinterpret_cast<uintptr_t>(y) & 15) != 0 || h->num_sms < 2 || !encode_fn()) return 0;
  TilePlan pl;
  memset(&pl, 0, sizeof(pl));
  if (!tile_static_dims(p, pl)) return 0;
  const int H = p.H[0], d = p.d, D = p.D;
  static const int force_tbr = getenv("VJF_B200_TILE_ROWS") ? atoi(getenv("VJF_B200_TILE_ROWS")) : 0;
  static const bool no_exact_env = getenv("VJF_B200_NO_EXACT") != nullptr;
  const bool no_exact = no_exact_env || g_tile_mode == 2;
  // spike counts (Poisson likelihood): are they exact in tf32?
  bool exact = false;
  if (p.lik == VJF_LIK_POISSON && !no_exact && vjf_observations_exact(h, y, (size_t)T * B * D, stream, &exact)) return -1;
  pl.CL0 = exact ? D / 32 : 0;
  pl.NBLKLO = pl.NBLK - pl.CL0 / 4;  // lo accumulator blocks, aligned with the blocks CL0 / 4 .. of the hi part
  // tiles of 32 trials; 64 when the lo image is small (exact observations) and every trial CTA gets more than one tile
  pl.TBR = force_tbr ? force_tbr : ((exact && ((B + 31) / 32 > h->max_slots - 1 || g_tile_mode == 3)) ? 64 : 32);
  // one tile per trial CTA (latency regime): the second observation buffer would never be used -- its space goes to the
  // weight ring instead (deeper prefetch of the layer-1 weight chunks)
  const int ntiles = (B + pl.TBR - 1) / pl.TBR;
  const int nbuf_pref = (ntiles <= h->max_slots - 1) ? 1 : 2;
  pl.RS = std::max(1, std::min(TK_NCW / pl.NCY, pl.TBR));
  const int TBR = pl.TBR, chunkB = TBR * 128;
  static const int force_ns = getenv("VJF_B200_TILE_NS") ? atoi(getenv("VJF_B200_TILE_NS")) : 0;
  const int ns_hi = force_ns ? force_ns : std::min(TK_MAXNS, pl.NCH);
  bool fits = false;
  for (int nbuf = nbuf_pref; nbuf >= 1 && !fits; --nbuf)
  for (int ns = ns_hi; ns >= (nbuf == 2 ? 3 : 2) && !fits; --ns) {
  pl.NBUF = nbuf; pl.NS = ns;
  int o = 0;
  pl.o_in[0] = o; o += pl.NCH * chunkB;
  pl.o_in[1] = o; o += (pl.NBUF > 1 ? pl.NCH * chunkB : 0);
  if (pl.NBUF == 1) pl.o_in[1] = pl.o_in[0];
  pl.o_inlo = o; o += (pl.NCH - pl.CL0) * chunkB;
  pl.o_pg = o; o += std::max(2 * pl.PWC, 2 * pl.HC) * chunkB;
  pl.o_ring = o; o += pl.NS * pl.SS;
  pl.o_uk = o; o += 2 * pl.ukimg;
  pl.o_f = o;
  pl.scratch_bytes = pl.o_uk - pl.o_in[0];
  int f = 0;
  auto take = [&](int n) { int at = f; f = (f + n + 3) & ~3; return at; };
  pl.ldhs = H + 1;
  pl.f_hs = take(TBR * pl.ldhs);
  pl.f_dec = take((d + 1) * D);
  pl.f_hm = take(H * d);
  pl.f_hv = take(H * d + d);
  pl.f_cen = take(p.R * p.du);
  pl.f_iw = take(p.R);
  pl.f_ex = take(TBR * std::max(p.E, 1));
  pl.f_eps = take(TBR * 2 * d);
  pl.f_xu = take(TBR * p.du);
  pl.f_xt = take(TBR * d); pl.f_mt = take(TBR * d); pl.f_lt = take(TBR * d); pl.f_pm = take(TBR * d); pl.f_dx = take(TBR * d);
  pl.f_gxt = take(TBR * d); pl.f_gmt = take(TBR * d); pl.f_glt = take(TBR * d); pl.f_plv = take(TBR);
  pl.f_gxp = take(pl.NCY * TBR * d + 4 * TBR);  // g_xt partials per observation chunk, then the partial sums of squares of the quadratic form
  pl.f_red = take(16 * VJF_NSCAL + 16);
  pl.f_misc = take(16);
  pl.f_bar = take(2 * 32);  // mbarriers (8 bytes each; the float area starts 16-byte aligned)
  pl.f_total = f;
  pl.smem_bytes = pl.o_f + f * 4 + 1024;  // + slack for the 1024-byte alignment of the base
  if ((size_t)pl.smem_bytes <= h->smem_limit) { fits = true; break; }
  }
  if (!fits) return 0;
  // scratch of the per-step flush (register accumulators of 15 warps) and workspace of the shared phases B1 / B2
  const int DX = d <= 2 ? 2 : (d == 3 ? 3 : (d == 4 ? 4 : 8));
  const int flush_floats = TK_NCW * (DX + 1) * 32 + TK_NCW * 2 * pl.HC * DX * 32 + 16 * VJF_NSCAL + 16 + 16;
  pl.flush_floats = (flush_floats + 31) & ~31;
  if (pl.flush_floats * 4 > pl.scratch_bytes) return 0;
  pl.flush_stage = ((pl.flush_floats + 128 * (std::max(pl.NQ, H) + 4)) * 4 <= pl.scratch_bytes) ? 1 : 0;  // staging block of the accumulator flush
  size_t b2 = 2 * 16 * 17 + p.R + 4 + 3 * ((size_t)p.d * p.R + 4) + 16 + 2 * VJF_NWARP + 8 + (size_t)p.R * (p.R | 1) + 64;
  if (p.R > 128) return 0;
  if (p.rls64) b2 = std::max(b2, vjf_rls64_floats(p));
  const int s_b1 = (int)((b2 + 3) & ~(size_t)3);
  if ((size_t)(s_b1 + 2048 + 8) * 4 > (size_t)(pl.o_uk - pl.o_pg)) return 0;
  // observations as a 3-D tensor [T][B][D]: boxes of {32 columns, TBR trials, 1 step}, 16-byte pieces swizzled within
  // 128-byte rows (= the K-major SWIZZLE_128B operand image); rows / columns past the end are zero-filled
  const cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)B, (cuuint64_t)T};
  const cuuint64_t gstr[2] = {(cuuint64_t)D * 4, (cuuint64_t)B * D * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)TBR, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(y), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 0;
  pl.on = 1;
  p.tp = pl;
  p.s_b1 = s_b1;
  p.w1k = h->w1k; p.uk = h->uk;
  p.B = B;
  p.TB = TBR;
  p.ntiles = (B + TBR - 1) / TBR;
  p.nslots = std::min(p.ntiles, h->max_slots - 1) + 1;
  p.overlap = 1; p.use_tma = 0; p.use_umma = 0;
  return 1;
}

int vjf_tile_launch(vjf_handle* h, StepParams& p, const CUtensorMap& map, cudaStream_t s) {
  VJF_CUDA_OK(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), s));
  VJF_CUDA_OK(cudaMemsetAsync(p.ctrl, 0, 8 * sizeof(unsigned), s));
  CUtensorMap m = map;
  void* args[] = {(void*)&p, (void*)&m};
  VJF_CUDA_OK(cudaLaunchCooperativeKernel((void*)vjf_tile_kernel, dim3(p.nslots), dim3(VJF_NT), args, (size_t)p.tp.smem_bytes, s));
  ++g_vjf_launches;
  return 0;
}
