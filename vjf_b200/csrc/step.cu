// Kernels + C ABI of the VJF filter/learning step.  See include/vjf_b200.h for the contract.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

// Process-wide registry of opened CUDA IPC mappings.  cudaIpcOpenMemHandle refuses to map an allocation twice ("resource
// already mapped"): two handles of one process may see the same peer allocation -- a second model connected while the first
// is still alive, or a peer that freed its exchange buffer and got the same memory back for the next model.  Mappings are
// shared and reference-counted instead.
namespace {
struct IpcEntry { cudaIpcMemHandle_t h; void* ptr; int refs; int device; };
std::vector<IpcEntry> g_ipc;
std::mutex g_ipc_mu;
int ipc_open(const cudaIpcMemHandle_t& hd, void** out) {
  int device = 0;
  VJF_CUDA_OK(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  for (auto& e : g_ipc)
    if (e.device == device && memcmp(&e.h, &hd, sizeof(hd)) == 0) { ++e.refs; *out = e.ptr; return 0; }
  void* ptr = nullptr;
  VJF_CUDA_OK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  g_ipc.push_back({hd, ptr, 1, device});
  *out = ptr;
  return 0;
}
void ipc_close(void* ptr) {
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  for (size_t i = 0; i < g_ipc.size(); ++i)
    if (g_ipc[i].ptr == ptr) {
      if (--g_ipc[i].refs == 0) { cudaIpcCloseMemHandle(ptr); g_ipc.erase(g_ipc.begin() + i); }
      return;
    }
}
}  // namespace

// ------------------------------------------------------------------------------------------
// error handling / accounting
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
long long g_vjf_launches = 0;

void vjf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* vjf_last_error(void) { return g_err; }
extern "C" int vjf_version(void) { return 100; }
extern "C" int64_t vjf_launch_count(void) { return g_vjf_launches; }
static int g_vjf_last_kind = 0;
extern "C" int32_t vjf_last_launch_kind(void) { return g_vjf_last_kind; }

// ------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------
static inline int64_t up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

extern "C" int vjf_get_layout(const vjf_config* c, vjf_layout* o) {
  if (!c || !o) { vjf_set_error("null argument"); return -1; }
  if (c->ydim < 1 || c->xdim < 1 || c->xdim > VJF_MAX_XDIM || c->udim < 0 || c->n_rbf < 1 || c->n_layers < 1 ||
      c->n_layers > VJF_MAX_LAYERS) {
    vjf_set_error("unsupported configuration: need 1<=xdim<=%d, 1<=len(hidden_sizes)<=%d, ydim>=1, n_rbf>=1",
                  VJF_MAX_XDIM, VJF_MAX_LAYERS);
    return -1;
  }
  for (int l = 0; l < c->n_layers; ++l)
    if (c->hidden[l] < 1) { vjf_set_error("hidden size must be positive"); return -1; }
  const int64_t D = c->ydim, d = c->xdim, u = c->udim, R = c->n_rbf;
  memset(o, 0, sizeof(*o));
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t at = off; off = up(off + n, 32); return at; };
  o->lik_logvar = take(1);
  o->dec_w = take(d * D);
  o->dec_b = take(D);
  int64_t in = D + u + 2 * d;
  for (int l = 0; l < c->n_layers; ++l) {
    o->mlp_w[l] = take(in * c->hidden[l]);
    o->mlp_b[l] = take(c->hidden[l]);
    in = c->hidden[l];
  }
  o->head_m_w = take(in * d);
  o->head_v_w = take(in * d);
  o->head_v_b = take(d);
  o->n_train = off;
  o->prior_mean = take(d);
  o->prior_logvar = take(d);
  o->tr_logvar = take(1);
  o->centroid = take(R * (d + u));
  o->logwidth = take(R);
  o->w_mean = take(R * d);
  o->w_chol = take(R * R);
  o->w_precision = take(R * R);
  o->w_pchol = take(R * R);
  o->lik_n = take(1);
  o->tr_n = take(1);
  o->total = off;
  if (o->total >= (int64_t)1 << 31) { vjf_set_error("state buffer too large"); return -1; }
  return 0;
}

static void lay_to_int(const vjf_layout& a, Lay& b) {
  b.lik_logvar = (int)a.lik_logvar; b.dec_w = (int)a.dec_w; b.dec_b = (int)a.dec_b;
  for (int l = 0; l < VJF_MAX_LAYERS; ++l) { b.mlp_w[l] = (int)a.mlp_w[l]; b.mlp_b[l] = (int)a.mlp_b[l]; }
  b.head_m_w = (int)a.head_m_w; b.head_v_w = (int)a.head_v_w; b.head_v_b = (int)a.head_v_b; b.n_train = (int)a.n_train;
  b.prior_mean = (int)a.prior_mean; b.prior_logvar = (int)a.prior_logvar; b.tr_logvar = (int)a.tr_logvar;
  b.centroid = (int)a.centroid; b.logwidth = (int)a.logwidth; b.w_mean = (int)a.w_mean; b.w_chol = (int)a.w_chol;
  b.w_precision = (int)a.w_precision; b.w_pchol = (int)a.w_pchol; b.lik_n = (int)a.lik_n; b.tr_n = (int)a.tr_n;
  b.total = (int)a.total;
}

// smallest x >= lo with x % 32 == tgt (tgt is a multiple of 4): bank-conflict-free fragment strides
static int pad_mod32(int lo, int tgt) {
  int x = (lo / 32) * 32 + tgt;
  while (x < lo) x += 32;
  return x;
}

// shared-memory plan for a tile of `tb` trials (rows padded to a multiple of 16 for the MMA tiles);
// returns the number of floats (maximum over phase A, B1 and B2)
static size_t plan_smem(StepParams& p, int tb, bool u_in_smem, bool dec_in_smem, bool w1_in_smem, bool in_split) {
  const int rows = (tb + 15) & ~15;
  size_t off = 0;
  auto take = [&](size_t n) { size_t at = off; off = (off + n + 3) & ~(size_t)3; return (int)at; };
  p.s_in = take((size_t)rows * p.K1p);
  p.in_split = in_split ? 1 : 0;
  p.s_inl = in_split ? take((size_t)rows * p.K1p) : p.s_in;
  p.s_g = take((size_t)((tb + 3) & ~3) * p.Dp);  // only real trials are read back
  const bool ext = p.ext != 0;  // large n_rbf (bigr.cu): nothing of size n_rbf lives in the tile kernels' shared memory
  p.s_phi = take(ext ? 0 : (size_t)rows * p.Rp);
  p.s_phil = take(ext ? 0 : (size_t)rows * p.Rp);
  for (int l = 0; l < p.L; ++l) p.s_act[l] = take((size_t)rows * p.Hp[l]);
  p.s_gpa = take((size_t)rows * p.Gp);
  p.s_gpal = take((size_t)rows * p.Gp);
  // the second (ping-pong) pair is only needed to back-propagate through more than one hidden layer
  p.s_gpb = (p.L > 1) ? take((size_t)rows * p.Gp) : p.s_gpa;
  p.s_gpbl = (p.L > 1) ? take((size_t)rows * p.Gp) : p.s_gpal;
  p.s_eps = take((size_t)rows * 2 * p.d);
  p.s_xu = take((size_t)rows * p.du);
  p.s_xt = take((size_t)rows * p.d); p.s_mt = take((size_t)rows * p.d); p.s_lt = take((size_t)rows * p.d);
  p.s_pm = take((size_t)rows * p.d); p.s_dx = take((size_t)rows * p.d); p.s_gxt = take((size_t)rows * p.d);
  p.s_gmt = take((size_t)rows * p.d); p.s_glt = take((size_t)rows * p.d); p.s_plv = take((size_t)rows);
  p.s_qp = take(ext ? 32 : (size_t)((p.R + 7) / 8) * 32 + 32);
  p.U_in_smem = u_in_smem ? 1 : 0;
  p.s_U = take(u_in_smem ? (size_t)((p.R + 7) & ~7) * p.ldu : 0);
  p.dec_in_smem = dec_in_smem ? 1 : 0;
  p.s_dec = take(dec_in_smem ? (size_t)(p.d + 1) * p.D : 0);
  p.W1_in_smem = w1_in_smem ? 1 : 0;
  p.s_W1 = take(w1_in_smem ? (size_t)p.K1 * p.ldw1 : 0);
  p.s_hm = take((size_t)p.H[p.L - 1] * p.d);
  p.s_hv = take((size_t)p.H[p.L - 1] * p.d + p.d);
  p.s_flag = take(12);  // [0] flag, [1] TMEM base, [2..7] three mbarriers, [8] non-finite partial seen, [9] barrier broadcast
  p.s_scf = take(VJF_NSCAL);
  p.s_b1 = take(2048 + 8);
  p.s_W = take(ext ? 0 : (size_t)p.R * p.d);
  p.s_c = take(ext ? 0 : (size_t)p.R * p.du);
  p.s_iw = take(ext ? 0 : (size_t)p.R);
  p.s_red = take((size_t)VJF_NWARP * VJF_NSCAL + 64);
  const size_t a = off;
  // phase B2: register path needs ~2(2R+d) + R + 2dR floats; the shared-memory fallback (R > 128)
  // [(2R+d)][ldm] + pivots; phase B1: 512 floats
  size_t b2 = 2 * 16 * 17 + p.R + 4 + 3 * ((size_t)p.d * p.R + 4) + 16 + 2 * VJF_NWARP + 8 + (size_t)p.R * (p.R | 1) + 64;
  if (p.R > 128) b2 = (size_t)(2 * p.R + p.d) * p.ldm + ((p.R + 3) & ~3) + 4 + 2 * VJF_NWARP + 8;
  if (p.rls64) b2 = std::max(b2, vjf_rls64_floats(p));
  if (ext) b2 = 1024;
  p.s_total = (int)std::max(std::max(a, b2), (size_t)1024);
  return (size_t)p.s_total;
}

// persistent != 0: plan for the cooperative kernel, where CTA 0 is the dedicated RLS CTA whenever every trial CTA
// gets exactly one tile (the overlapped schedule)
static int plan_tiles(vjf_handle* h, StepParams& p, int B, int max_slots, int persistent = 0) {
  if (B < 1 || B > h->cfg.max_trials) { vjf_set_error("trials B=%d outside [1, max_trials=%d]", B, h->cfg.max_trials); return -1; }
  p.overlap = 0;
  if (persistent && max_slots > 1 && (B + VJF_TB_MAX - 1) / VJF_TB_MAX <= max_slots - 1) { p.overlap = 1; max_slots -= 1; }
  const int want = std::min(std::max((int)up((B + max_slots - 1) / max_slots, 4), 4), VJF_TB_MAX);
  const size_t limit = h->smem_limit;
  bool u_smem = !p.ext && (size_t)((p.R + 7) & ~7) * p.ldu * 4 <= 96 * 1024;
  bool dec_smem = (size_t)(p.d + 1) * p.D * 4 <= 32 * 1024;
  bool w1_smem = (size_t)p.K1 * p.ldw1 * 4 <= 72 * 1024;
  bool in_split = true;
  int tb = want;
  for (;;) {
    if (plan_smem(p, tb, u_smem, dec_smem, w1_smem, in_split) * 4 <= limit) break;
    if (in_split && (size_t)16 * p.K1p * 4 > 48 * 1024) { in_split = false; continue; }  // wide observations: split on the fly
    if (w1_smem && tb <= want - 8) { w1_smem = false; tb = want; continue; }
    if (tb > 4) { tb -= 4; continue; }
    if (w1_smem) { w1_smem = false; tb = want; continue; }
    if (dec_smem) { dec_smem = false; tb = want; continue; }
    if (u_smem) { u_smem = false; tb = want; continue; }
    vjf_set_error("configuration does not fit in %zu bytes of shared memory (ydim=%d n_rbf=%d)", limit, p.D, p.R);
    return -1;
  }
  p.B = B;
  p.TB = tb;
  p.ntiles = (B + tb - 1) / tb;
  if (p.overlap && p.ntiles > max_slots) p.overlap = 0;  // the tile had to shrink to fit shared memory
  // layer-1 weight gradient on tcgen05 (umma.cuh): one tile per CTA (the operand pair aliases the staged W1, which
  // must not be needed by a later tile), one hidden layer of at most 64 units (M = 64), N = roundup(K1, 8) <= 256
  {
    const int rows = (tb + 15) & ~15, nk = (p.K1 + 7) & ~7;
    static const bool no_umma = getenv("VJF_B200_NO_UMMA") != nullptr;
    p.umma_nk = nk;
    static const bool no_tma = getenv("VJF_B200_NO_TMA") != nullptr;
    p.use_tma = (!no_tma && p.overlap && (p.D & 3) == 0 && (p.H[0] & 3) == 0 && ((p.H[p.L - 1] * p.d) & 3) == 0) ? 1 : 0;
    p.use_umma = (!no_umma && p.overlap && p.L == 1 && p.H[0] <= 64 && p.Gp >= 64 && p.W1_in_smem && nk <= 256 &&
                  (size_t)p.K1 * p.ldw1 >= (size_t)2 * nk * rows && (size_t)p.K1 * p.ldw1 >= 4 * 32 * 64 /* epilogue staging */) ? 1 : 0;
  }
  p.nslots = p.overlap ? p.ntiles + 1 : std::min(p.ntiles, max_slots + (persistent && max_slots < h->max_slots ? 1 : 0));
  return 0;
}

// ------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------
extern "C" int vjf_create(const vjf_config* cfg, float* state, vjf_handle** out) {
  if (!cfg || !state || !out) { vjf_set_error("null argument"); return -1; }
  vjf_layout lay;
  if (vjf_get_layout(cfg, &lay)) return -1;
  if (cfg->max_trials < 1) { vjf_set_error("max_trials must be >= 1"); return -1; }
  if (cfg->likelihood != VJF_LIK_POISSON && cfg->likelihood != VJF_LIK_GAUSSIAN) { vjf_set_error("unknown likelihood id"); return -1; }
  vjf_handle* h = (vjf_handle*)calloc(1, sizeof(vjf_handle));
  h->cfg = *cfg;
  h->lay64 = lay;
  h->state = state;
  VJF_CUDA_OK(cudaGetDevice(&h->device));
  cudaDeviceProp prop;
  VJF_CUDA_OK(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) {
    vjf_set_error("vjf_b200 is built for sm_100a (B200); found compute capability %d.%d -- there is no fallback path", prop.major, prop.minor);
    free(h);
    return -3;
  }
  h->num_sms = prop.multiProcessorCount;
  h->smem_limit = prop.sharedMemPerBlockOptin;
  VJF_CUDA_OK(cudaFuncSetAttribute(vjf_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  VJF_CUDA_OK(cudaFuncSetAttribute(vjf_phase_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  VJF_CUDA_OK(cudaFuncSetAttribute(vjf_phase_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  VJF_CUDA_OK(cudaFuncSetAttribute(vjf_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_limit));
  h->max_slots = h->num_sms;  // one persistent CTA per SM

  StepParams& p = h->base;
  memset(&p, 0, sizeof(p));
  p.D = cfg->ydim; p.d = cfg->xdim; p.u = cfg->udim; p.R = cfg->n_rbf; p.L = cfg->n_layers;
  p.K1 = p.D + p.u + 2 * p.d; p.E = p.u + 2 * p.d; p.du = p.d + p.u;
  // strides chosen so that the mma fragment loads are bank-conflict free (see mma.cuh)
  p.K1p = pad_mod32((int)up(p.K1, 16), 4);
  p.Dp = (int)up(p.D, 4);
  p.Rp = pad_mod32((int)up(p.R, 16), 4);
  p.ldu = pad_mod32((int)up(p.R, 8), 8);
  p.Hpmax = 4;
  int hmax = 1;
  for (int l = 0; l < p.L; ++l) {
    p.H[l] = cfg->hidden[l];
    p.Hp[l] = pad_mod32((int)up(p.H[l], 16), 4);
    p.Hpmax = std::max(p.Hpmax, p.Hp[l]);
    hmax = std::max(hmax, p.H[l]);
  }
  p.Gp = pad_mod32((int)up(hmax, 8), 8);
  p.ldw1 = pad_mod32((int)up(p.H[0], 4), 8);
  p.lik = cfg->likelihood;
  p.world = 1;
  lay_to_int(lay, p.lay);
  p.G = p.lay.n_train;
  p.pa = p.G;
  p.ext = p.R > VJF_BIGR_MIN ? 1 : 0;  // large n_rbf: the RLS statistics are GEMM outputs (bigr.cu), not slot sums
  p.pb = p.ext ? p.pa : (int)up(p.pa + (int64_t)p.R * p.R, 4);
  p.ps = p.ext ? p.pb : (int)up(p.pb + (int64_t)p.R * p.d, 4);
  p.PS = p.ps + VJF_NSCAL;
  p.ldm = (p.R + 1) | 1;
  p.state = state;

  const size_t part_bytes = (size_t)h->max_slots * p.PS * sizeof(float);
  VJF_CUDA_OK(cudaMalloc(&h->partials, part_bytes));
  VJF_CUDA_OK(cudaMemset(h->partials, 0, part_bytes));
  // reduced vector, followed by the row-padded mirror of the recognition layer-1 weight (TMA source, 128-byte aligned)
  const size_t red_floats = (size_t)up(p.PS, 32) + (size_t)up((int64_t)p.K1 * p.ldw1, 32) + (p.ext ? 32 : (size_t)((p.R + 7) & ~7) * p.ldu);
  VJF_CUDA_OK(cudaMalloc(&h->reduced, red_floats * sizeof(float)));
  VJF_CUDA_OK(cudaMemset(h->reduced, 0, red_floats * sizeof(float)));
  p.w1_mirror = h->reduced + up(p.PS, 32);
  p.u_mirror = p.w1_mirror + up((int64_t)p.K1 * p.ldw1, 32);  // [roundup(R, 8)][ldu] row-padded copy of w_chol
  VJF_CUDA_OK(cudaMalloc(&h->sync_words, 64 * sizeof(unsigned)));
  VJF_CUDA_OK(cudaMemset(h->sync_words, 0, 64 * sizeof(unsigned)));
  p.partials = h->partials; p.reduced = h->reduced;
  p.barrier = h->sync_words; p.status = h->sync_words + 16; p.ctrl = h->sync_words + 32;
  if (!p.ext) {
    VJF_CUDA_OK(cudaMalloc(&h->P64, (size_t)p.R * p.R * sizeof(double)));
    VJF_CUDA_OK(cudaMemset(h->P64, 0, (size_t)p.R * p.R * sizeof(double)));
    p.P64 = h->P64;
  }
  if (vjf_tile_create(h)) return -2;
  if (p.ext && vjf_bigr_create(h)) return -2;
  if (vjf_wide_create(h)) return -2;
  *out = h;
  return 0;
}

extern "C" int vjf_destroy(vjf_handle* h) {
  if (!h) return 0;
  cudaFree(h->partials); cudaFree(h->reduced); cudaFree(h->sync_words); cudaFree(h->w1k); cudaFree(h->uk);
  for (int r = 0; r < h->comm_world; ++r) if (r != h->comm_rank && h->peer[r]) ipc_close(h->peer[r]);
  cudaFree(h->xbuf);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->stage_y[i]); cudaFree(h->stage_u[i]); cudaFree(h->stage_eps[i]);
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
  }
  cudaFree(h->stage_mu); cudaFree(h->stage_lv); cudaFree(h->stage_loss); cudaFree(h->fc_w); cudaFree(h->wk_ws); cudaFree(h->P64); vjf_bigr_destroy(h); vjf_wide_destroy(h);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->compute_stream) cudaStreamDestroy(h->compute_stream);
  free(h);
  return 0;
}

__global__ void vjf_init_state_kernel(float* st, Lay lay, int R, int d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R * R) {
    const float v = (i / R == i % R) ? 1.0f : 0.0f;  // vjf/module.py:52-54
    st[lay.w_chol + i] = v; st[lay.w_precision + i] = v; st[lay.w_pchol + i] = v;
  }
  if (i < R * d) st[lay.w_mean + i] = 0.f;            // module.py:46
  if (i < R) st[lay.logwidth + i] = 0.f;              // module.py:21
  if (i < d) { st[lay.prior_mean + i] = 0.f; st[lay.prior_logvar + i] = 0.f; }  // model.py:66-67
  if (i == 0) {
    st[lay.lik_logvar] = logf(0.1f);                  // likelihood.py:16
    st[lay.tr_logvar] = 0.f;                          // model.py:331
    st[lay.lik_n] = 0.f; st[lay.tr_n] = 0.f;
  }
}

extern "C" int vjf_init_state(vjf_handle* h, void* stream) {
  if (!h) { vjf_set_error("null handle"); return -1; }
  const int R = h->base.R, n = R * R;
  vjf_init_state_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->state, h->base.lay, R, h->base.d);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// step / run
// ------------------------------------------------------------------------------------------
static int launch_persistent(vjf_handle* h, StepParams& p, cudaStream_t s) {
  VJF_CUDA_OK(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), s));
  VJF_CUDA_OK(cudaMemsetAsync(p.ctrl, 0, 8 * sizeof(unsigned), s));
  void* args[] = {(void*)&p};
  VJF_CUDA_OK(cudaLaunchCooperativeKernel((void*)vjf_persistent_kernel, dim3(p.nslots), dim3(VJF_NT), args,
                                          (size_t)p.s_total * sizeof(float), s));
  ++g_vjf_launches;
  return 0;
}

static int check_ptrs(const vjf_handle* h, const void* y, const float* u, const float* qm, const float* ql, uint32_t flags,
                      const float* mu, const float* lv) {
  if (!h) { vjf_set_error("null handle"); return -1; }
  if (!y || !mu || !lv) { vjf_set_error("null y / output pointer"); return -1; }
  if (h->cfg.udim > 0 && !u) { vjf_set_error("udim=%d but u is NULL", h->cfg.udim); return -1; }
  if (!(flags & VJF_FLAG_PRIOR_Q0) && (!qm || !ql)) { vjf_set_error("previous posterior is NULL and VJF_FLAG_PRIOR_Q0 is not set"); return -1; }
  return 0;
}

// the throughput tile pipeline when the shapes are in its plan, else the persistent kernel of k_persistent.cu
static int launch_time_loop(vjf_handle* h, StepParams& p, int T, int B, cudaStream_t s) {
  if (p.ext) {
    if (p.world > 1) { vjf_set_error("n_rbf > %d: the sharded run is not implemented for the large-n_rbf path", VJF_BIGR_MIN); return -1; }
    g_vjf_last_kind = 2;
    return vjf_bigr_time_loop(h, p, T, B, s);
  }
  CUtensorMap map;
  const int use_tile = vjf_tile_plan(h, p, p.y, p.y_dtype, T, B, &map, s);
  if (use_tile < 0) return -2;
  g_vjf_last_kind = use_tile ? 1 : 0;
  if (use_tile) return vjf_tile_launch(h, p, map, s);
  if (vjf_wide_applies(h, p, B)) { g_vjf_last_kind = 3; return vjf_wide_time_loop(h, p, T, B, s); }
  if (plan_tiles(h, p, B, h->max_slots, 1)) return -1;
  return launch_persistent(h, p, s);
}

// development aid, debug builds only (python -m vjf_b200.build --debug): per-phase timestamps for the next launches
static long long* g_dbg_ptr = nullptr;
static int g_dbg_cta = 1;
#ifdef VJF_DEBUG_STAMPS
extern "C" void vjf_debug_set_stamps(long long* dev_ptr) { g_dbg_ptr = dev_ptr; }
extern "C" void vjf_debug_set_cta(int cta) { g_dbg_cta = cta; }
#endif

extern "C" int vjf_run(vjf_handle* h, int32_t T, int32_t B, const void* y, int32_t y_dtype, const float* u,
                       const float* q0_mean, const float* q0_logvar, const float* eps, uint64_t seed, uint64_t step0,
                       uint32_t flags, float lr, float* mu, float* logvar, float* losses, void* stream) {
  if (check_ptrs(h, y, u, q0_mean, q0_logvar, flags, mu, logvar)) return -1;
  if (T < 1) { vjf_set_error("T must be >= 1"); return -1; }
  if (y_dtype != VJF_Y_F32 && y_dtype != VJF_Y_U8) { vjf_set_error("unknown y dtype"); return -1; }
  if (B < 1 || B > h->cfg.max_trials) { vjf_set_error("trials B=%d outside [1, max_trials=%d]", B, h->cfg.max_trials); return -1; }
  StepParams p = h->base;
  p.Bglobal = B;
  p.y = y; p.y_dtype = y_dtype; p.u_in = u; p.q0m = q0_mean; p.q0l = q0_logvar; p.eps = eps;
  p.mu = mu; p.logvar = logvar; p.losses = losses;
  p.seed = seed; p.step0 = step0; p.trial_offset = 0; p.flags = flags; p.lr = lr; p.T = T;
  p.dbg = g_dbg_ptr;
  p.dbg_cta = g_dbg_cta;
  return launch_time_loop(h, p, T, B, (cudaStream_t)stream);
}


// ---- sharded run: exchange buffers over CUDA IPC ----
static size_t xbuf_floats(const StepParams& p) { return (size_t)VJF_MAX_RANKS * 2 * ((p.PS + 127) & ~127); }

extern "C" int vjf_comm_local_handle(vjf_handle* h, void* out_handle64) {
  if (!h || !out_handle64) { vjf_set_error("null argument"); return -1; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!h->xbuf) {
    const size_t PSx = (h->base.PS + 127) & ~127;
    h->xbuf_bytes = xbuf_floats(h->base) * sizeof(float) + (size_t)VJF_MAX_RANKS * (PSx / 128) * sizeof(unsigned) + 256;
    VJF_CUDA_OK(cudaMalloc(&h->xbuf, h->xbuf_bytes));
    VJF_CUDA_OK(cudaMemset(h->xbuf, 0, h->xbuf_bytes));
    VJF_CUDA_OK(cudaDeviceSynchronize());
  }
  else {
    // a second connect: the flags still hold the epochs of the previous session, which a restarted epoch counter would accept
    VJF_CUDA_OK(cudaMemset(h->xbuf, 0, h->xbuf_bytes));
    VJF_CUDA_OK(cudaDeviceSynchronize());
  }
  VJF_CUDA_OK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(out_handle64), h->xbuf));
  return 0;
}

extern "C" int vjf_comm_connect(vjf_handle* h, int32_t rank, int32_t world, const void* handles) {
  if (!h || !handles || world < 1 || world > VJF_MAX_RANKS || rank < 0 || rank >= world) { vjf_set_error("bad comm arguments (world <= %d)", VJF_MAX_RANKS); return -1; }
  if (!h->xbuf) { vjf_set_error("call vjf_comm_local_handle first"); return -1; }
  const cudaIpcMemHandle_t* hs = reinterpret_cast<const cudaIpcMemHandle_t*>(handles);
  for (int r = 0; r < h->comm_world; ++r) if (r != h->comm_rank && h->peer[r]) { ipc_close(h->peer[r]); h->peer[r] = nullptr; }
  h->comm_world = 0;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { h->peer[r] = h->xbuf; continue; }
    void* ptr = nullptr;
    if (ipc_open(hs[r], &ptr)) return -2;
    h->peer[r] = reinterpret_cast<float*>(ptr);
  }
  h->comm_rank = rank; h->comm_world = world; h->comm_epoch = 0;
  return 0;
}

extern "C" int vjf_run_sharded(vjf_handle* h, int32_t T, int32_t B_local, int32_t B_global, uint64_t trial_offset, const void* y,
                               int32_t y_dtype, const float* u, const float* q0_mean, const float* q0_logvar, const float* eps,
                               uint64_t seed, uint64_t step0, uint32_t flags, float lr, float* mu, float* logvar, float* losses,
                               void* stream) {
  if (check_ptrs(h, y, u, q0_mean, q0_logvar, flags, mu, logvar)) return -1;
  if (h->comm_world < 1) { vjf_set_error("vjf_comm_connect has not been called"); return -1; }
  if (T < 1 || B_global < B_local) { vjf_set_error("bad T / batch sizes"); return -1; }
  if (B_local < 1 || B_local > h->cfg.max_trials) { vjf_set_error("trials B=%d outside [1, max_trials=%d]", B_local, h->cfg.max_trials); return -1; }
  StepParams p = h->base;
  p.Bglobal = B_global;
  p.y = y; p.y_dtype = y_dtype; p.u_in = u; p.q0m = q0_mean; p.q0l = q0_logvar; p.eps = eps;
  p.mu = mu; p.logvar = logvar; p.losses = losses;
  p.seed = seed; p.step0 = step0; p.trial_offset = trial_offset; p.flags = flags; p.lr = lr; p.T = T;
  p.world = h->comm_world; p.rank = h->comm_rank; p.PSx = (p.PS + 127) & ~127; p.epoch0 = h->comm_epoch;
  for (int r = 0; r < p.world; ++r) p.peer[r] = h->peer[r];
  h->comm_epoch += (unsigned)T;
  return launch_time_loop(h, p, T, B_local, (cudaStream_t)stream);
}

extern "C" int vjf_step(vjf_handle* h, int32_t B, const float* y, const float* u, const float* q_mean, const float* q_logvar,
                        const float* eps, uint64_t seed, uint64_t step_index, uint32_t flags, float lr, float* out_mean,
                        float* out_logvar, float* out_loss, void* stream) {
  return vjf_run(h, 1, B, y, VJF_Y_F32, u, q_mean, q_logvar, eps, seed, step_index, flags, lr, out_mean, out_logvar,
                 out_loss, stream);
}

int vjf_plan_tiles_public(vjf_handle* h, StepParams& p, int B) { return plan_tiles(h, p, B, h->max_slots); }

int vjf_internal_reduce(const StepParams& p, cudaStream_t s) {
  vjf_reduce_kernel<<<(p.PS - p.red_begin + 127) / 128, VJF_NT, (size_t)(p.s_b1 + 2048 + 8) * sizeof(float), s>>>(p);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int vjf_set_rls_precision(vjf_handle* h, int32_t bits) {
  if (!h || (bits != 32 && bits != 64)) { vjf_set_error("vjf_set_rls_precision: bits must be 32 or 64"); return -1; }
  if (h->base.ext && bits == 64) { vjf_set_error("n_rbf > %d: the large-n_rbf path factorises in fp32", VJF_BIGR_MIN); return -1; }
  StepParams probe = h->base;
  probe.rls64 = 1;
  if (bits == 64 && (vjf_rls64_floats(probe) + 2048 + 64) * sizeof(float) > h->smem_limit) {
    vjf_set_error("n_rbf=%d: the double-precision factorisation workspace does not fit in shared memory", h->base.R);
    return -1;
  }
  h->base.rls64 = bits == 64;
  return 0;
}

extern "C" int64_t vjf_reduce_size(vjf_handle* h) { return h ? h->base.PS : -1; }
extern "C" float* vjf_reduce_buffer(vjf_handle* h) { return h ? h->reduced : nullptr; }

extern "C" int vjf_step_phase_a(vjf_handle* h, int32_t B_local, int32_t B_global, const float* y, int32_t y_dtype,
                                const float* u, const float* q_mean, const float* q_logvar, const float* eps, uint64_t seed,
                                uint64_t step_index, uint64_t trial_offset, uint32_t flags, float* out_mean, float* out_logvar,
                                void* stream) {
  if (check_ptrs(h, y, u, q_mean, q_logvar, flags, out_mean, out_logvar)) return -1;
  StepParams p = h->base;
  if (plan_tiles(h, p, B_local, h->max_slots)) return -1;
  p.Bglobal = B_global;
  p.y = y; p.y_dtype = y_dtype; p.u_in = u; p.q0m = q_mean; p.q0l = q_logvar; p.eps = eps;
  p.mu = out_mean; p.logvar = out_logvar; p.losses = nullptr;
  p.seed = seed; p.step0 = step_index; p.trial_offset = trial_offset; p.flags = flags; p.lr = 0.f; p.T = 1;
  cudaStream_t s = (cudaStream_t)stream;
  vjf_phase_a_kernel<<<p.nslots, VJF_NT, (size_t)p.s_total * sizeof(float), s>>>(p);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  if (vjf_internal_reduce(p, s)) return -2;
  return 0;
}

extern "C" int vjf_step_phase_b(vjf_handle* h, int32_t B_global, uint32_t flags, float lr, float* out_loss, void* stream) {
  if (!h) { vjf_set_error("null handle"); return -1; }
  StepParams p = h->base;
  if (plan_tiles(h, p, 1, h->max_slots)) return -1;  // only the B2 workspace matters here
  p.Bglobal = B_global; p.B = B_global;
  p.flags = flags; p.lr = lr; p.T = 1; p.losses = out_loss;
  const int grid = std::max(1, std::min(h->num_sms, (p.lay.n_train + VJF_NT - 1) / VJF_NT));
  vjf_phase_b_kernel<<<grid, VJF_NT, (size_t)p.s_total * sizeof(float), (cudaStream_t)stream>>>(p);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int vjf_get_status(vjf_handle* h, void* stream, uint32_t* out, int32_t clear) {
  if (!h || !out) { vjf_set_error("null argument"); return -1; }
  cudaStream_t s = (cudaStream_t)stream;
  VJF_CUDA_OK(cudaMemcpyAsync(out, h->base.status, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (clear) VJF_CUDA_OK(cudaMemsetAsync(h->base.status, 0, sizeof(uint32_t), s));
  VJF_CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

// ------------------------------------------------------------------------------------------
// Philox tape (lets a caller reproduce, outside the kernel, the very numbers it draws)
// ------------------------------------------------------------------------------------------
__global__ void vjf_philox_kernel(unsigned long long seed, unsigned long long step, unsigned long long off, int B, int d,
                                  float* out) {
  const int nblk = (d + 3) >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * nblk) return;
  const int b = i / (2 * nblk), r = i - b * 2 * nblk, which = r / nblk, blk = r - which * nblk;
  float z[4];
  philox_normal4(seed, step, off + b, which, blk, z);
  for (int k = 0; k < 4; ++k)
    if (blk * 4 + k < d) out[((size_t)which * B + b) * d + blk * 4 + k] = z[k];
}

extern "C" int vjf_philox_normal(uint64_t seed, uint64_t step_index, uint64_t trial_offset, int32_t B, int32_t xdim,
                                 float* eps_out, void* stream) {
  if (!eps_out || B < 1 || xdim < 1) { vjf_set_error("bad argument"); return -1; }
  const int n = B * 2 * ((xdim + 3) / 4);
  vjf_philox_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(seed, step_index, trial_offset, B, xdim, eps_out);
  ++g_vjf_launches;
  VJF_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// end-to-end entry with host buffers: chunked, double-buffered H2D overlapped with compute
// ------------------------------------------------------------------------------------------
static int ensure(void** p, size_t* have, size_t need) {
  if (*have >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  VJF_CUDA_OK(cudaMalloc(p, need));
  *have = need;
  return 0;
}

// sharded != 0: the chunks run through vjf_run_sharded (trials of this rank; B_global / trial_offset as there)
static int run_host_impl(vjf_handle* h, int sharded, int32_t T, int32_t B, int32_t B_global, uint64_t trial_offset, const void* y_host,
                         int32_t y_dtype, const float* u_host, const float* eps_host, uint64_t seed, uint64_t step0, uint32_t flags, float lr,
                         float* mu_host, float* logvar_host, float* losses_host, int32_t chunk_steps) {
  if (!h || !y_host || !mu_host || !logvar_host) { vjf_set_error("null argument"); return -1; }
  if (h->cfg.udim > 0 && !u_host) { vjf_set_error("udim=%d but u is NULL", h->cfg.udim); return -1; }
  if (!(flags & VJF_FLAG_PRIOR_Q0)) { vjf_set_error("vjf_run_host starts from the prior: set VJF_FLAG_PRIOR_Q0"); return -1; }
  if (T < 1 || chunk_steps < 1) { vjf_set_error("T and chunk_steps must be >= 1"); return -1; }
  const int D = h->cfg.ydim, d = h->cfg.xdim, u = h->cfg.udim;
  const size_t ysz = (y_dtype == VJF_Y_U8) ? 1 : 4;
  const int C = std::min(chunk_steps, T);
  if (!h->copy_stream) {
    VJF_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    VJF_CUDA_OK(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      VJF_CUDA_OK(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
      VJF_CUDA_OK(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
  }
  // device staging: y/u/eps double-buffered per chunk; trajectory for the whole chunk pair
  const size_t yb = (size_t)C * B * D * ysz, ub = (size_t)C * B * u * 4, eb = (size_t)C * 2 * B * d * 4;
  const size_t tb = (size_t)(C + 1) * B * d * 4, lb = (size_t)C * 4 * 4;
  for (int i = 0; i < 2; ++i) {
    if (ensure(&h->stage_y[i], &h->stage_y_sz[i], yb)) return -2;
    if (u && ensure((void**)&h->stage_u[i], &h->stage_u_sz[i], ub)) return -2;
    if (eps_host && ensure((void**)&h->stage_eps[i], &h->stage_eps_sz[i], eb)) return -2;
  }
  // trajectory staging holds [carry | C steps] for two chunks in flight
  if (ensure((void**)&h->stage_mu, &h->stage_mu_sz, 2 * tb)) return -2;
  if (ensure((void**)&h->stage_lv, &h->stage_lv_sz, 2 * tb)) return -2;
  if (ensure((void**)&h->stage_loss, &h->stage_loss_sz, 2 * lb)) return -2;
  cudaStream_t cs = h->copy_stream, ks = h->compute_stream;
  const int nchunks = (T + C - 1) / C;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1, t0 = c * C, n = std::min(C, T - t0);
    if (c >= 2) VJF_CUDA_OK(cudaStreamWaitEvent(cs, h->ev_done[buf], 0));
    VJF_CUDA_OK(cudaMemcpyAsync(h->stage_y[buf], (const char*)y_host + (size_t)t0 * B * D * ysz, (size_t)n * B * D * ysz,
                                cudaMemcpyHostToDevice, cs));
    if (u) VJF_CUDA_OK(cudaMemcpyAsync(h->stage_u[buf], u_host + (size_t)t0 * B * u, (size_t)n * B * u * 4, cudaMemcpyHostToDevice, cs));
    if (eps_host) VJF_CUDA_OK(cudaMemcpyAsync(h->stage_eps[buf], eps_host + (size_t)t0 * 2 * B * d, (size_t)n * 2 * B * d * 4, cudaMemcpyHostToDevice, cs));
    VJF_CUDA_OK(cudaEventRecord(h->ev_copied[buf], cs));
    VJF_CUDA_OK(cudaStreamWaitEvent(ks, h->ev_copied[buf], 0));
    // trajectory slab of this chunk: row 0 = carry (q of the last step of the previous chunk)
    float* mu_d = h->stage_mu + (size_t)buf * (C + 1) * B * d;
    float* lv_d = h->stage_lv + (size_t)buf * (C + 1) * B * d;
    float* loss_d = h->stage_loss + (size_t)buf * C * 4;
    if (c > 0) {
      const int pb = (c - 1) & 1, pn = std::min(C, T - (c - 1) * C);
      VJF_CUDA_OK(cudaMemcpyAsync(mu_d, h->stage_mu + ((size_t)pb * (C + 1) + pn) * B * d, (size_t)B * d * 4, cudaMemcpyDeviceToDevice, ks));
      VJF_CUDA_OK(cudaMemcpyAsync(lv_d, h->stage_lv + ((size_t)pb * (C + 1) + pn) * B * d, (size_t)B * d * 4, cudaMemcpyDeviceToDevice, ks));
    }
    const uint32_t fl = (c == 0) ? flags : (flags & ~(uint32_t)VJF_FLAG_PRIOR_Q0);
    const int rc = sharded
        ? vjf_run_sharded(h, n, B, B_global, trial_offset, h->stage_y[buf], y_dtype, u ? h->stage_u[buf] : nullptr, mu_d, lv_d,
                          eps_host ? h->stage_eps[buf] : nullptr, seed, step0 + t0, fl, lr, mu_d + (size_t)B * d, lv_d + (size_t)B * d, loss_d, ks)
        : vjf_run(h, n, B, h->stage_y[buf], y_dtype, u ? h->stage_u[buf] : nullptr, mu_d, lv_d, eps_host ? h->stage_eps[buf] : nullptr, seed,
                  step0 + t0, fl, lr, mu_d + (size_t)B * d, lv_d + (size_t)B * d, loss_d, ks);
    if (rc) return -2;
    VJF_CUDA_OK(cudaMemcpyAsync(mu_host + (size_t)t0 * B * d, mu_d + (size_t)B * d, (size_t)n * B * d * 4, cudaMemcpyDeviceToHost, ks));
    VJF_CUDA_OK(cudaMemcpyAsync(logvar_host + (size_t)t0 * B * d, lv_d + (size_t)B * d, (size_t)n * B * d * 4, cudaMemcpyDeviceToHost, ks));
    if (losses_host) VJF_CUDA_OK(cudaMemcpyAsync(losses_host + (size_t)t0 * 4, loss_d, (size_t)n * 16, cudaMemcpyDeviceToHost, ks));
    VJF_CUDA_OK(cudaEventRecord(h->ev_done[buf], ks));
  }
  VJF_CUDA_OK(cudaStreamSynchronize(ks));
  VJF_CUDA_OK(cudaStreamSynchronize(cs));
  return 0;
}

extern "C" int vjf_run_host(vjf_handle* h, int32_t T, int32_t B, const void* y_host, int32_t y_dtype, const float* u_host,
                            const float* eps_host, uint64_t seed, uint64_t step0, uint32_t flags, float lr, float* mu_host,
                            float* logvar_host, float* losses_host, int32_t chunk_steps) {
  return run_host_impl(h, 0, T, B, B, 0, y_host, y_dtype, u_host, eps_host, seed, step0, flags, lr, mu_host, logvar_host, losses_host, chunk_steps);
}

extern "C" int vjf_run_sharded_host(vjf_handle* h, int32_t T, int32_t B_local, int32_t B_global, uint64_t trial_offset, const void* y_host,
                                    int32_t y_dtype, const float* u_host, const float* eps_host, uint64_t seed, uint64_t step0, uint32_t flags,
                                    float lr, float* mu_host, float* logvar_host, float* losses_host, int32_t chunk_steps) {
  if (!h || h->comm_world < 1) { vjf_set_error("vjf_comm_connect has not been called"); return -1; }
  return run_host_impl(h, 1, T, B_local, B_global, trial_offset, y_host, y_dtype, u_host, eps_host, seed, step0, flags, lr, mu_host, logvar_host,
                       losses_host, chunk_steps);
}
