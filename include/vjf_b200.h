/*
 * vjf_b200.h -- C ABI of the B200-native VJF filter/learning step.
 *
 * The reference (catniplab/vjf) is pure Python: its only boundary for this path is the Python API
 * VJF.make_model / VJF.filter / VJF.fit (vjf/model.py:309-319, :179-221, :223-307).  This library is
 * what a Python shim (vjf_b200/model.py, or the stub shown in INTEGRATION.md) binds underneath those
 * three calls with ctypes.  No torch types cross this boundary: plain pointers, sizes and a CUDA
 * stream handle.  All device pointers are BORROWED fp32 pointers to contiguous row-major buffers; the
 * library allocates its workspace once in vjf_create() and nothing per step.
 *
 * Every function returns 0 on success, <0 on an argument/CUDA error (vjf_last_error() has the text).
 * Numerical events (non-finite ELBO terms, failed Cholesky) never fail a call -- as in the reference
 * (vjf/model.py:138-145, vjf/module.py:104-112) -- they are OR-ed into a device status word that
 * vjf_get_status() reads back.
 */
#ifndef VJF_B200_H
#define VJF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VJF_MAX_LAYERS 4
#define VJF_MAX_XDIM 16

/* likelihood ids: PoissonLikelihood / GaussianLikelihood, vjf/likelihood.py:43-66 / :9-40 */
#define VJF_LIK_POISSON 0
#define VJF_LIK_GAUSSIAN 1

/* flags of one step == keyword arguments of VJF.filter (vjf/model.py:179-180) */
#define VJF_FLAG_SGD 1            /* sgd=True : backward + clip_grad_value_(1) + SGD step (:206-211) */
#define VJF_FLAG_UPDATE 2         /* update=True: running noise variances + RLS (:215-216, :156-177) */
#define VJF_FLAG_WARMUP 4         /* warm_up=True: no dynamics term in the loss, no RLS (:148-149, :370-371) */
#define VJF_FLAG_DECODER_FROZEN 8 /* decoder.requires_grad_(False) after warm-up (:283) */
#define VJF_FLAG_PRIOR_Q0 16      /* qs=None: first step starts from the prior (:107-108, :80-95) */

/* status bits (device status word) */
#define VJF_ST_RECON_NONFINITE 1   /* l_recon non-finite -> replaced by 0, no gradient (model.py:138-139) */
#define VJF_ST_DYN_NONFINITE 2     /* l_dynamics non-finite (model.py:141-142) */
#define VJF_ST_ENTROPY_NONFINITE 4 /* entropy non-finite (model.py:144-145) */
#define VJF_ST_MSE_NONFINITE 8     /* the reference would trip `assert isfinite(mse)` (functional.py:60) */
#define VJF_ST_CHOL_FAILED 16      /* RLS Cholesky pivot <= 0: state left unchanged ('RLS failed.', module.py:112) */
#define VJF_ST_COMM_TIMEOUT 32     /* a peer's contribution did not arrive within the timeout (sharded run) */

/* dtype of the observation buffer handed to vjf_run / vjf_run_host */
#define VJF_Y_F32 0
#define VJF_Y_U8 1 /* spike counts 0..255 (Poisson); converted on load */

typedef struct vjf_config {
  int32_t ydim, xdim, udim, n_rbf;     /* VJF.make_model(ydim, xdim, udim, n_rbf, ...) */
  int32_t n_layers;                    /* len(hidden_sizes) */
  int32_t hidden[VJF_MAX_LAYERS];      /* hidden_sizes */
  int32_t likelihood;                  /* VJF_LIK_* */
  int32_t max_trials;                  /* largest batch (trials) a step will see on this device */
} vjf_config;

/* Offsets (in floats) of every tensor inside the flat fp32 state buffer.  Linear weights are stored
 * INPUT-major, weight_t[in][out] (the transpose of torch's nn.Linear.weight[out][in]), so that the
 * kernels read them coalesced; the Python shim exposes them as transposed views.
 * The trainable block [0, n_train) has the same layout as the gradient vector. */
typedef struct vjf_layout {
  int64_t lik_logvar;                  /* 1      likelihood.logvar (Gaussian)            likelihood.py:16 */
  int64_t dec_w, dec_b;                /* [d][D], [D]   decoder.decode                    model.py:24      */
  int64_t mlp_w[VJF_MAX_LAYERS];       /* [in_l][H_l]   recognition.mlp.{2l}.weight^T     recognition.py:20-25 */
  int64_t mlp_b[VJF_MAX_LAYERS];       /* [H_l] */
  int64_t head_m_w, head_v_w, head_v_b;/* [H_L][d] x2, [d]  recognition.mean / .logvar    recognition.py:27-28 */
  int64_t n_train;                     /* size of the trainable block */
  int64_t prior_mean, prior_logvar;    /* [d] each       VJF.mean / VJF.logvar            model.py:66-67   */
  int64_t tr_logvar;                   /* 1      transition.logvar (state noise)          model.py:331     */
  int64_t centroid, logwidth;          /* [R][d+u], [R]  RBF                              module.py:20-21  */
  int64_t w_mean, w_chol, w_precision, w_pchol; /* [R][d], [R][R] x3  LinearRegression    module.py:46-54  */
  int64_t lik_n, tr_n;                 /* sample counters, stored as exact floats         likelihood.py:17, model.py:332 */
  int64_t total;                       /* floats in the state buffer */
} vjf_layout;

typedef struct vjf_handle vjf_handle;

const char* vjf_last_error(void);
int vjf_version(void);

/* Fill `out` for a configuration (pure host arithmetic, needs no GPU). */
int vjf_get_layout(const vjf_config* cfg, vjf_layout* out);

/* Create the per-model context on the current CUDA device.  `state` is the flat device buffer
 * (vjf_layout.total floats) that holds every parameter; it stays owned by the caller and is updated
 * in place by the step functions.  Replaces VJF.__init__'s module/optimizer wiring (model.py:51-78). */
int vjf_create(const vjf_config* cfg, float* state, vjf_handle** out);
int vjf_destroy(vjf_handle* h);

/* Write the reference's initial values for the non-random tensors into `state` on `stream`
 * (W=0, w_chol=P=w_pchol=I, logwidth=0, obs logvar=log 0.1, state logvar=0, prior 0). */
int vjf_init_state(vjf_handle* h, void* stream);

/*
 * One filter step == VJF.filter(y, u, qs, sgd=, update=, warm_up=) (vjf/model.py:179-221).
 *   y        [B][ydim]   observations of this time step
 *   u        [B][udim]   control input, or NULL (udim == 0)
 *   q_mean, q_logvar [B][xdim]  previous posterior; ignored with VJF_FLAG_PRIOR_Q0
 *   eps      [2][B][xdim] N(0,1) draws for xs then xt (vjf/util.py:11-13), or NULL to draw them in
 *            the kernel from Philox4x32-10 keyed by (seed, step_index)
 *   out_mean, out_logvar [B][xdim]  the new posterior q_t
 *   out_loss [4] = loss, -l_recon, -l_dynamics, entropy (the verbose tuple, model.py:218-219)
 * Parameters and RLS state inside `state` are updated in place.
 */
int vjf_step(vjf_handle* h, int32_t B, const float* y, const float* u, const float* q_mean, const float* q_logvar,
             const float* eps, uint64_t seed, uint64_t step_index, uint32_t flags, float lr, float* out_mean,
             float* out_logvar, float* out_loss, void* stream);

/*
 * T consecutive steps in ONE persistent launch == the time loop of VJF.fit (vjf/model.py:252-261):
 * step t reads q_{t-1} from the trajectory it is writing.
 *   y [T][B][ydim] (dtype y_dtype), u [T][B][udim] or NULL, eps [T][2][B][xdim] or NULL (Philox),
 *   q0_mean/q0_logvar: posterior before the first step (ignored with VJF_FLAG_PRIOR_Q0),
 *   mu, logvar [T][B][xdim]: the filtered trajectory, losses [T][4].
 */
int vjf_run(vjf_handle* h, int32_t T, int32_t B, const void* y, int32_t y_dtype, const float* u, const float* q0_mean,
            const float* q0_logvar, const float* eps, uint64_t seed, uint64_t step0, uint32_t flags, float lr,
            float* mu, float* logvar, float* losses, void* stream);

/*
 * Same as vjf_run but every buffer except the model state is a HOST pointer (pinned or pageable):
 * observations are streamed host->device in chunks on a copy stream overlapped with the compute of
 * the previous chunk, the trajectory and losses are copied back; returns after everything landed.
 * This is the end-to-end entry the `e2e` benchmark number times.
 */
int vjf_run_host(vjf_handle* h, int32_t T, int32_t B, const void* y_host, int32_t y_dtype, const float* u_host,
                 const float* eps_host, uint64_t seed, uint64_t step0, uint32_t flags, float lr, float* mu_host,
                 float* logvar_host, float* losses_host, int32_t chunk_steps);

/* ---- split step for trials sharded over several GPUs (one process per GPU) ----
 * phase A: trial-parallel forward + hand-derived backward of the LOCAL trials; leaves the local sums
 *          (gradients, RLS statistics phi^T phi / phi^T dx, loss sums) in the device vector
 *          vjf_reduce_buffer() of vjf_reduce_size() floats.  The caller all-reduces (sum) that vector
 *          across ranks -- the single exchange of the step -- then calls phase B.
 * phase B: clip + SGD, running variances, RLS factorisation on the reduced sums, identically on
 *          every rank.  B_global = total trials over all ranks (the reference's batch means). */
int64_t vjf_reduce_size(vjf_handle* h);
float* vjf_reduce_buffer(vjf_handle* h);
int vjf_step_phase_a(vjf_handle* h, int32_t B_local, int32_t B_global, const float* y, int32_t y_dtype, const float* u,
                     const float* q_mean, const float* q_logvar, const float* eps, uint64_t seed, uint64_t step_index,
                     uint64_t trial_offset, uint32_t flags, float* out_mean, float* out_logvar, void* stream);
int vjf_step_phase_b(vjf_handle* h, int32_t B_global, uint32_t flags, float lr, float* out_loss, void* stream);

/* ---- trials sharded over GPUs, exchange INSIDE the persistent kernel over NVLink peer memory ----
 * Every rank (one process per GPU) owns an exchange buffer; vjf_comm_local_handle() returns its 64-byte CUDA IPC
 * handle, the caller gathers the handles of all ranks (torch.distributed, MPI, ...) and passes them to
 * vjf_comm_connect().  vjf_run_sharded() is vjf_run() for the local block of trials: after the local slot
 * reduction each CTA pushes its 512-byte chunk of sums to every peer's inbox with peer stores, waits for the
 * matching chunks of the other ranks, adds them in rank order (=> bit-identical replicas) and goes on with the
 * SGD / RLS phases -- compute and the all-reduce live in one kernel, no NCCL call on the step path.
 * All ranks must call vjf_run_sharded with the same T / flags / lr in the same order. */
#define VJF_MAX_RANKS 8
int vjf_comm_local_handle(vjf_handle* h, void* out_handle64);
int vjf_comm_connect(vjf_handle* h, int32_t rank, int32_t world, const void* handles /* [world][64] */);
int vjf_run_sharded(vjf_handle* h, int32_t T, int32_t B_local, int32_t B_global, uint64_t trial_offset, const void* y,
                    int32_t y_dtype, const float* u, const float* q0_mean, const float* q0_logvar, const float* eps,
                    uint64_t seed, uint64_t step0, uint32_t flags, float lr, float* mu, float* logvar, float* losses,
                    void* stream);

/* vjf_run_host for the local block of trials of a sharded run: chunked, double-buffered host->device copies of this rank's
 * observations overlapped with the fused compute + exchange kernel of the previous chunk, device->host copies of the
 * trajectory and losses.  All ranks must call it with the same T / chunk_steps / flags / lr. */
int vjf_run_sharded_host(vjf_handle* h, int32_t T, int32_t B_local, int32_t B_global, uint64_t trial_offset, const void* y_host,
                         int32_t y_dtype, const float* u_host, const float* eps_host, uint64_t seed, uint64_t step0, uint32_t flags,
                         float lr, float* mu_host, float* logvar_host, float* losses_host, int32_t chunk_steps);

/* Precision of the recursive least squares (LinearRegression.rls, vjf/module.py:79-112): 32 (default; the reference's default
 * dtype) or 64.  With 64 the accumulated w_precision is shadowed in double and the factorisation runs in double: the reference
 * reaches that robustness by switching the whole model to float64 (script/example.py:12); the fp32 recursion stops absorbing
 * samples and its Cholesky fails after ~1e7 samples (VJF_ST_CHOL_FAILED).  Needed for long horizons (T = 100 000). */
int vjf_set_rls_precision(vjf_handle* h, int32_t bits);

/* n_rbf > 160 (BASELINE config 3, n_rbf = 1024) runs a different schedule under the same entry points (vjf_step / vjf_run /
 * vjf_run_host): per time step a short sequence of launches with phi w_chol and phi^T phi as tcgen05 GEMMs over all trials and a
 * blocked multi-CTA Cholesky (csrc/bigr.cu); vjf_last_launch_kind() reports 2.  Not available there: the sharded run,
 * vjf_rls_initialize, vjf_weight_kalman, the double-precision RLS.  vjf_bigr_buffer returns device views of its workspace for tests
 * and profiling: 0 phi [B][n_rbf], 1 p_mean [B][xdim], 2 p_logvar [B], 3 split-K partial sums of phi^T phi; NULL otherwise. */
float* vjf_bigr_buffer(vjf_handle* h, int32_t which);

/* Wide observations (ydim above the tile pipeline's 480 columns, up to 2048; BASELINE config 4: ydim 2000, xdim 8): one hidden
 * layer of <= 128 units, xdim <= 8, n_rbf <= 128, fp32 observations run a per-step launch sequence under the same entry points
 * (vjf_step / vjf_run / vjf_run_host / vjf_run_sharded): the two contractions over the observation columns (recognition layer-1
 * forward, vjf/recognition.py:38, and its weight gradient) as tcgen05 GEMMs over all trials of the step, one trial-parallel kernel
 * for the rest, the serial RLS half of the step on a side stream beside the next step's forward GEMM (csrc/wide.cu);
 * vjf_last_launch_kind() reports 3.  Sharded: the reduced vector is summed by a pull all-reduce over the peer-mapped exchange
 * buffers of vjf_comm_connect.  A non-finite ELBO term skips the SGD step (the split path's rule, vjf/model.py:212-214). */

/* development aid of the wide-observation path: device pointer of its globaltimer stamps (8 steps x 8 words: RLS start / end,
 * mid kernel start / wait entered / wait left / end of CTA 0, SGD end, end of the last mid CTA), NULL unless VJF_WIDE_STAMPS was set
 * in the environment when the handle was created (scripts/c4_time.py prints the timeline) */
unsigned long long* vjf_wide_stamps(vjf_handle* h);

/* status word: OR of VJF_ST_* since the last clear (synchronises the stream) */
int vjf_get_status(vjf_handle* h, void* stream, uint32_t* out, int32_t clear);

/* Fill eps[n] with the same Philox N(0,1) stream the kernels draw in-kernel:
 * element (step, which in {0,1}, trial, i) for trial in [trial_offset, trial_offset+B). */
int vjf_philox_normal(uint64_t seed, uint64_t step_index, uint64_t trial_offset, int32_t B, int32_t xdim, float* eps_out,
                      void* stream);

/* kernel-launch accounting for the benchmark (`gpu_launches`) */
int64_t vjf_launch_count(void);
/* which kernel ran the most recent vjf_run / vjf_step / vjf_run_sharded time loop: 1 = throughput tile pipeline (every
 * contraction on tcgen05, observations by TMA tensor copies), 0 = the general persistent kernel (shapes outside the tile plan:
 * several hidden layers, hidden width not a multiple of 32, uint8 observations, ...), 2 = large-n_rbf launch sequence, 3 = wide-
 * observation launch sequence */
int32_t vjf_last_launch_kind(void);
/* kernel selection of the time loop (process-wide; for tests and A/B timing): 0 automatic (default), 1 general persistent kernel
 * only (no tile pipeline, no wide-observation path), 2 tile pipeline without its exact-observation mode, 3 tile pipeline with 64-trial tiles whenever the observations are
 * exact in tf32 (spike counts) */
int vjf_set_tile_mode(int32_t mode);

/* ---- whole-trajectory RLS re-initialisation: RBFDS.initialize + LinearRegression.initialize
 * (vjf/model.py:379-388, vjf/module.py:144-150).  xs, xt [N][xdim], u [N][udim] or NULL; the caller
 * has already written the re-drawn centroids / logwidth into `state`.  Streams phi^T phi over N
 * samples instead of the reference's (N,N) temporary. */
int vjf_rls_initialize(vjf_handle* h, int64_t N, const float* xs, const float* xt, const float* u, void* stream);

/* ---- weight-space Kalman update with diffusion: LinearRegression.kalman (vjf/module.py:114-142) = kalman.predict with A = I,
 * Q = diffusion * I followed by kalman.joseph_update AS WRITTEN (vjf/kalman.py:102-145) with H = features of [xs, u] and
 * R = v * I.  xs [N][xdim], u [N][udim] or NULL, target [N][xdim].  Updates w_mean and w_chol (lower Cholesky factor of the
 * posterior weight covariance) in `state`.  Evaluated in information form from the streamed statistics phi^T phi, phi^T target:
 * no N x N innovation covariance. */
int vjf_weight_kalman(vjf_handle* h, int64_t N, const float* xs, const float* target, const float* u, float v, float diffusion,
                      void* stream);

/* ---- forecast: RBFDS.forecast with sampled weights (vjf/model.py:342-361, module.py:70-73).
 * x [n_step+1][B][xdim] (x[0] given), yhat [n_step+1][B][ydim] or NULL, w_eps [n_step][R][xdim],
 * x_eps [n_step][B][xdim] or NULL (noise=False). */
int vjf_forecast(vjf_handle* h, int32_t n_step, int32_t B, float* x, float* yhat, const float* u, const float* w_eps,
                 const float* x_eps, void* stream);

/* ---- batched small-matrix Kalman operator (vjf/kalman.py:15-145, vjf/numerical.py:8-19).
 * P independent problems, one warp each; problem p uses x[p] (n x nb, state-major as in the
 * reference), L[p] (n x n Cholesky factor), A[p], Q[p] (n x n), H[p] (m x n), R[p] (m x m), y[p] (m x nb).
 * Limits: n <= 16, m <= 32, nb <= 32.  joseph_update reproduces the reference AS WRITTEN (S^-1 applied
 * twice, elementwise sqrt(R)); see SURVEY.md 8a K3. */
int vjf_kalman_predict_batched(int32_t P, int32_t n, int32_t m, int32_t nb, const float* x, const float* L,
                               const float* A, const float* Q, const float* H, float* yhat, float* xhat, float* Lhat,
                               int32_t* info, void* stream);
int vjf_kalman_update_batched(int32_t P, int32_t n, int32_t m, int32_t nb, const float* y, const float* yhat,
                              const float* xhat, const float* Lhat, const float* H, const float* R, float* x_out,
                              float* L_out, int32_t* info, void* stream);
int vjf_kalman_joseph_update_batched(int32_t P, int32_t n, int32_t m, int32_t nb, const float* y, const float* yhat,
                                     const float* xhat, const float* Lhat, const float* H, const float* R, float* x_out,
                                     float* L_out, int32_t* info, void* stream);
/* symmetrize: upper triangle mirrored down; positivize: eigen-clamp (>= eps) via cyclic Jacobi */
int vjf_symmetrize_batched(int32_t P, int32_t n, const float* a, float* out, void* stream);
int vjf_positivize_batched(int32_t P, int32_t n, const float* a, float eps, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VJF_B200_H */
