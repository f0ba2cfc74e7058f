/*
 * nma_b200.h — C-ABI of the B200-native NMA (Neural Moving Average) ELBO step.
 *
 * The reference (mehrnazmo/VIforSSMs) has no FFI: its only process/device seam
 * is the TensorFlow session call
 *     sess.run([self.train_step, self.merged], feed_dict={time_feats, mask, shift})
 * (AR.py:300-301; fitz_nag_NVP.py:389-390; SV_dense.py:341-342).  Everything that
 * one call executes — window gather, NMA flow, ELBO terms, gradients, clip +
 * Adamax — is what this library replaces.  Each entry point cites the reference
 * lines it stands in for.
 *
 * Conventions: every pointer named d_* is a DEVICE pointer owned by the caller;
 * every call is asynchronous on the given cudaStream_t (passed as void*), is
 * CUDA-graph capturable, allocates nothing after nma_create, returns 0 on
 * success and a negative code on failure (message via nma_last_error()).
 * No exception ever crosses this boundary.
 */
#ifndef NMA_B200_H
#define NMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMA_MAX_CHAN   32
#define NMA_MAX_ARRAYS 8
#define NMA_MAX_FLOWS  8

/* model kinds (which _ELBO is evaluated) */
#define NMA_MODEL_AR  0   /* AR.py:168-187 */
#define NMA_MODEL_FHN 1   /* fitz_nag_NVP.py:232-266 */
#define NMA_MODEL_SV  2   /* SV_dense.py:203-234 */
#define NMA_MODEL_LV  3   /* lotka_volterra_partial_batch_fix_theta.py:265-371 (fixed theta) */
#define NMA_MODEL_LVR 4   /* lotka_volterra_partial.py:234-297 (learned theta): same flow kernels as NMA_MODEL_LV, own ELBO branch */
#define NMA_MODEL_LVB 5   /* lotka_volterra_partial_batch.py:300-371 (learned softplus-theta, p_val windows): NMA_MODEL_LV's
                           * flow, observation and x0 terms, the plain bivariate transition density, d/dtheta of all of it */

/* objectives (which scalar is differentiated) */
#define NMA_OBJ_ELBO    0 /* -sum_rows scale*(sde - logq + obs)   AR.py:184-185,228-229 */
#define NMA_OBJ_NEG_OBS 1 /* -sum_rows obs_log_prob               AR.py:201-202 (pre-train) */
#define NMA_OBJ_PATH_SQ 2 /* sum (lf_sample - target)^2           fitz_nag_NVP.py:288-289; SV_dense.py:251-252 */

typedef struct nma_config {
    int32_t model;
    int32_t p;          /* max rows per step: MC samples == subsequences (AR.py:117,263-265) */
    int32_t K;          /* kernel_len */
    int32_t B;          /* batch_dims */
    int32_t D;          /* flow_dims: 1, or 2 = two latent components interleaved on the time axis */
    int32_t F;          /* no_flows */
    int32_t C;          /* network_dims[0]; must be 50 */
    int32_t H;          /* hidden 1x1 layers = len(network_dims)-2 */
    int32_t bn;         /* inference-mode batch-norm affine after each hidden layer (fitz_nag_NVP.py:93) */
    int32_t Cf;         /* channels of time_feats */
    int32_t feat_aug;   /* SV_dense.py:53: features at slot+1 plus first differences of the first Cf-2 */
    int32_t dtheta;
    int32_t n_arrays;   /* number of base arrays given to nma_set_series */
    int32_t obs_array;  /* base array evaluated by the observation term */
    int32_t bin_array;  /* base array holding the observation indicator */
    int32_t head_offset;
    int32_t chan_array[NMA_MAX_CHAN];  /* time_feats[r, j, c] = base[chan_array[c]][D*idx[r] + j + chan_offset[c]] */
    int32_t chan_offset[NMA_MAX_CHAN];
    double  scale;      /* T / batch_dims (AR.py:184) */
    float   dt;
    float   obs_std;
    float   x0[2];
    int32_t n_pinned;   /* leading states of the series pinned to x0 by mask / shift: 1 (0 reads as 1); p_val for the
                         * Lotka-Volterra batch scripts (lotka_volterra_partial_batch.py:237-240) */
    int32_t reserved;
} nma_config;

typedef struct nma_handle_s* nma_handle;

const char* nma_last_error(void);
int nma_version(void);

/* Model assembly — VI_SSM.__init__ + build_flow (AR.py:115-159,189-238): sizes every workspace. */
int nma_create(const nma_config* cfg, nma_handle* out);
int nma_destroy(nma_handle h);

/* number of fp32 values in the flat parameter blob / per-flow section table
 * (TF variable creation order, AR.py:53-78).  offsets: int64[F * 32] (see nma_api.cu). */
int64_t nma_param_count(nma_handle h);
int nma_param_layout(nma_handle h, int64_t* offsets_out, int32_t n);
/* bytes of device workspace held by the handle */
int64_t nma_workspace_bytes(nma_handle h);

/* Padded base arrays exactly as VI_SSM.__init__ builds them (AR.py:135-150), already cast to fp32
 * (the feed_dict cast, AR.py:300-301).  Pointers are borrowed until nma_destroy / the next call. */
int nma_set_series(nma_handle h, const float* const* d_arrays, const int64_t* lengths, int32_t n_arrays);

/* A1 — window gather (AR.py:267-288): d_time_feats [p,L0,Cf], d_mask/d_shift [p,D,B+1] (may be NULL). */
int nma_gather(nma_handle h, const int64_t* d_idx, int32_t p, float* d_time_feats, float* d_mask,
               float* d_shift, void* stream);

/* A2-A8 — one fused ELBO + gradient evaluation (the body of sess.run(train_step), AR.py:300 →
 * AR.py:44-110,168-187,226-229).  d_eps [p,L0] and d_idx [p] are injected by the host
 * (indices stay bit-exact numpy draws, AR.py:263-265).
 *   d_terms       [p,4]  sde_log_prob, obs_log_prob, lf_log_prob (logq), base_log_prob
 *   d_lf          [p,L_F] final flow sample (lf_sample before any reshape)
 *   d_grad_params [n_params]  d objective / d params   (overwritten)
 *   d_grad_theta  [p,dtheta]  d objective / d theta    (overwritten)
 *   d_flags       [p]    bit0 = non-finite ELBO term in that row (fitz_nag_NVP.py:378-381 needs it)
 * d_eps == NULL: the library draws the base noise itself, as the reference does in-graph (AR.py:31-35): Philox4x32-10 +
 * Box-Muller keyed by (seed, device-resident draw counter) - nma_set_seed; nma_philox_normal(stream_id 0) reproduces it.
 * On a handle with a communicator (nma_comm_create / nma_comm_init) the call also issues the all-reduce of every flow's
 * gradient section on the library's side stream; nma_comm_wait(h, stream) orders `stream` behind them.
 */
int nma_elbo_fwd_bwd(nma_handle h, const float* d_params, const float* d_eps, const float* d_theta,
                     const int64_t* d_idx, int32_t p, int32_t objective, float path_target,
                     float* d_terms, float* d_lf, float* d_grad_params, float* d_grad_theta,
                     uint32_t* d_flags, void* stream);

/* forward only — save_paths (AR.py:323-362; fitz_nag_NVP.py:409-448) */
int nma_forward_paths(nma_handle h, const float* d_params, const float* d_eps, const float* d_theta,
                      const int64_t* d_idx, int32_t p, float* d_terms, float* d_lf, void* stream);

/* ---- the whole iteration: what ONE sess.run([self.train_step, self.merged], feed_dict) executes (AR.py:300-301) ----
 * noise (in-graph in the reference, AR.py:31-35,117-118) -> theta ~ q(theta), log q(theta) (AR.py:376-391) -> flow, ELBO
 * terms, gradients (AR.py:44-110,168-187,228-229) incl. log prior(theta) - log q(theta) (AR.py:178-185) and the backward
 * pass through the theta posterior -> gradient all-reduce over the time shards (if a communicator is set) ->
 * tf.global_norm + clip_by_global_norm + Adamax over EVERY variable (AR.py:230-234; optimisers/adamax.py:42-58) -> the
 * logged scalars (AR.py:207-224).  No host round trip, no library kernel in between; CUDA-graph capturable.
 *   d_blob / d_grad / d_m / d_v : [nma_param_count(h) + nma_theta_flow_param_count(dtheta, nb)]  NMA variables followed by
 *                                 the theta posterior's (creation order: per layer 4 x (kernel [in][out], bias))
 *   d_scalars [8] : mean ELBO, scale*mean sde, mean log q(theta), scale*mean obs, scale*mean log q(path), global norm,
 *                   rows with a non-finite term, draw counter after the step
 *   d_theta_out   : [p][dtheta] the theta sample (may be NULL);  d_lf_out : [p][L_F] the final flow sample (may be NULL)
 * nma_set_theta_flow with nb = 0 declares "no posterior": theta is the constant prior_mean in every row
 * (lotka_volterra_partial_batch_fix_theta.py:190), the blob holds the NMA variables only.                              */
typedef struct nma_step_opts {
    int32_t objective;      /* NMA_OBJ_* */
    float   path_target;
    int32_t prior_on;       /* 1: + log prior(theta) - log q(theta) in the objective (AR.py:184-185); 0: pre-training heads */
    int32_t obs_in_elbo;    /* 1: the observation term is part of the ELBO (0 for SV_dense.py:238-241) */
    int32_t tf_mask_grad;   /* 1: TensorFlow's masked_dense semantics - masked kernel entries of the theta posterior get a
                             *    gradient (it counts in tf.global_norm) and are reset by the kernel constraint after the update */
    float   lr, beta1, beta2, eps, clip;   /* clip <= 0: no clipping (the pre-train optimisers, AR.py:201-202) */
} nma_step_opts;
/* theta posterior of AR.py:376-391 (see nma_theta_flow_fwd below for the layouts) and the diagonal Gaussian prior of
 * AR.py:178-182 (host arrays of dtheta floats).  Device pointers are borrowed until nma_destroy. */
int nma_set_theta_flow(nma_handle h, const float* d_masks, const int32_t* d_perms, int32_t nb, int32_t relu,
                       float base_loc, float base_scale, const float* prior_mean, const float* prior_scale,
                       int32_t softplus_out);
/* softplus_out = 1: the chain ends in tfb.Softplus (lotka_volterra_partial_batch.py:741) - theta = softplus(flow output),
 * log q carries its Jacobian - and the prior is the Softplus-transformed diagonal Gaussian of ibid. :358-365. */
int64_t nma_theta_flow_param_count(int32_t dtheta, int32_t nb);
int nma_train_step(nma_handle h, float* d_blob, float* d_grad, float* d_m, float* d_v, const int64_t* d_idx, int32_t p,
                   const nma_step_opts* opts, float* d_scalars, float* d_theta_out, float* d_lf_out, void* stream);
/* seed and draw counter of the in-library noise (the counter lives on the device and advances once per call that draws) */
int nma_set_seed(nma_handle h, uint64_t seed, uint64_t counter);
int nma_get_counter(nma_handle h, uint64_t* counter_out);
/* the normals a call draws for (seed, counter): stream_id 0 = eps [p*L0] (loc 0, scale 1), 1 = the theta posterior's base
 * sample [p*dtheta] (loc, scale of the base distribution) */
int nma_philox_normal(float* d_out, int64_t n, uint64_t seed, uint64_t counter, uint32_t stream_id, float loc, float scale,
                      void* stream);
/* borrowed views of what the last nma_train_step left in the workspace (any pointer argument may be NULL) */
int nma_step_buffers(nma_handle h, float** d_eps, float** d_z0, float** d_theta, float** d_logq_theta, float** d_terms,
                     float** d_row_elbo);

/* ---- multi-GPU: the step's one collective, issued by the library on its own side stream ----
 * nma_comm_unique_id + nma_comm_create build an NCCL communicator owned by the handle (every rank calls create with the id
 * rank 0 made); nma_comm_init adopts an ncclComm_t the caller owns.  libnccl.so.2 is resolved at run time.  Destroy every
 * CUDA graph that captured a step before nma_comm_destroy / nma_destroy. */
int nma_comm_unique_id(char* out128);
int nma_comm_create(nma_handle h, const char* id128, int32_t rank, int32_t world);
int nma_comm_init(nma_handle h, void* nccl_comm);
int nma_comm_destroy(nma_handle h);
int nma_comm_world(nma_handle h);
int nma_comm_wait(nma_handle h, void* stream);
int nma_comm_allreduce(nma_handle h, float* d_buf, int64_t count, void* stream);

/* The K-tap conv (AR.py:61-62) and its data gradient run on the tcgen05 tensor cores (3xTF32 split, fp32
 * accumulate) whenever the configuration allows it (flow_dims = 1, kernel_len <= 190); the FP32 SIMT kernels
 * remain as the correctness anchor.  on = 0 selects the SIMT conv, on = 1 the tensor-core conv (error if the
 * configuration does not support it).  The environment variable NMA_TC=0 sets the default to SIMT.
 * `on` is a bit set: bit 0 the conv, bit 1 the feature MLP and the head backward on tcgen05 as well, bit 2 the conv
 * GEMMs (forward, data gradient, weight gradient) in the 2-term bfloat16 split on kind::f16 - twice the tensor rate,
 * 16 instead of 22 significand bits per operand, measured 4e-6..8e-6 on the step's gradients (bar 1e-4); it needs
 * bits 0 and 1 and an AR-type model (flow_dims = 1, one hidden layer, no batch-norm), error otherwise.
 * NMA_TC_BF16=1 makes it the default where supported.  nma_get_tensor_cores returns the same bit set. */
int nma_set_tensor_cores(nma_handle h, int32_t on);
int nma_get_tensor_cores(nma_handle h);

/* Test hook: the bare tensor-core contraction on caller data.  d_in [Q][56] fp32, d_w [K][51][50] (conv1d kernel
 * layout), d_out [Q][64].  mode 0: out[q][n] = sum_k sum_c in[q+k][c] w[k][c][n]; mode 1 (data gradient):
 * out[q][n] = sum_k sum_f in[q+k][f] w[K-1-k][n][f].  nacc = 1 or 2 accumulators of 128 positions per CTA.
 * mode + 2: the same contraction in the bf16 split.  Synchronises the stream; allocates its own scratch. */
int nma_tc_conv_raw(const float* d_in, const float* d_w, int32_t mode, int32_t nacc, float* d_out, int64_t Q,
                    int32_t K, void* stream);

/* Test hook: bare tensor-core weight gradient.  d_in, d_da [Q][56] fp32, d_gw [K][51][50] (accumulated into):
 * gw[k][c][f] += sum_{q <= Q-K} in[q+k][c] * da[q][f]. */
int nma_tc_wgrad_raw(const float* d_in, const float* d_da, float* d_gw, int64_t Q, int32_t K, void* stream);
/* ... and in the bf16 split (MN-major operands straight from the conv operand layout) */
int nma_tc_wgrad_raw_bf(const float* d_in, const float* d_da, float* d_gw, int64_t Q, int32_t K, void* stream);

/* Measurement hooks (no reference counterpart): re-launch one stage of the last step on the workspace it
 * left behind (stage: 0 conv_fwd, 1 conv_dgrad, 2 conv_wgrad, 3 epi_bwd, 4 feat_fwd, 5 feat_bwd), and the
 * number of kernels this library has launched so far in this process. */
int nma_launch_stage(nma_handle h, int32_t stage, int32_t flow, const float* d_params, const float* d_eps,
                     const int64_t* d_idx, int32_t p, float* d_grad_params, void* stream);
int64_t nma_launch_count(void);

/* A9 — tf.global_norm + clip_by_global_norm + AdamaxOptimizer._apply_dense
 * (AR.py:230-234; optimisers/adamax.py:42-58).  d_norm_out[0] = global norm (pre-clip).
 * d_scratch: >= 1024 floats. */
int nma_adamax_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n,
                    float lr, float beta1, float beta2, float eps, float clip,
                    float* d_norm_out, float* d_scratch, void* stream);

/* A12 — AR(1) series simulation (AR_dat_gen.py:11-15) as an affine-map prefix scan: single pass, decoupled look-back,
 * warp-shuffle tile scan (16 B of HBM traffic per element).
 * x[0] = x0; x[i] = a*x[i-1] + b + c*z[i-1]  (i = 1..n);  d_z: n standard normals. d_x: n+1.
 * d_scratch: nma_scan_scratch_bytes(n) bytes, 16-byte aligned (cleared by the call).
 * nma_scan_affine: the same scan over per-element maps, x[i] = A[i-1]*x[i-1] + D[i-1] (the Euler-Maruyama recursions of
 * the stochastic-volatility generator, SV_dense.py:211-223, are of this form). */
int64_t nma_scan_scratch_bytes(int64_t n);
int nma_scan_affine(const double* d_A, const double* d_D, double* d_x, int64_t n, double x0, void* d_scratch,
                    int64_t scratch_bytes, void* stream);
int nma_scan_ar1(const double* d_z, double* d_x, int64_t n, double x0, double a, double b, double c,
                 void* d_scratch, int64_t scratch_bytes, void* stream);
/* A13 — hold-fill + time-till-next-observation (AR_dat_gen.py:17-31) for every-`impute`-th sampling. */
int nma_time_till(const double* d_obs, int64_t n, int32_t impute, double* d_obs_fill, double* d_obs_binary,
                  double* d_time_till, void* stream);

/* A11 on the device - the theta posterior of AR.py:376-391 (num_bijectors inverse masked-autoregressive-flow layers,
 * masked_autoregressive_default_template(hidden_layers=[5,5,5]), fixed permutations in between), one thread per row.
 * d_params: the flow's variables in creation order, per layer 4 x (kernel [in][out], bias); d_masks: the four block
 * masks of one layer, concatenated (d*5 + 25 + 25 + 5*2d floats); d_perms [nb-1][d]; relu: 0 = elu template, 1 = relu.
 *   fwd: z0 [p][d] ~ N(base_loc, base_scale) -> theta [p][d], log q(theta) [p]
 *   bwd: dL/dtheta [p][d], dL/dlogq [p] (may be null) -> d_g_params (ACCUMULATED into), d_g_z0 [p][d] (may be null)
 * Checked against the host autograd module on the B200 (tests/test_gpu_lvr_theta.py) and, formula by formula, on the CPU
 * (tests/test_theta_flow_formulas.py).  _bwd_ex: constant d/dlogq when d_g_logq is NULL; mask_grad = 1 gives TensorFlow's
 * masked_dense gradient (non-zero at masked kernel entries), nma_theta_flow_constrain re-applies the masks after an update. */
int nma_theta_flow_fwd(const float* d_params, const float* d_masks, const int32_t* d_perms, const float* d_z0,
                       int32_t p, int32_t d, int32_t nb, int32_t relu, float base_loc, float base_scale,
                       float* d_theta, float* d_logq, void* stream);
int nma_theta_flow_bwd(const float* d_params, const float* d_masks, const int32_t* d_perms, const float* d_z0,
                       int32_t p, int32_t d, int32_t nb, int32_t relu, const float* d_g_theta, const float* d_g_logq,
                       float* d_g_params, float* d_g_z0, void* stream);
int nma_theta_flow_bwd_ex(const float* d_params, const float* d_masks, const int32_t* d_perms, const float* d_z0,
                          int32_t p, int32_t d, int32_t nb, int32_t relu, const float* d_g_theta, const float* d_g_logq,
                          float g_logq_const, int32_t mask_grad, float* d_g_params, float* d_g_z0, void* stream);
int nma_theta_flow_constrain(float* d_flow_params, const float* d_masks, int32_t d, int32_t nb, void* stream);

/* A14 - rolling variances of the stochastic-volatility features (SV_dense.py:159-170):
 * d_var[i] = np.var(x[i : i+K]) for i in [0, n-K), on the float32 series, bit-exact with numpy's float32 np.var
 * (pairwise sums, two passes). */
int nma_rolling_var(const float* d_x, int64_t n, int32_t K, float* d_var, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NMA_B200_H */
