// Internal declarations shared by the NMA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nma_b200.h"

#define NMA_C   50          // network_dims[0] in every reference script
#define NMA_C1  51          // conv input channels: previous sample + C features (AR.py:58-59)
#define NMA_MAXH 4

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "this library is written for sm_100a (B200) only"
#endif

struct FlowDims {
    int L;      // length of the flow's input sample x^(i)         (AR.py:132: L0 - i*K)
    int Lin;    // conv input positions = L-1                      (AR.py:53,58: [:, :-1])
    int N;      // conv output positions = L-K = length of x^(i+1) (AR.py:61-62 'valid')
    int LP;     // pitch (floats) of [c][Lin] tensors
    int NP;     // pitch of [c][N] tensors
};

// offsets (floats) into the flat parameter blob, TF creation order (AR.py:53-78)
struct FlowParamOff {
    int64_t featw[4], featb[4];
    int64_t convw, convb;
    int64_t thw[3], thb[3];
    int64_t hidw[NMA_MAXH], hidb[NMA_MAXH], gam[NMA_MAXH], bet[NMA_MAXH];
    int64_t headw, headb;
};

struct FlowWs {
    float* x;        // [p][L]      input sample of this flow (x^(0) = eps); index F = final sample
    float* dx;       // [p][L]      d objective / d x^(i)
    float* a[5];     // a[0]: [p][Cf_in][LP] gathered features; a[1..4]: [p][C][LP] feature-MLP activations
    float* h[NMA_MAXH + 1];  // [p][C][NP] post-ELU activations e_0..e_H of the conv head
    float* s;        // [p][NP]     pre-softplus scale logits
    float* dA;       // [p][C][NP]  gradient w.r.t. the conv pre-activation
    float* df;       // [p][C][LP]  gradient w.r.t. the feature channels of the conv input
    float* df3;      // LV only: [p][C][LWP] gradient w.r.t. the third feature activation (input of the wide 4th layer)
    float* tb;       // [p][3][C]   theta-MLP activations t1, t2, b(+conv bias)
    float* dtb;      // [p][C]      sum_m dA
    float* wpk;      // packed conv weights for the forward conv  [Cin=51][5][KP][12]
    float* wdpk;     // packed conv weights for the data-gradient conv [Cin=50][6][KP][12]
    // ---- tensor-core path (nma_tc.cuh): flattened interleaved operands, q = row*Lin + slot ----
    // (bf16 mode: the same buffers hold [8][Q][8 x bf16] units, see TcP<true>)
    float* tin_hi;   // [14][tin_Q][4]  conv input (channel 0 = x^(i), 1..50 = feature activations), 3xTF32 hi part
    float* tin_lo;   //                 lo part
    float* dat_hi;   // [14][dat_Q][4]  dA (gradient w.r.t. the conv pre-activation) at q = K-1 + row*Lin + m
    float* dat_lo;
    float* wtc_f;    // [K][14][64 hi | 64 lo rows][4] packed taps of the forward conv
    float* wtc_d;    // same, data-gradient conv (flipped, transposed)
    float* wtc_feat; // [10][14][64 hi | 64 lo rows][4] packed feature-MLP kernels: 4 forward, 4 transposed, + the hidden 1x1 kernel transposed / forward
    long long tin_Q, dat_Q;
};

// nma_train_step (nma_step.cu): the theta posterior registered with nma_set_theta_flow and the per-row buffers of the
// part of sess.run(train_step) that surrounds nma_elbo_fwd_bwd (AR.py:117-118,178-185,226-234)
struct StepWs {
    float* eps;              // [p][L0]      base noise drawn in the library (eps == NULL)
    float* z0;               // [p][8]       base sample of the theta posterior
    float* theta;            // [p][dtheta]
    float* u;                // [p][8]       pre-softplus sample (posteriors that end in a Softplus bijector)
    float* logq_theta;       // [p]
    float* g_theta;          // [p][dtheta]  d objective / d theta (ELBO part + prior part)
    float* terms;            // [p][4]
    float* row_elbo;         // [p]
    uint32_t* flags;         // [p]
    float* norm;             // [1] global norm, [1..1024] partial sums of k_sumsq
    unsigned long long* counter;   // [2]  device-resident draw counter (Philox offset), bumped by the last kernel of a call
    unsigned long long seed;
    // theta posterior (borrowed device pointers)
    const float* tf_masks;
    const int32_t* tf_perms;
    int tf_nb, tf_relu, tf_set, tf_softplus;
    float tf_base_loc, tf_base_scale;
    float prior_mean[8], prior_scale[8];
};

// NCCL communicator used for the gradient all-reduce (nma_comm.cu); the library resolves libnccl.so.2 at run time
struct CommState {
    void* comm;              // ncclComm_t
    int owned;               // created by nma_comm_create (destroyed by nma_comm_destroy / nma_destroy)
    int world, rank;
    cudaStream_t side;       // the collective's stream
    cudaEvent_t ev_ready[NMA_MAX_FLOWS + 2];   // a gradient section is complete on the compute stream
    cudaEvent_t ev_done;     // all collectives of the step have been issued and finished on `side`
    int pending;             // collectives issued since the last join
};

struct nma_handle_s {
    nma_config cfg;
    int L0, S, Cf_in, feat_off, KP;
    FlowDims fd[NMA_MAX_FLOWS + 1];
    FlowParamOff po[NMA_MAX_FLOWS];
    int64_t n_params;
    FlowWs ws[NMA_MAX_FLOWS + 1];
    const float* base[NMA_MAX_ARRAYS];
    int64_t base_len[NMA_MAX_ARRAYS];
    void* arena;
    int64_t arena_bytes;
    int sm_count;
    int dev;
    int tc_ok;       // the tensor-core conv supports this configuration
    int use_tc;      // ... and is switched on (default; NMA_TC=0 or nma_set_tensor_cores(h, 0) selects the FP32 SIMT conv)
    int tc_nacc;     // 128-position accumulators per CTA (2 when the tile fits in shared memory, else 1)
    int use_tc_persist;  // persistent warp-specialised conv kernels (nma_tc_conv2.cu); NMA_TC_PERSIST=0 keeps one tile per CTA
    int use_tc_feat; // feature MLP on the tensor cores as well (needs use_tc; NMA_TC_FEAT=0 keeps the FP32 SIMT kernels)
    int bf16_ok;     // the bf16-split conv path covers this configuration (persistent conv kernels + tensor-core head backward)
    int dgrad_wide;  // bf16 split only: data gradient as N = 256 instructions with the stacked kernel as A (k_conv_dgrad_tcq);
                     // (default) NMA_DGRAD_WIDE=0 selects the M = 128 positions form k_conv_dgrad_tcp<true>
    int tap_pairs;   // bf16 split only: forward / data-gradient taps in pairs, 7 instead of 8 instructions per pair
                     // (nma_tc_conv2.cu); default for even kernel_len, NMA_TAP_PAIRS=0 / 1 forces it off / on; needs dgrad_wide
    int use_bf16;    // conv GEMMs (forward, data gradient, weight gradient) in the 2-term bf16 split on kind::f16
                     // (nma_tc.cuh); NMA_TC_BF16=1 or bit 2 of nma_set_tensor_cores
    // Lotka-Volterra (lotka_volterra_partial_batch_fix_theta.py:71-82): every flow's feature MLP runs over the whole
    // window (LW = L0 - 1 positions), ends in a dense layer as wide as the flow's conv input and is transposed, so the
    // conv has 1 + LW input channels.  For the other models conv_cin = 51 and feat_out[i] = 50.
    int is_lv, conv_cin, LW, LWP;
    int feat_out[NMA_MAX_FLOWS];
    // channel-split scratch of the SIMT conv kernels at small row counts (nma_conv_core.cuh: conv_split_reduce)
    float* split_part;
    unsigned* split_ticket;
    // second stream for the kernels of a step that do not depend on each other (weight packing next to the feature
    // forward; the conv weight gradient next to data gradient + feature backward): used when the launch does not fill the
    // machine (the scripts' own row counts), joined back with events - also inside a captured graph
    cudaStream_t aux, aux2;  // aux: weight packing, conv weight gradients; aux2: feature backward (+ the last flow's data gradient)
    cudaEvent_t ev_fork, ev_join, ev_join2;
    int aux_pending;         // step_forward_backward left flow 0's conv / feature backward running on `aux` (step_aux_join)
    // ---- whole-iteration entry point (nma_step.cu: nma_train_step) ----
    StepWs step;
    // ---- gradient all-reduce inside the library (nma_comm.cu) ----
    CommState comm;
};

// device-side copy of what kernels need about the series and the channel table
struct SeriesView {
    const float* base[NMA_MAX_ARRAYS];
    long long len[NMA_MAX_ARRAYS];
    int chan_array[NMA_MAX_CHAN];
    int chan_offset[NMA_MAX_CHAN];
    int Cf, D, feat_aug;
};

void nma_set_error(const char* fmt, ...);
void nma_count_launch(int n);   // kernels launched by this library (bench.py reports it)
#define NMA_CHECK_CUDA(call)                                                           \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            nma_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return -2;                                                                 \
        }                                                                              \
    } while (0)

SeriesView nma_series_view(const nma_handle_s* h);

// launchers (one per kernel family); all asynchronous on `st`
int launch_gather(nma_handle_s* h, const int64_t* idx, int p, float* tf, float* mask, float* shift, cudaStream_t st);
// which: 1 = conv tap / 1x1 kernels, 2 = feature kernels (tensor-core feature path), 3 = both
int launch_pack_weights(nma_handle_s* h, const float* params, bool need_bwd, cudaStream_t st, int which = 3);
int launch_theta_fwd(nma_handle_s* h, const float* params, const float* theta, int p, cudaStream_t st);
int launch_feat_fwd(nma_handle_s* h, const float* params, const int64_t* idx, int p, bool save, cudaStream_t st);
int launch_conv_fwd(nma_handle_s* h, int flow, const float* params, int p, bool save, cudaStream_t st);
int launch_conv_fwd_tc(nma_handle_s* h, int flow, const float* params, int p, bool save, cudaStream_t st);
int launch_conv_dgrad_tc(nma_handle_s* h, int flow, int p, cudaStream_t st);
int launch_conv_wgrad_tc(nma_handle_s* h, int flow, int p, float* grad_params, cudaStream_t st);
int launch_conv_dgrad_tcp(nma_handle_s* h, int flow, int p, cudaStream_t st);
int launch_conv_wgrad_bf(nma_handle_s* h, int flow, int p, float* grad_params, cudaStream_t st);
int conv_fwd_tcp_supported(const nma_handle_s* h);
int launch_conv_fwd_tcp(nma_handle_s* h, int flow, const float* params, int p, bool save, cudaStream_t st);
int launch_pack_weights_tc(nma_handle_s* h, const float* params, bool need_bwd, cudaStream_t st, int which = 3);
int launch_pack_w1x1_fwd(nma_handle_s* h, int i, const float* params, cudaStream_t st);
int launch_pack_w1x1_t(nma_handle_s* h, int i, const float* params, cudaStream_t st);
int launch_pack_feat_tc(nma_handle_s* h, const float* params, bool need_bwd, cudaStream_t st);
int epi_bwd_tc_supported(const nma_handle_s* h);
int launch_epi_bwd_tc(nma_handle_s* h, int flow, const float* params, int p, int objective, float* grad_params,
                      cudaStream_t st);
int launch_feat_bwd_tc(nma_handle_s* h, int flow, const float* params, int p, float* grad_params, cudaStream_t st);
int launch_feat_fwd_tc(nma_handle_s* h, const float* params, const int64_t* idx, const float* eps, int p, bool save,
                       cudaStream_t st);
struct FlowEpiArgs;
void fill_flow_epi_args(nma_handle_s* h, int flow, const float* params, bool save, FlowEpiArgs& e);
int launch_elbo(nma_handle_s* h, const float* theta, const float* eps, const int64_t* idx, int p, int objective,
                float path_target, float* terms, float* lf, float* grad_theta, uint32_t* flags, bool want_grad,
                cudaStream_t st);
int launch_epi_bwd(nma_handle_s* h, int flow, const float* params, int p, int objective, float* grad_params,
                   cudaStream_t st);
int launch_conv_dgrad(nma_handle_s* h, int flow, int p, cudaStream_t st);
int launch_conv_wgrad(nma_handle_s* h, int flow, int p, float* grad_params, cudaStream_t st);
int launch_feat_bwd(nma_handle_s* h, int flow, const float* params, int p, float* grad_params, cudaStream_t st);
int launch_theta_bwd(nma_handle_s* h, const float* params, const float* theta, int p, float* grad_params,
                     float* grad_theta, int flow, cudaStream_t st);   // flow < 0: every flow in one launch
// nma_comm.cu: all-reduce(sum) of grad[off, off+count) on the side stream once `st` has reached this point; no-op without
// a communicator.  comm_join makes `st` wait for everything issued so far.
int comm_allreduce_after(nma_handle_s* h, float* d_buf, int64_t count, int slot, cudaStream_t st);
int comm_join(nma_handle_s* h, cudaStream_t st);
void comm_release(nma_handle_s* h);
// nma_step.cu
int launch_philox_normal(float* d_out, int64_t n, unsigned long long seed, const unsigned long long* d_counter,
                         unsigned long long counter_host, uint32_t stream_id, float loc, float scale, cudaStream_t st);
int launch_counter_bump(nma_handle_s* h, cudaStream_t st);
int step_forward_backward(nma_handle_s* h, const float* d_params, const float* d_eps, const float* d_theta,
                          const int64_t* d_idx, int p, int objective, float path_target, float* d_terms, float* d_lf,
                          float* d_grad_params, float* d_grad_theta, uint32_t* d_flags, bool per_flow_collective,
                          cudaStream_t st, bool defer_last_join = false);
// waits (on st) for what step_forward_backward(defer_last_join = true) left running on the handle's second stream
int step_aux_join(nma_handle_s* h, cudaStream_t st);
// Lotka-Volterra instances (nma_lv.cu)
int launch_lv_feat_fwd(nma_handle_s* h, const float* params, const int64_t* idx, const float* eps, int p, bool save,
                       cudaStream_t st);
int launch_lv_conv_dgrad(nma_handle_s* h, int flow, const float* params, int p, cudaStream_t st);
int launch_lv_conv_wgrad(nma_handle_s* h, int flow, int p, float* grad_params, cudaStream_t st);
int launch_lv_feat4_bwd(nma_handle_s* h, int flow, const float* params, int p, float* grad_params, cudaStream_t st);

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
// ELU as exp(z) - 1 through ex2.approx (one MUFU) and a select: no divergent branch per element, ~6 instructions
// instead of expm1f's ~30.  Absolute error <= ~2e-7 on outputs in (-1, 0]; expm1f's relative accuracy near 0 is
// irrelevant downstream (the value is added to O(1) sums).  The fused feature kernel went from 314 M to 220 M warp
// instructions per 2048 rows with this change alone (profiles/r01_feat_tc.md).
// ex2.approx.ftz directly: __expf wraps the MUFU in a denormal-range guard (FSETP + two predicated FMULs) that cannot
// matter here - an argument below -126 gives exp - 1 == -1 either way.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float elu_f(float z) {
    const float e = ex2_ftz(fminf(z, 0.f) * 1.4426950408889634f) - 1.f;
    return z > 0.f ? z : e;
}
// derivative of ELU expressed through its output e = elu(z): z>0 -> 1, else e+1
__device__ __forceinline__ float elu_grad_from_out(float e) { return e > 0.f ? 1.f : e + 1.f; }
// tf.nn.softplus, numerically stable: max(x,0) + log1p(exp(-|x|))
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// series access (A1): time_feats[r, slot, c] = base[chan_array[c]][win0 + slot + chan_offset[c]]
__device__ __forceinline__ float series_val(const SeriesView& sv, int c, long long pos) {
    const int a = sv.chan_array[c];
    const long long q = pos + sv.chan_offset[c];
    return (q >= 0 && q < sv.len[a]) ? __ldg(sv.base[a] + q) : 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sums of 16 per-lane values over the 32 lanes of a warp in 16 shuffles instead of 16 x 5: every step halves the number of
// values a lane still carries (it keeps the half selected by one bit of its lane index and sends the other half to its
// partner).  Returns, in EVERY lane, the full sum of value number  j = 8*bit4 + 4*bit3 + 2*bit2 + bit1  of the lane index
// (the two lanes of a pair hold the same value); *j_out receives j.
__device__ __forceinline__ float warp_sum16_transposed(const float (&v)[16], int lane, int* j_out) {
    float w[8], x[4], y[2], z;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b3 ? w[i] : w[i + 4], keep = b3 ? w[i + 4] : w[i];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b2 ? x[i] : x[i + 2], keep = b2 ? x[i + 2] : x[i];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const float send = b1 ? y[0] : y[1], keep = b1 ? y[1] : y[0];
        z = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    z += __shfl_xor_sync(0xffffffffu, z, 1);
    *j_out = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
    return z;
}
