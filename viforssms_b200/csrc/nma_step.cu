// nma_train_step: everything the reference executes inside ONE `sess.run([self.train_step, self.merged], feed_dict)`
// (AR.py:300-301) as one C-ABI call, without a single library (ATen / cuBLAS) kernel in between:
//   base noise eps ~ N(0,1) [p, L0] and the theta posterior's base sample (AR.py:31-35, 117-118: sampled in-graph)
//     -> counter-based Philox4x32-10 + Box-Muller, keyed by (seed, device-resident draw counter): replayable in a CUDA graph
//   theta = theta_dist.sample(p), log q(theta)                      (AR.py:117-118, 376-391)  k_theta_flow_fwd
//   flow, ELBO terms, backward                                      (AR.py:44-110, 168-187)   step_forward_backward
//   + log prior(theta) - log q(theta), their gradient w.r.t. theta  (AR.py:178-185)           k_step_tail
//   backward through the theta posterior                            (AR.py:228-229)           k_theta_flow_bwd
//   gradient all-reduce over the time shards                                       nma_comm.cu
//   tf.global_norm, clip_by_global_norm, Adamax                     (AR.py:230-234)           nma_adamax_step
//   kernel_constraint of masked_dense (mask re-applied after the update)                      k_apply_mask
//   the logged scalars (AR.py:207-224)                                                        k_step_scalars
#include "nma_common.cuh"

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), the counter-based generator TensorFlow / cuRAND / PyTorch also use.
// counter = (element block lo, element block hi, draw counter lo, draw counter hi ^ stream id << 24), key = seed.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}

// two uniforms -> two standard normals (Box-Muller); u1 in (0, 1], u2 in [0, 1)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

__global__ void __launch_bounds__(256) k_philox_normal(float* __restrict__ out, int64_t n, unsigned long long seed,
                                                       const unsigned long long* __restrict__ d_counter,
                                                       unsigned long long counter_host, uint32_t stream_id, float loc,
                                                       float scale) {
    const unsigned long long ctr = d_counter ? d_counter[0] : counter_host;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int64_t nblk = (n + 3) >> 2;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (int64_t)gridDim.x * blockDim.x) {
        const uint4 c = make_uint4((uint32_t)b, (uint32_t)((unsigned long long)b >> 32), (uint32_t)ctr,
                                   (uint32_t)(ctr >> 32) ^ (stream_id << 24));
        const uint4 r = philox4x32_10(c, key);
        const float2 n0 = box_muller(r.x, r.y), n1 = box_muller(r.z, r.w);
        const float v[4] = {n0.x, n0.y, n1.x, n1.y};
        const int64_t e = b << 2;
        if (e + 3 < n && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
            *reinterpret_cast<float4*>(out + e) =
                make_float4(fmaf(scale, v[0], loc), fmaf(scale, v[1], loc), fmaf(scale, v[2], loc), fmaf(scale, v[3], loc));
        } else {
            for (int j = 0; j < 4; ++j)
                if (e + j < n) out[e + j] = fmaf(scale, v[j], loc);
        }
    }
}

int launch_philox_normal(float* d_out, int64_t n, unsigned long long seed, const unsigned long long* d_counter,
                         unsigned long long counter_host, uint32_t stream_id, float loc, float scale, cudaStream_t st) {
    if (!d_out || n < 1) { nma_set_error("philox: bad argument"); return -1; }
    int64_t blocks = ((n + 3) / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_philox_normal<<<(unsigned)blocks, 256, 0, st>>>(d_out, n, seed, d_counter, counter_host, stream_id, loc, scale);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_counter_bump(unsigned long long* c) {
    if (threadIdx.x == 0 && blockIdx.x == 0) c[0] += 1ull;
}
int launch_counter_bump(nma_handle_s* h, cudaStream_t st) {
    k_counter_bump<<<1, 32, 0, st>>>(h->step.counter);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Test / host hook: the normals nma_train_step (stream 1: theta base sample, scaled) and eps == NULL (stream 0) draw
// for a given (seed, counter).
extern "C" int nma_philox_normal(float* d_out, int64_t n, uint64_t seed, uint64_t counter, uint32_t stream_id, float loc,
                                 float scale, void* stream) {
    return launch_philox_normal(d_out, n, seed, nullptr, counter, stream_id, loc, scale, (cudaStream_t)stream);
}

extern "C" int nma_set_seed(nma_handle h, uint64_t seed, uint64_t counter) {
    if (!h) { nma_set_error("null handle"); return -1; }
    h->step.seed = seed;
    NMA_CHECK_CUDA(cudaMemcpy(h->step.counter, &counter, sizeof(counter), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int nma_get_counter(nma_handle h, uint64_t* out) {
    if (!h || !out) { nma_set_error("null argument"); return -1; }
    NMA_CHECK_CUDA(cudaMemcpy(out, h->step.counter, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------
// theta posterior registration
// ---------------------------------------------------------------------------
extern "C" int nma_set_theta_flow(nma_handle h, const float* d_masks, const int32_t* d_perms, int32_t nb, int32_t relu,
                                  float base_loc, float base_scale, const float* prior_mean, const float* prior_scale,
                                  int32_t softplus_out) {
    if (!h) { nma_set_error("null handle"); return -1; }
    // nb == 0: no posterior - theta is the constant prior_mean in every row (lotka_volterra_partial_batch_fix_theta.py:190)
    if (nb < 0 || nb > 8 || (nb > 0 && !d_masks) || (nb > 1 && !d_perms) || !prior_mean || !prior_scale || h->cfg.dtheta > 8) {
        nma_set_error("nma_set_theta_flow: bad argument");
        return -1;
    }
    h->step.tf_masks = d_masks; h->step.tf_perms = d_perms; h->step.tf_nb = nb; h->step.tf_relu = relu;
    h->step.tf_base_loc = base_loc; h->step.tf_base_scale = base_scale;
    h->step.tf_softplus = (softplus_out && nb > 0) ? 1 : 0;
    for (int k = 0; k < h->cfg.dtheta; ++k) {
        h->step.prior_mean[k] = prior_mean[k];
        h->step.prior_scale[k] = prior_scale[k];
        if (nb > 0 && !(prior_scale[k] > 0.f)) { nma_set_error("nma_set_theta_flow: prior scale must be positive"); return -1; }
    }
    h->step.tf_set = 1;
    return 0;
}

static int tf_layer_floats(int d) { return (d * 5 + 5) + 2 * (5 * 5 + 5) + (5 * 2 * d + 2 * d); }
extern "C" int64_t nma_theta_flow_param_count(int32_t d, int32_t nb) { return (int64_t)nb * tf_layer_floats(d); }

// ---------------------------------------------------------------------------
// per-row tail of the objective: log prior(theta) - log q(theta) and the upstream gradient of the theta posterior
// ---------------------------------------------------------------------------
struct TailArgs {
    const float* theta;        // [p][d]
    const float* logq_theta;   // [p]
    const float* terms;        // [p][4]
    const float* grad_theta;   // [p][d]   d(-sum_rows scale * (sde - logq + obs)) / d theta  (or the pre-train objective's)
    float* g_theta;            // [p][d]
    float* row_elbo;           // [p]
    const float* u;            // [p][d] pre-softplus sample when the posterior ends in a Softplus bijector (else null)
    int p, d, prior_on;
    float scale, obs_weight;
    float mean[8], sd[8];
};

__global__ void __launch_bounds__(128) k_step_tail(TailArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.p) return;
    float prior = 0.f;
    for (int k = 0; k < a.d; ++k) {
        if (a.u) {
            // theta = softplus(u); prior = TransformedDistribution(MultivariateNormalDiag, Softplus) (lotka_volterra_partial_batch.py
            // :358-365): log N(u; mean, sd) - log sigmoid(u).  The posterior carries the same Jacobian, so in prior - log q(theta)
            // it cancels: the flow receives d/du = d/dtheta * sigmoid(u) + (u - mean) / sd^2 and d/dlog q_u = 1.
            const float u = a.u[(size_t)r * a.d + k];
            const float sg = sigmoid_f(u);
            const float z = (u - a.mean[k]) / a.sd[k];
            prior += -0.5f * z * z - 0.5f * 1.8378770664093453f - logf(a.sd[k]) - logf(sg);
            a.g_theta[(size_t)r * a.d + k] = a.grad_theta[(size_t)r * a.d + k] * sg + (a.prior_on ? z / a.sd[k] : 0.f);
        } else {
            const float th = a.theta[(size_t)r * a.d + k];
            const float z = (th - a.mean[k]) / a.sd[k];
            // MultivariateNormalDiag(prior_mean, prior_scale).log_prob(theta)  (AR.py:178-182)
            prior += -0.5f * z * z - 0.5f * 1.8378770664093453f - logf(a.sd[k]);
            // objective = -sum_rows (... + prior - log q(theta)): d/dtheta gains +z/sd
            a.g_theta[(size_t)r * a.d + k] = a.grad_theta[(size_t)r * a.d + k] + (a.prior_on ? z / a.sd[k] : 0.f);
        }
    }
    const float* t = a.terms + (size_t)r * 4;
    a.row_elbo[r] = a.scale * (t[0] - t[2] + a.obs_weight * t[1]) + (a.prior_on ? prior - a.logq_theta[r] : 0.f);   // AR.py:184
}

// the scalars the reference logs every iteration (AR.py:207-224): means over the p rows
//   out[0] ELBO  [1] scale*sde  [2] log q(theta)  [3] scale*obs  [4] scale*logq(path)  [5] global norm
//   [6] number of rows with a non-finite term  [7] draw counter after this step
__global__ void __launch_bounds__(256) k_step_scalars(const float* __restrict__ row_elbo, const float* __restrict__ terms,
                                                      const float* __restrict__ logq_theta,
                                                      const uint32_t* __restrict__ flags, const float* __restrict__ norm,
                                                      unsigned long long* __restrict__ counter, int p, float scale,
                                                      float* __restrict__ out) {
    __shared__ float red[6][8];
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = threadIdx.x; r < p; r += blockDim.x) {
        acc[0] += row_elbo[r];
        acc[1] += terms[(size_t)r * 4 + 0];
        acc[2] += logq_theta[r];
        acc[3] += terms[(size_t)r * 4 + 1];
        acc[4] += terms[(size_t)r * 4 + 2];
        acc[5] += flags[r] ? 1.f : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float v = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t[6];
        for (int k = 0; k < 6; ++k) {
            t[k] = 0.f;
            for (int w = 0; w < 8; ++w) t[k] += red[k][w];
        }
        const float ip = 1.f / (float)p;
        out[0] = t[0] * ip; out[1] = scale * t[1] * ip; out[2] = t[2] * ip; out[3] = scale * t[3] * ip;
        out[4] = scale * t[4] * ip; out[5] = norm[0]; out[6] = t[5];
        counter[0] += 1ull;
        out[7] = (float)counter[0];
    }
}

// kernel_constraint of tf.contrib.distributions' masked_dense: after every update the kernels are multiplied by their
// block masks again (the gradient of a masked entry is NOT zero in TensorFlow - the mask is not part of the forward
// pass - it counts in tf.global_norm and its update is wiped here).  Layout of one layer: 4 x (kernel, bias).
__global__ void k_apply_mask(float* __restrict__ params, const float* __restrict__ masks, int d, int nb) {
    const int LP = (d * 5 + 5) + 2 * (5 * 5 + 5) + (5 * 2 * d + 2 * d);
    const int ksz[4] = {d * 5, 25, 25, 5 * 2 * d}, bsz[4] = {5, 5, 5, 2 * d};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nb * LP; t += gridDim.x * blockDim.x) {
        const int k = t / LP;
        int o = t - k * LP, mo = 0;
        for (int l = 0; l < 4; ++l) {
            if (o < ksz[l]) { params[t] *= masks[mo + o]; break; }
            o -= ksz[l]; mo += ksz[l];
            if (o < bsz[l]) break;
            o -= bsz[l];
        }
    }
}

extern "C" int nma_theta_flow_constrain(float* d_flow_params, const float* d_masks, int32_t d, int32_t nb, void* stream) {
    if (!d_flow_params || !d_masks || d < 1 || d > 8 || nb < 1) { nma_set_error("nma_theta_flow_constrain: bad argument"); return -1; }
    k_apply_mask<<<(nb * tf_layer_floats(d) + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_flow_params, d_masks, d, nb);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// the iteration
// ---------------------------------------------------------------------------
// Softplus bijector at the end of the posterior chain (lotka_volterra_partial_batch.py:741): theta = softplus(u),
// log q(theta) = log q_u(u) - sum log sigmoid(u)
__global__ void k_theta_softplus(const float* __restrict__ u, float* __restrict__ theta, float* __restrict__ logq, int p, int d) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p) return;
    float lj = 0.f;
    for (int k = 0; k < d; ++k) {
        const float v = u[(size_t)r * d + k];
        theta[(size_t)r * d + k] = softplus_f(v);
        lj += logf(sigmoid_f(v));
    }
    logq[r] -= lj;
}

struct FixedTheta { float* theta; float* logq; int p, d; float v[8]; };
__global__ void k_fixed_theta(FixedTheta a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.p) return;
    for (int k = 0; k < a.d; ++k) a.theta[(size_t)r * a.d + k] = a.v[k];
    a.logq[r] = 0.f;
}

extern "C" int nma_train_step(nma_handle h, float* d_blob, float* d_grad, float* d_m, float* d_v, const int64_t* d_idx,
                              int32_t p, const nma_step_opts* o, float* d_scalars, float* d_theta_out, float* d_lf_out,
                              void* stream) {
    if (!h || !d_blob || !d_grad || !d_m || !d_v || !d_idx || !o || !d_scalars) { nma_set_error("nma_train_step: null argument"); return -1; }
    if (p < 1 || p > h->cfg.p) { nma_set_error("p=%d outside 1..%d (nma_create sized the workspace)", p, h->cfg.p); return -1; }
    if (!h->base[0]) { nma_set_error("nma_set_series has not been called"); return -1; }
    if (!h->step.tf_set) { nma_set_error("nma_train_step: nma_set_theta_flow has not been called"); return -1; }
    if (o->objective < 0 || o->objective > 2) { nma_set_error("unknown objective %d", o->objective); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    StepWs& w = h->step;
    const int d = h->cfg.dtheta, nb = w.tf_nb;
    const int64_t n_flow = (int64_t)nb * tf_layer_floats(d);
    const int64_t n_total = h->n_params + n_flow;
    float* flow_params = d_blob + h->n_params;
    float* flow_grad = d_grad + h->n_params;
    int rc;
    // 1. noise: eps [p][L0] (stream 0) and the theta posterior's base sample [p][d] (stream 1)
    if ((rc = launch_philox_normal(w.eps, (int64_t)p * h->L0, w.seed, w.counter, 0, 0u, 0.f, 1.f, st))) return rc;
    // 2. theta ~ q(theta)  (or the constant of the fixed-theta script)
    if (nb > 0) {
        if ((rc = launch_philox_normal(w.z0, (int64_t)p * d, w.seed, w.counter, 0, 1u, w.tf_base_loc, w.tf_base_scale, st))) return rc;
        if ((rc = nma_theta_flow_fwd(flow_params, w.tf_masks, w.tf_perms, w.z0, p, d, nb, w.tf_relu, w.tf_base_loc,
                                     w.tf_base_scale, w.tf_softplus ? w.u : w.theta, w.logq_theta, stream)))
            return rc;
        if (w.tf_softplus) {
            k_theta_softplus<<<(p + 127) / 128, 128, 0, st>>>(w.u, w.theta, w.logq_theta, p, d);
            nma_count_launch(1);
        }
        NMA_CHECK_CUDA(cudaMemsetAsync(flow_grad, 0, (size_t)n_flow * 4, st));
    } else {
        FixedTheta f;
        f.theta = w.theta; f.logq = w.logq_theta; f.p = p; f.d = d;
        for (int k = 0; k < 8; ++k) f.v[k] = w.prior_mean[k];
        k_fixed_theta<<<(p + 127) / 128, 128, 0, st>>>(f);
        nma_count_launch(1);
    }
    // 3. flow, ELBO terms, backward down to d/dtheta (per-flow all-reduce on the side stream when sharded)
    if ((rc = step_forward_backward(h, d_blob, w.eps, w.theta, d_idx, p, o->objective, o->path_target, w.terms, d_lf_out,
                                    d_grad, w.g_theta, w.flags, true, st, true)))
        return rc;
    // 4. prior / entropy of theta and the gradient entering the theta posterior
    {
        TailArgs a;
        a.theta = w.theta; a.logq_theta = w.logq_theta; a.terms = w.terms; a.grad_theta = w.g_theta; a.g_theta = w.g_theta;
        a.u = (w.tf_softplus && nb > 0) ? w.u : nullptr;
        a.row_elbo = w.row_elbo; a.p = p; a.d = d; a.prior_on = (o->prior_on && nb > 0) ? 1 : 0; a.scale = (float)h->cfg.scale;
        a.obs_weight = o->obs_in_elbo ? 1.f : 0.f;
        for (int k = 0; k < 8; ++k) { a.mean[k] = w.prior_mean[k]; a.sd[k] = k < d ? w.prior_scale[k] : 1.f; }
        k_step_tail<<<(p + 127) / 128, 128, 0, st>>>(a);
        nma_count_launch(1);
    }
    // 5. backward through the theta posterior: d/dlogq(theta) = +1 when the entropy term is part of the objective
    if (nb > 0 && (rc = nma_theta_flow_bwd_ex(flow_params, w.tf_masks, w.tf_perms, w.z0, p, d, nb, w.tf_relu, w.g_theta,
                                              nullptr, o->prior_on ? 1.f : 0.f, o->tf_mask_grad, flow_grad, nullptr, stream)))
        return rc;
    // (flow 0's conv / feature backward may still be running on the handle's second stream: small launches only)
    if ((rc = step_aux_join(h, st))) return rc;
    // 6. the theta posterior's section of the gradient, then wait for every collective of this step
    if (h->comm.comm) {
        if (nb > 0 && (rc = comm_allreduce_after(h, flow_grad, n_flow, h->cfg.F, st))) return rc;
        if ((rc = comm_join(h, st))) return rc;
    }
    // 7. clip + Adamax over every variable (one global norm, AR.py:228-234), then masked_dense's kernel constraint
    if ((rc = nma_adamax_step(d_blob, d_grad, d_m, d_v, n_total, o->lr, o->beta1, o->beta2, o->eps, o->clip, w.norm,
                              w.norm + 16, stream)))
        return rc;
    if (o->tf_mask_grad && nb > 0) {
        if ((rc = nma_theta_flow_constrain(flow_params, w.tf_masks, d, nb, stream))) return rc;
    }
    // 8. logged scalars; bumps the draw counter
    k_step_scalars<<<1, 256, 0, st>>>(w.row_elbo, w.terms, w.logq_theta, w.flags, w.norm, w.counter, p, (float)h->cfg.scale,
                                      d_scalars);
    nma_count_launch(1);
    if (d_theta_out) NMA_CHECK_CUDA(cudaMemcpyAsync(d_theta_out, w.theta, (size_t)p * d * 4, cudaMemcpyDeviceToDevice, st));
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// views of what the last nma_train_step left on the device (borrowed; valid until nma_destroy)
extern "C" int nma_step_buffers(nma_handle h, float** d_eps, float** d_z0, float** d_theta, float** d_logq_theta,
                                float** d_terms, float** d_row_elbo) {
    if (!h) { nma_set_error("null handle"); return -1; }
    if (d_eps) *d_eps = h->step.eps;
    if (d_z0) *d_z0 = h->step.z0;
    if (d_theta) *d_theta = h->step.theta;
    if (d_logq_theta) *d_logq_theta = h->step.logq_theta;
    if (d_terms) *d_terms = h->step.terms;
    if (d_row_elbo) *d_row_elbo = h->step.row_elbo;
    return 0;
}
