// A11 on the device: the variational posterior over theta - `num_bijectors` inverse masked-autoregressive-flow layers
// with fixed permutations in between (AR.py:376-391; fitz_nag_NVP.py:480-494; SV_dense.py:428-442;
// masked_autoregressive_default_template(hidden_layers=[5,5,5])) - as two kernels, one thread per row:
//   k_theta_flow_fwd : z0 [p][d] -> theta [p][d], log q(theta) [p]
//   k_theta_flow_bwd : dL/dtheta [p][d], dL/dlogq [p] -> gradient w.r.t. the ~580 flow parameters (summed over rows)
// The host-side autograd module (viforssms_b200/theta_flow.py) costs ~400 tiny launches per step, which is most of
// a p = 50 step; these two launches replace them.
//
// The arithmetic below is replayed statement by statement in float64 against autograd of theta_flow.py
// (tests/test_theta_flow_formulas.py); on the B200 the kernels agree with the host module to 2.6e-7 on the parameter
// gradients (tests/test_gpu_lvr_theta.py).  nma_train_step (nma_step.cu) is their caller.
//
// mask_grad: TensorFlow's masked_dense keeps masked kernel entries at zero through the initialiser and a
// kernel_constraint applied after every update, but does not multiply the mask in the forward pass, so the gradient it
// hands to tf.global_norm (AR.py:230) is NOT zero at masked entries.  mask_grad = 1 reproduces that (the caller re-applies
// the mask after the update, nma_theta_flow_constrain); mask_grad = 0 gives the gradient of the masked forward pass.
//
// Per layer k, with the masked MLP  out = W3.(act(W2.(act(W1.(act(W0.z + b0)) + b1)) + b2)) + b3  (W_i = kernel * mask,
// sizes d -> 5 -> 5 -> 5 -> 2d):  shift_j = out[2j], ls_j = clip(out[2j+1], -5, 3) (value clipped, gradient passed
// through: clip_by_value_preserve_gradient),  z'_j = (z_j - shift_j) exp(-ls_j),  log q += sum_j ls_j,  then z' is
// permuted (z''_j = z'_{perm[j]}) except after the last layer.
#include "nma_common.cuh"

#define TF_H 5          // hidden width of masked_autoregressive_default_template(hidden_layers=[5,5,5])
#define TF_DMAX 8       // theta dimension (3..5 in the reference scripts)
#define TF_NBMAX 8      // bijector layers (4 or 5 in the reference scripts)
#define TF_LOG2PI 1.8378770664093453f

struct ThetaFlowArgs {
    const float* params;       // [nb][ (d*5+5) + (5*5+5) + (5*5+5) + (5*2d+2d) ]  kernels row-major [in][out], then bias
    const float* masks;        // [d*5 + 25 + 25 + 5*2d]  the four block masks of one layer (shared by all layers)
    const float* z0;           // [p][d]
    const int* perms;          // [nb-1][d]
    float* theta;              // [p][d]
    float* logq;               // [p]
    const float* g_theta;      // [p][d]   (bwd)
    const float* g_logq;       // [p]      (bwd; may be null: 0)
    float* g_params;           // [n]      (bwd; accumulated into)
    float* g_z0;               // [p][d]   (bwd; may be null)
    int p, d, nb, relu;
    float base_loc, base_scale;
    float g_logq_const;        // (bwd) used when g_logq is null
    int mask_grad;             // (bwd) 1: TensorFlow's masked_dense gradient (masked entries receive h * g as well)
};

__device__ __forceinline__ float tf_act(float a, int relu) { return relu ? fmaxf(a, 0.f) : elu_f(a); }
// derivative of the activation expressed through its output
__device__ __forceinline__ float tf_dact(float h, int relu) { return relu ? (h > 0.f ? 1.f : 0.f) : (h > 0.f ? 1.f : h + 1.f); }

__host__ __device__ inline int tf_layer_params(int d) { return (d * TF_H + TF_H) + 2 * (TF_H * TF_H + TF_H) + (TF_H * 2 * d + 2 * d); }

// masked MLP of one layer: h0 = z [d], h1..h3 [5] post-activation, out [2d]
__device__ __forceinline__ void tf_mlp(const float* __restrict__ P, const float* __restrict__ M, int d, int relu,
                                       const float* z, float (&h1)[TF_H], float (&h2)[TF_H], float (&h3)[TF_H],
                                       float (&out)[2 * TF_DMAX]) {
    const float* W0 = P;                       const float* b0 = W0 + d * TF_H;
    const float* W1 = b0 + TF_H;               const float* b1 = W1 + TF_H * TF_H;
    const float* W2 = b1 + TF_H;               const float* b2 = W2 + TF_H * TF_H;
    const float* W3 = b2 + TF_H;               const float* b3 = W3 + TF_H * 2 * d;
    const float* M0 = M; const float* M1 = M0 + d * TF_H; const float* M2 = M1 + TF_H * TF_H; const float* M3 = M2 + TF_H * TF_H;
#pragma unroll
    for (int o = 0; o < TF_H; ++o) {
        float a = __ldg(b0 + o);
        for (int i = 0; i < d; ++i) a = fmaf(z[i], __ldg(W0 + i * TF_H + o) * __ldg(M0 + i * TF_H + o), a);
        h1[o] = tf_act(a, relu);
    }
#pragma unroll
    for (int o = 0; o < TF_H; ++o) {
        float a = __ldg(b1 + o);
#pragma unroll
        for (int i = 0; i < TF_H; ++i) a = fmaf(h1[i], __ldg(W1 + i * TF_H + o) * __ldg(M1 + i * TF_H + o), a);
        h2[o] = tf_act(a, relu);
    }
#pragma unroll
    for (int o = 0; o < TF_H; ++o) {
        float a = __ldg(b2 + o);
#pragma unroll
        for (int i = 0; i < TF_H; ++i) a = fmaf(h2[i], __ldg(W2 + i * TF_H + o) * __ldg(M2 + i * TF_H + o), a);
        h3[o] = tf_act(a, relu);
    }
    for (int o = 0; o < 2 * d; ++o) {
        float a = __ldg(b3 + o);
#pragma unroll
        for (int i = 0; i < TF_H; ++i) a = fmaf(h3[i], __ldg(W3 + i * 2 * d + o) * __ldg(M3 + i * 2 * d + o), a);
        out[o] = a;
    }
}

__global__ void __launch_bounds__(128) k_theta_flow_fwd(ThetaFlowArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.p) return;
    const int d = a.d, LP = tf_layer_params(d);
    float z[TF_DMAX], zn[TF_DMAX], h1[TF_H], h2[TF_H], h3[TF_H], out[2 * TF_DMAX];
    float lp = 0.f;
    for (int j = 0; j < d; ++j) {
        z[j] = a.z0[(size_t)r * d + j];
        const float q = (z[j] - a.base_loc) / a.base_scale;
        lp += -0.5f * q * q - 0.5f * TF_LOG2PI - logf(a.base_scale);
    }
    for (int k = 0; k < a.nb; ++k) {
        tf_mlp(a.params + (size_t)k * LP, a.masks, d, a.relu, z, h1, h2, h3, out);
        for (int j = 0; j < d; ++j) {
            const float ls = fminf(fmaxf(out[2 * j + 1], -5.f), 3.f);
            zn[j] = (z[j] - out[2 * j]) * expf(-ls);
            lp += ls;
        }
        for (int j = 0; j < d; ++j) z[j] = (k < a.nb - 1) ? zn[a.perms[k * d + j]] : zn[j];
    }
    for (int j = 0; j < d; ++j) a.theta[(size_t)r * d + j] = z[j];
    a.logq[r] = lp;
}

__global__ void __launch_bounds__(128) k_theta_flow_bwd(ThetaFlowArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = r < a.p;            // every lane walks the whole kernel: the parameter gradients are warp sums
    const bool mg = a.mask_grad != 0;
    const int d = a.d, LP = tf_layer_params(d);
    float zin[TF_NBMAX][TF_DMAX];          // input of every layer (recomputed forward)
    float z[TF_DMAX], zn[TF_DMAX], h1[TF_H], h2[TF_H], h3[TF_H], out[2 * TF_DMAX];
    for (int j = 0; j < d; ++j) z[j] = valid ? a.z0[(size_t)r * d + j] : 0.f;
    for (int k = 0; k < a.nb; ++k) {
        for (int j = 0; j < d; ++j) zin[k][j] = z[j];
        tf_mlp(a.params + (size_t)k * LP, a.masks, d, a.relu, z, h1, h2, h3, out);
        for (int j = 0; j < d; ++j) {
            const float ls = fminf(fmaxf(out[2 * j + 1], -5.f), 3.f);
            zn[j] = (z[j] - out[2 * j]) * expf(-ls);
        }
        for (int j = 0; j < d; ++j) z[j] = (k < a.nb - 1) ? zn[a.perms[k * d + j]] : zn[j];
    }
    const float glp = valid ? (a.g_logq ? a.g_logq[r] : a.g_logq_const) : 0.f;
    float gz[TF_DMAX], gzn[TF_DMAX];       // gradient w.r.t. the layer's (permuted) output, then w.r.t. its input
    for (int j = 0; j < d; ++j) gz[j] = valid ? a.g_theta[(size_t)r * d + j] : 0.f;
    for (int k = a.nb - 1; k >= 0; --k) {
        const float* P = a.params + (size_t)k * LP;
        float* G = a.g_params + (size_t)k * LP;
        const float* M = a.masks;
        // un-permute: z''_j = z'_{perm[j]}
        if (k < a.nb - 1) {
            for (int j = 0; j < d; ++j) gzn[j] = 0.f;
            for (int j = 0; j < d; ++j) {
                const int src = a.perms[k * d + j];
                for (int i = 0; i < d; ++i) gzn[i] += (i == src) ? gz[j] : 0.f;       // (no dynamic register indexing)
            }
        } else {
            for (int j = 0; j < d; ++j) gzn[j] = gz[j];
        }
        for (int j = 0; j < d; ++j) z[j] = zin[k][j];
        tf_mlp(P, M, d, a.relu, z, h1, h2, h3, out);
        // z'_j = (z_j - s_j) exp(-ls_j);  log q += ls_j
        float gout[2 * TF_DMAX];
        for (int j = 0; j < d; ++j) {
            const float ls = fminf(fmaxf(out[2 * j + 1], -5.f), 3.f);
            const float e = expf(-ls);
            gz[j] = gzn[j] * e;                                         // direct path to the layer input
            gout[2 * j] = -gzn[j] * e;
            gout[2 * j + 1] = -gzn[j] * (z[j] - out[2 * j]) * e + glp;   // clip: value only, gradient passes
        }
        // ---- back through the masked MLP ----
        const float* W0 = P;                       const float* W1 = W0 + d * TF_H + TF_H;
        const float* W2 = W1 + TF_H * TF_H + TF_H; const float* W3 = W2 + TF_H * TF_H + TF_H;
        float* G0 = G;                             float* G1 = G0 + d * TF_H + TF_H;
        float* G2 = G1 + TF_H * TF_H + TF_H;       float* G3 = G2 + TF_H * TF_H + TF_H;
        const float* M0 = M; const float* M1 = M0 + d * TF_H; const float* M2 = M1 + TF_H * TF_H; const float* M3 = M2 + TF_H * TF_H;
        float g3[TF_H], g2[TF_H], g1[TF_H];
        // layer 3 (no activation): out = h3 W3 + b3
#pragma unroll
        for (int i = 0; i < TF_H; ++i) g3[i] = 0.f;
        for (int o = 0; o < 2 * d; ++o) {
            const float go = gout[o];
            const float sb = warp_sum(go);
            if (lane == 0) atomicAdd(G3 + TF_H * 2 * d + o, sb);
#pragma unroll
            for (int i = 0; i < TF_H; ++i) {
                const float m = __ldg(M3 + i * 2 * d + o);
                const float sw = warp_sum(h3[i] * go) * (mg ? 1.f : m);
                if (lane == 0 && (mg || m != 0.f)) atomicAdd(G3 + i * 2 * d + o, sw);
                g3[i] = fmaf(go, __ldg(W3 + i * 2 * d + o) * m, g3[i]);
            }
        }
        // layer 2: h3 = act(h2 W2 + b2)
#pragma unroll
        for (int i = 0; i < TF_H; ++i) g2[i] = 0.f;
#pragma unroll
        for (int o = 0; o < TF_H; ++o) {
            const float go = g3[o] * tf_dact(h3[o], a.relu);
            const float sb = warp_sum(go);
            if (lane == 0) atomicAdd(G2 + TF_H * TF_H + o, sb);
#pragma unroll
            for (int i = 0; i < TF_H; ++i) {
                const float m = __ldg(M2 + i * TF_H + o);
                const float sw = warp_sum(h2[i] * go) * (mg ? 1.f : m);
                if (lane == 0 && (mg || m != 0.f)) atomicAdd(G2 + i * TF_H + o, sw);
                g2[i] = fmaf(go, __ldg(W2 + i * TF_H + o) * m, g2[i]);
            }
        }
        // layer 1: h2 = act(h1 W1 + b1)
#pragma unroll
        for (int i = 0; i < TF_H; ++i) g1[i] = 0.f;
#pragma unroll
        for (int o = 0; o < TF_H; ++o) {
            const float go = g2[o] * tf_dact(h2[o], a.relu);
            const float sb = warp_sum(go);
            if (lane == 0) atomicAdd(G1 + TF_H * TF_H + o, sb);
#pragma unroll
            for (int i = 0; i < TF_H; ++i) {
                const float m = __ldg(M1 + i * TF_H + o);
                const float sw = warp_sum(h1[i] * go) * (mg ? 1.f : m);
                if (lane == 0 && (mg || m != 0.f)) atomicAdd(G1 + i * TF_H + o, sw);
                g1[i] = fmaf(go, __ldg(W1 + i * TF_H + o) * m, g1[i]);
            }
        }
        // layer 0: h1 = act(z W0 + b0)
#pragma unroll
        for (int o = 0; o < TF_H; ++o) {
            const float go = g1[o] * tf_dact(h1[o], a.relu);
            const float sb = warp_sum(go);
            if (lane == 0) atomicAdd(G0 + d * TF_H + o, sb);
            for (int i = 0; i < d; ++i) {
                const float m = __ldg(M0 + i * TF_H + o);
                const float sw = warp_sum(z[i] * go) * (mg ? 1.f : m);
                if (lane == 0 && (mg || m != 0.f)) atomicAdd(G0 + i * TF_H + o, sw);
                gz[i] = fmaf(go, __ldg(W0 + i * TF_H + o) * m, gz[i]);
            }
        }
    }
    if (valid && a.g_z0)
        for (int j = 0; j < d; ++j) a.g_z0[(size_t)r * d + j] = gz[j];
}

// ---------------------------------------------------------------------------
// The same backward pass specialised on the theta dimension (3, 4, 5 in the reference scripts): every per-row array is
// statically indexed (registers instead of the generic kernel's 496-byte local frame), permutations are applied with
// selects, and the masked kernels, biases and masks are staged in shared memory once per block.  The generic kernel above
// took 0.12 ms at ANY row count - a serial chain of dependent local-memory and global loads - which was 14 % of the
// reference's own p = 50 iteration.
// ---------------------------------------------------------------------------
// Bias and kernel gradient of one masked dense layer of the theta posterior, summed over the rows of a warp: the NO bias
// sums (go) and the NI x NO kernel sums (in[i] * go[o], kernel layout [i][o]) are taken 16 at a time through
// warp_sum16_transposed and added to the gradient blob by the even lanes.  Masked kernel entries receive a gradient only
// under TensorFlow's masked_dense semantics (mg), as before.
template <int NI, int NO>
__device__ __forceinline__ void tf_layer_wgrad(const float (&in)[NI], const float (&go)[NO], const float* mask, bool mg,
                                               float* Gw, float* Gb, int lane) {
    constexpr int N = (NI + 1) * NO;
#pragma unroll
    for (int g0 = 0; g0 < N; g0 += 16) {
        float part[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int f = g0 + k;
            part[k] = f < NO ? go[f < NO ? f : 0] : (f < N ? in[f < N ? (f - NO) / NO : 0] * go[(f - NO) % NO] : 0.f);
        }
        int j;
        const float s = warp_sum16_transposed(part, lane, &j);
        const int f = g0 + j;
        if (!(lane & 1) && f < N) {
            if (f < NO) atomicAdd(Gb + f, s);
            else if (mg || mask[f - NO] != 0.f) atomicAdd(Gw + (f - NO), s);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(128) k_theta_flow_bwd_t(ThetaFlowArgs a) {
    constexpr int LP = (D * TF_H + TF_H) + 2 * (TF_H * TF_H + TF_H) + (TF_H * 2 * D + 2 * D);
    constexpr int O0 = 0, OB0 = D * TF_H, O1 = OB0 + TF_H, OB1 = O1 + TF_H * TF_H, O2 = OB1 + TF_H, OB2 = O2 + TF_H * TF_H,
                  O3 = OB2 + TF_H, OB3 = O3 + TF_H * 2 * D;
    constexpr int MK = D * TF_H + 2 * TF_H * TF_H + TF_H * 2 * D;          // mask entries of one layer
    constexpr int M1 = D * TF_H, M2 = M1 + TF_H * TF_H, M3 = M2 + TF_H * TF_H;
    __shared__ float sP[TF_NBMAX * LP];      // kernels already multiplied by their masks, biases as they are
    __shared__ float sM[MK];
    __shared__ int sPerm[TF_NBMAX * D];
    const int nb = a.nb;
    for (int t = threadIdx.x; t < MK; t += blockDim.x) sM[t] = a.masks[t];
    for (int t = threadIdx.x; t < (nb - 1) * D; t += blockDim.x) sPerm[t] = a.perms[t];
    __syncthreads();
    for (int t = threadIdx.x; t < nb * LP; t += blockDim.x) {
        const int o = t % LP;
        float m = 1.f;
        if (o < OB0) m = sM[o];
        else if (o >= O1 && o < OB1) m = sM[M1 + (o - O1)];
        else if (o >= O2 && o < OB2) m = sM[M2 + (o - O2)];
        else if (o >= O3 && o < OB3) m = sM[M3 + (o - O3)];
        sP[t] = a.params[t] * m;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = r < a.p;
    const bool mg = a.mask_grad != 0;
    const int relu = a.relu;

    auto mlp = [&](const float* P, const float (&z)[D], float (&h1)[TF_H], float (&h2)[TF_H], float (&h3)[TF_H], float (&out)[2 * D]) {
#pragma unroll
        for (int o = 0; o < TF_H; ++o) {
            float v = P[OB0 + o];
#pragma unroll
            for (int i = 0; i < D; ++i) v = fmaf(z[i], P[O0 + i * TF_H + o], v);
            h1[o] = tf_act(v, relu);
        }
#pragma unroll
        for (int o = 0; o < TF_H; ++o) {
            float v = P[OB1 + o];
#pragma unroll
            for (int i = 0; i < TF_H; ++i) v = fmaf(h1[i], P[O1 + i * TF_H + o], v);
            h2[o] = tf_act(v, relu);
        }
#pragma unroll
        for (int o = 0; o < TF_H; ++o) {
            float v = P[OB2 + o];
#pragma unroll
            for (int i = 0; i < TF_H; ++i) v = fmaf(h2[i], P[O2 + i * TF_H + o], v);
            h3[o] = tf_act(v, relu);
        }
#pragma unroll
        for (int o = 0; o < 2 * D; ++o) {
            float v = P[OB3 + o];
#pragma unroll
            for (int i = 0; i < TF_H; ++i) v = fmaf(h3[i], P[O3 + i * 2 * D + o], v);
            out[o] = v;
        }
    };

    float zin[TF_NBMAX][D];
    float z[D], zn[D], h1[TF_H], h2[TF_H], h3[TF_H], out[2 * D];
#pragma unroll
    for (int j = 0; j < D; ++j) z[j] = valid ? a.z0[(size_t)r * D + j] : 0.f;
    // the bijector loops stay rolled: unrolled 8 ways the kernel is ~47 k instructions of straight-line code, more than
    // the instruction cache holds, and at the scripts' row counts (one or two warps) instruction fetch is all it waits for
#pragma unroll 1
    for (int k = 0; k < nb; ++k) {
        {
#pragma unroll
            for (int j = 0; j < D; ++j) zin[k][j] = z[j];
            mlp(sP + k * LP, z, h1, h2, h3, out);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float ls = fminf(fmaxf(out[2 * j + 1], -5.f), 3.f);
                zn[j] = (z[j] - out[2 * j]) * expf(-ls);
            }
            if (k < nb - 1) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const int src = sPerm[k * D + j];
                    float v = 0.f;
#pragma unroll
                    for (int i = 0; i < D; ++i) v = (i == src) ? zn[i] : v;
                    z[j] = v;
                }
            } else {
#pragma unroll
                for (int j = 0; j < D; ++j) z[j] = zn[j];
            }
        }
    }
    const float glp = valid ? (a.g_logq ? a.g_logq[r] : a.g_logq_const) : 0.f;
    float gz[D], gzn[D];
#pragma unroll
    for (int j = 0; j < D; ++j) gz[j] = valid ? a.g_theta[(size_t)r * D + j] : 0.f;
#pragma unroll 1
    for (int k = nb - 1; k >= 0; --k) {
        {
            const float* P = sP + k * LP;
            float* G = a.g_params + (size_t)k * LP;
            if (k < nb - 1) {       // un-permute: z''_j = z'_{perm[j]}
#pragma unroll
                for (int i = 0; i < D; ++i) gzn[i] = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const int src = sPerm[k * D + j];
#pragma unroll
                    for (int i = 0; i < D; ++i) gzn[i] += (i == src) ? gz[j] : 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < D; ++j) gzn[j] = gz[j];
            }
#pragma unroll
            for (int j = 0; j < D; ++j) z[j] = zin[k][j];
            mlp(P, z, h1, h2, h3, out);
            float gout[2 * D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float ls = fminf(fmaxf(out[2 * j + 1], -5.f), 3.f);
                const float e = expf(-ls);
                gz[j] = gzn[j] * e;
                gout[2 * j] = -gzn[j] * e;
                gout[2 * j + 1] = -gzn[j] * (z[j] - out[2 * j]) * e + glp;
            }
            float g3[TF_H], g2[TF_H], g1[TF_H];
#pragma unroll
            for (int i = 0; i < TF_H; ++i) g3[i] = g2[i] = g1[i] = 0.f;
            // Per layer: the bias and kernel gradients are sums over the rows of a warp of (NI + 1) x NO products.  They go
            // through warp_sum16_transposed 16 at a time (16 shuffles per 16 sums) - one warp_sum per entry was 750
            // shuffles per bijector and most of this kernel's instruction stream.
            // layer 3 (no activation)
            tf_layer_wgrad<TF_H, 2 * D>(h3, gout, sM + M3, mg, G + O3, G + OB3, lane);
#pragma unroll
            for (int o = 0; o < 2 * D; ++o)
#pragma unroll
                for (int i = 0; i < TF_H; ++i) g3[i] = fmaf(gout[o], P[O3 + i * 2 * D + o], g3[i]);
            float go2[TF_H], go1[TF_H], go0[TF_H];
#pragma unroll
            for (int o = 0; o < TF_H; ++o) go2[o] = g3[o] * tf_dact(h3[o], relu);
            tf_layer_wgrad<TF_H, TF_H>(h2, go2, sM + M2, mg, G + O2, G + OB2, lane);
#pragma unroll
            for (int o = 0; o < TF_H; ++o)
#pragma unroll
                for (int i = 0; i < TF_H; ++i) g2[i] = fmaf(go2[o], P[O2 + i * TF_H + o], g2[i]);
#pragma unroll
            for (int o = 0; o < TF_H; ++o) go1[o] = g2[o] * tf_dact(h2[o], relu);
            tf_layer_wgrad<TF_H, TF_H>(h1, go1, sM + M1, mg, G + O1, G + OB1, lane);
#pragma unroll
            for (int o = 0; o < TF_H; ++o)
#pragma unroll
                for (int i = 0; i < TF_H; ++i) g1[i] = fmaf(go1[o], P[O1 + i * TF_H + o], g1[i]);
#pragma unroll
            for (int o = 0; o < TF_H; ++o) go0[o] = g1[o] * tf_dact(h1[o], relu);
            tf_layer_wgrad<D, TF_H>(z, go0, sM, mg, G + O0, G + OB0, lane);
#pragma unroll
            for (int o = 0; o < TF_H; ++o)
#pragma unroll
                for (int i = 0; i < D; ++i) gz[i] = fmaf(go0[o], P[O0 + i * TF_H + o], gz[i]);
        }
    }
    if (valid && a.g_z0)
#pragma unroll
        for (int j = 0; j < D; ++j) a.g_z0[(size_t)r * D + j] = gz[j];
}

static int tf_check(int32_t p, int32_t d, int32_t nb, const void* a, const void* b, const void* c) {
    if (p < 1 || d < 1 || d > TF_DMAX || nb < 1 || nb > TF_NBMAX || !a || !b || !c) {
        nma_set_error("nma_theta_flow: bad argument (p=%d, d=%d <= %d, nb=%d <= %d)", p, d, TF_DMAX, nb, TF_NBMAX);
        return -1;
    }
    return 0;
}

extern "C" int nma_theta_flow_fwd(const float* d_params, const float* d_masks, const int32_t* d_perms, const float* d_z0,
                                  int32_t p, int32_t d, int32_t nb, int32_t relu, float base_loc, float base_scale,
                                  float* d_theta, float* d_logq, void* stream) {
    if (tf_check(p, d, nb, d_params, d_masks, d_z0)) return -1;
    if (!d_theta || !d_logq || (nb > 1 && !d_perms)) { nma_set_error("nma_theta_flow_fwd: null pointer"); return -1; }
    ThetaFlowArgs a = {};
    a.params = d_params; a.masks = d_masks; a.perms = d_perms; a.z0 = d_z0; a.theta = d_theta; a.logq = d_logq;
    a.p = p; a.d = d; a.nb = nb; a.relu = relu; a.base_loc = base_loc; a.base_scale = base_scale;
    k_theta_flow_fwd<<<(p + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int nma_theta_flow_bwd_ex(const float* d_params, const float* d_masks, const int32_t* d_perms, const float* d_z0,
                                     int32_t p, int32_t d, int32_t nb, int32_t relu, const float* d_g_theta,
                                     const float* d_g_logq, float g_logq_const, int32_t mask_grad, float* d_g_params,
                                     float* d_g_z0, void* stream) {
    if (tf_check(p, d, nb, d_params, d_masks, d_z0)) return -1;
    if (!d_g_theta || !d_g_params || (nb > 1 && !d_perms)) { nma_set_error("nma_theta_flow_bwd: null pointer"); return -1; }
    ThetaFlowArgs a = {};
    a.params = d_params; a.masks = d_masks; a.perms = d_perms; a.z0 = d_z0;
    a.g_theta = d_g_theta; a.g_logq = d_g_logq; a.g_params = d_g_params; a.g_z0 = d_g_z0;
    a.p = p; a.d = d; a.nb = nb; a.relu = relu; a.g_logq_const = g_logq_const; a.mask_grad = mask_grad;
    const int blocks = (p + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (d == 3) k_theta_flow_bwd_t<3><<<blocks, 128, 0, st>>>(a);
    else if (d == 4) k_theta_flow_bwd_t<4><<<blocks, 128, 0, st>>>(a);
    else if (d == 5) k_theta_flow_bwd_t<5><<<blocks, 128, 0, st>>>(a);
    else k_theta_flow_bwd<<<blocks, 128, 0, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int nma_theta_flow_bwd(const float* d_params, const float* d_masks, const int32_t* d_perms, const float* d_z0,
                                  int32_t p, int32_t d, int32_t nb, int32_t relu, const float* d_g_theta,
                                  const float* d_g_logq, float* d_g_params, float* d_g_z0, void* stream) {
    return nma_theta_flow_bwd_ex(d_params, d_masks, d_perms, d_z0, p, d, nb, relu, d_g_theta, d_g_logq, 0.f, 0, d_g_params,
                                 d_g_z0, stream);
}
