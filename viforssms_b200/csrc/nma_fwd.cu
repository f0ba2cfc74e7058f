// Forward kernels of the NMA ELBO step: weight packing, theta-bias MLP (AR.py:63-68), window gather +
// feature MLP (AR.py:53-56, 267-283), fused conv + head + affine flow layer (AR.py:58-89), ELBO terms
// (AR.py:168-187).  sm_100a only.
#include <string.h>
#include "nma_conv_core.cuh"
#include "nma_flow_epi.cuh"

// ---------------------------------------------------------------------------
// nma_gather: materialise the feed the reference builds on the host every iteration
// ---------------------------------------------------------------------------
__global__ void k_gather(SeriesView sv, const int64_t* __restrict__ idx, int p, int L0, int B, float x0a, float x0b, int npin,
                         float* __restrict__ tf, float* __restrict__ mask, float* __restrict__ shift) {
    const int r = blockIdx.x;
    const long long win0 = (long long)sv.D * idx[r];
    const int n = L0 * sv.Cf;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int slot = t / sv.Cf, c = t - slot * sv.Cf;
        tf[(size_t)r * n + t] = series_val(sv, c, win0 + slot);
    }
    // mask_vals = [0,1,1,...], shift_vals = [x0,0,0,...] sliced at idx (AR.py:145-148,285-288); the Lotka-Volterra batch
    // scripts pin the first p_val states (lotka_volterra_partial_batch.py:237-240)
    const int nm = sv.D * (B + 1);
    for (int t = threadIdx.x; t < nm; t += blockDim.x) {
        const int d = t / (B + 1), j = t - d * (B + 1);
        const bool first = (idx[r] + j) < npin;
        if (mask) mask[(size_t)r * nm + t] = first ? 0.f : 1.f;
        if (shift) shift[(size_t)r * nm + t] = first ? (d == 0 ? x0a : x0b) : 0.f;
    }
}

int launch_gather(nma_handle_s* h, const int64_t* idx, int p, float* tf, float* mask, float* shift, cudaStream_t st) {
    SeriesView sv = nma_series_view(h);
    k_gather<<<p, 256, 0, st>>>(sv, idx, p, h->L0, h->cfg.B, h->cfg.x0[0], h->cfg.x0[1], h->cfg.n_pinned > 0 ? h->cfg.n_pinned : 1,
                                tf, mask, shift);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// weight packing: conv kernel [K][51][50] -> per input channel slabs [groups][KP][12]
// ---------------------------------------------------------------------------
__global__ void k_pack_fwd(const float* __restrict__ W, int K, int KP, int cin, float* __restrict__ out) {
    // out[c][g][k][12], c < cin (51; 1 + window for the Lotka-Volterra conv), g<5
    const int n = cin * 5 * KP * CONV_WPAD;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        int q = t % CONV_WPAD;
        int k = (t / CONV_WPAD) % KP;
        int g = (t / (CONV_WPAD * KP)) % 5;
        int c = t / (CONV_WPAD * KP * 5);
        float v = 0.f;
        if (q < 10 && k < K) v = W[((size_t)k * cin + c) * NMA_C + g * 10 + q];
        out[t] = v;
    }
}
__global__ void k_pack_dgrad(const float* __restrict__ W, int K, int KP, float* __restrict__ out) {
    // out[f][g][k'][12], f<50 (input = conv output channel), g<6; output o=(g,q): g<5 -> c = 1+10g+q, g==5,q==0 -> c=0
    // tap k' uses W[K-1-k'][c][f]  (full correlation of dA with the flipped kernel)
    const int n = NMA_C * 6 * KP * CONV_WPAD;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        int q = t % CONV_WPAD;
        int k = (t / CONV_WPAD) % KP;
        int g = (t / (CONV_WPAD * KP)) % 6;
        int f = t / (CONV_WPAD * KP * 6);
        float v = 0.f;
        int c = -1;
        if (g < 5 && q < 10) c = 1 + 10 * g + q;
        if (g == 5 && q == 0) c = 0;
        if (c >= 0 && k < K) v = W[((size_t)(K - 1 - k) * NMA_C1 + c) * NMA_C + f];
        out[t] = v;
    }
}

int launch_pack_weights(nma_handle_s* h, const float* params, bool need_bwd, cudaStream_t st, int which) {
    if (h->use_tc) return launch_pack_weights_tc(h, params, need_bwd, st, which);
    if (!(which & 1)) return 0;
    for (int i = 0; i < h->cfg.F; ++i) {
        k_pack_fwd<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, h->KP, h->conv_cin, h->ws[i].wpk);
        nma_count_launch(1);
        if (need_bwd && !h->is_lv) k_pack_dgrad<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, h->KP, h->ws[i].wdpk);
    }
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// theta-bias MLP (AR.py:63-68): three linear layers, no activation. tb[r] = {t1, t2, b + conv_bias}
// ---------------------------------------------------------------------------
struct ThetaFwdArgs {
    const float* w[NMA_MAX_FLOWS][3];
    const float* b[NMA_MAX_FLOWS][3];
    const float* convb[NMA_MAX_FLOWS];
    float* tb[NMA_MAX_FLOWS];
};
__global__ void k_theta_fwd(ThetaFwdArgs a, const float* __restrict__ theta, int dth) {
    const int r = blockIdx.x, i = blockIdx.y, f = threadIdx.x;
    __shared__ float s_in[64];
    __shared__ float s_t[64];
    if (f < dth) s_in[f] = theta[(size_t)r * dth + f];
    __syncthreads();
    float* out = a.tb[i] + (size_t)r * 3 * NMA_C;
    float v = 0.f;
    if (f < NMA_C) {
        v = a.b[i][0][f];
        for (int k = 0; k < dth; ++k) v = fmaf(s_in[k], a.w[i][0][k * NMA_C + f], v);
        out[f] = v;
        s_t[f] = v;
    }
    __syncthreads();
    if (f < NMA_C) {
        v = a.b[i][1][f];
        for (int k = 0; k < NMA_C; ++k) v = fmaf(s_t[k], a.w[i][1][k * NMA_C + f], v);
        out[NMA_C + f] = v;
    }
    __syncthreads();
    if (f < NMA_C) s_t[f] = v;
    __syncthreads();
    if (f < NMA_C) {
        v = a.b[i][2][f] + a.convb[i][f];
        for (int k = 0; k < NMA_C; ++k) v = fmaf(s_t[k], a.w[i][2][k * NMA_C + f], v);
        out[2 * NMA_C + f] = v;
    }
}

int launch_theta_fwd(nma_handle_s* h, const float* params, const float* theta, int p, cudaStream_t st) {
    ThetaFwdArgs a;
    for (int i = 0; i < h->cfg.F; ++i) {
        for (int l = 0; l < 3; ++l) {
            a.w[i][l] = params + h->po[i].thw[l];
            a.b[i][l] = params + h->po[i].thb[l];
        }
        a.convb[i] = params + h->po[i].convb;
        a.tb[i] = h->ws[i].tb;
    }
    k_theta_fwd<<<dim3(p, h->cfg.F), 64, 0, st>>>(a, theta, h->cfg.dtheta);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// feature MLP (A1 + A2): gather the window of one row for one flow, run 4x dense(C, elu)
// ---------------------------------------------------------------------------
#define FEAT_THREADS 256
#define FEAT_WPITCH 60   // 5 groups x 12 floats

// Y[g][j] = elu(b[g] + sum_f W[f][g] X[f][j]);  thread item = (4 positions) x (10 outputs)
__device__ __forceinline__ void dense_tile_elu(const float* Xs, int ldx, int nin, const float* Wsm, const float* bsm,
                                               float* Ys, int ldy, int npos4, float* gout, int gld, int nvalid) {
    for (int it = threadIdx.x; it < npos4 * 5; it += blockDim.x) {
        const int jg = it % npos4, gg = it / npos4;
        float4 acc[10];
#pragma unroll
        for (int g = 0; g < 10; ++g) {
            const float b = bsm[gg * 10 + g];
            acc[g] = make_float4(b, b, b, b);
        }
        for (int f = 0; f < nin; ++f) {
            const float4 xv = *reinterpret_cast<const float4*>(Xs + f * ldx + 4 * jg);
            const float4* w4 = reinterpret_cast<const float4*>(Wsm + f * FEAT_WPITCH + gg * 12);
            const float4 wa = w4[0], wb = w4[1], wc = w4[2];
            const float w[10] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y};
#pragma unroll
            for (int g = 0; g < 10; ++g) {
                acc[g].x = fmaf(w[g], xv.x, acc[g].x);
                acc[g].y = fmaf(w[g], xv.y, acc[g].y);
                acc[g].z = fmaf(w[g], xv.z, acc[g].z);
                acc[g].w = fmaf(w[g], xv.w, acc[g].w);
            }
        }
#pragma unroll
        for (int g = 0; g < 10; ++g) {
            float4 o = make_float4(elu_f(acc[g].x), elu_f(acc[g].y), elu_f(acc[g].z), elu_f(acc[g].w));
            // keep the pad columns (>= nvalid) at zero so downstream padded reads stay finite
            const int j0 = 4 * jg;
            if (j0 + 0 >= nvalid) o.x = 0.f;
            if (j0 + 1 >= nvalid) o.y = 0.f;
            if (j0 + 2 >= nvalid) o.z = 0.f;
            if (j0 + 3 >= nvalid) o.w = 0.f;
            *reinterpret_cast<float4*>(Ys + (gg * 10 + g) * ldy + j0) = o;
            if (gout) *reinterpret_cast<float4*>(gout + (size_t)(gg * 10 + g) * gld + j0) = o;
        }
    }
}

// stage a dense kernel [nin][50] (+bias[50]) from global into the padded smem layout [nin][60]
__device__ __forceinline__ void stage_dense_w(const float* __restrict__ W, const float* __restrict__ b, int nin,
                                              float* Wsm, float* bsm) {
    for (int t = threadIdx.x; t < nin * FEAT_WPITCH; t += blockDim.x) {
        const int f = t / FEAT_WPITCH, q = t - f * FEAT_WPITCH;
        const int gg = q / 12, g = q - gg * 12;
        Wsm[t] = (g < 10) ? W[f * NMA_C + gg * 10 + g] : 0.f;
    }
    for (int t = threadIdx.x; t < NMA_C; t += blockDim.x) bsm[t] = b[t];
}

struct FeatArgs {
    const float* w[NMA_MAX_FLOWS][4];
    const float* b[NMA_MAX_FLOWS][4];
    float* a[NMA_MAX_FLOWS][5];
    float* x0;        // ws[0].x : aligned copy of eps
    float* tin_hi[NMA_MAX_FLOWS];   // tensor-core layout of the conv input (null: SIMT conv)
    float* tin_lo[NMA_MAX_FLOWS];
    long long tin_Q[NMA_MAX_FLOWS];
    int Lin[NMA_MAX_FLOWS], LP[NMA_MAX_FLOWS];
    int XP0;
};

__global__ void __launch_bounds__(FEAT_THREADS) k_feat_fwd(FeatArgs fa, SeriesView sv, const int64_t* __restrict__ idx,
                                                           const float* __restrict__ eps, int L0, int K, int Cf_in,
                                                           int feat_off, int save) {
    extern __shared__ __align__(128) float smem[];
    const int r = blockIdx.x, i = blockIdx.y;
    const int Lin = fa.Lin[i], LP = fa.LP[i];
    const int lds = LP;                               // smem pitch
    float* T0 = smem;                                 // [50][lds]
    float* T1 = T0 + NMA_C * lds;                     // [50][lds]
    float* Wsm = T1 + NMA_C * lds;                    // [50][60]
    float* bsm = Wsm + NMA_C * FEAT_WPITCH;           // [64]
    const long long win0 = (long long)sv.D * idx[r];

    if (i == 0) {   // aligned copy of the base sample (x^(0) = eps, AR.py:31-32)
        for (int t = threadIdx.x; t < fa.XP0; t += blockDim.x)
            fa.x0[(size_t)r * fa.XP0 + t] = (t < L0) ? eps[(size_t)r * L0 + t] : 0.f;
    }
    // gather: slot = i*K + j + feat_off  (AR.py:192-193 then :53; SV_dense.py:53 uses [1:])
    float* ga0 = save ? fa.a[i][0] + (size_t)r * Cf_in * LP : nullptr;
    for (int t = threadIdx.x; t < Cf_in * LP; t += blockDim.x) {
        const int c = t / LP, j = t - c * LP;
        float v = 0.f;
        if (j < Lin) {
            const long long pos = win0 + (long long)i * K + j + feat_off;
            if (c < sv.Cf)
                v = series_val(sv, c, pos);
            else
                v = series_val(sv, c - sv.Cf, pos) - series_val(sv, c - sv.Cf, pos - 1);
        }
        T0[c * lds + j] = v;
        if (ga0) ga0[t] = v;
    }
    const int npos4 = LP / 4;
    float* cur = T0;
    float* nxt = T1;
    for (int l = 0; l < 4; ++l) {
        __syncthreads();
        const int nin = (l == 0) ? Cf_in : NMA_C;
        stage_dense_w(fa.w[i][l], fa.b[i][l], nin, Wsm, bsm);
        __syncthreads();
        float* gout = (save || l == 3) ? fa.a[i][l + 1] + (size_t)r * NMA_C * LP : nullptr;
        dense_tile_elu(cur, lds, nin, Wsm, bsm, nxt, lds, npos4, gout, LP, Lin);
        float* t = cur; cur = nxt; nxt = t;
    }
    if (fa.tin_hi[i]) {
        // conv input in the tensor-core layout [channel/4][q = r*Lin + slot][4], split for 3xTF32.
        // channel 0 is the flow's input sample: eps for flow 0; flows > 0 get it from the previous flow's epilogue.
        __syncthreads();
        float* th = fa.tin_hi[i];
        float* tl = fa.tin_lo[i];
        const long long Q = fa.tin_Q[i], qrow = (long long)r * Lin;
        for (int t = threadIdx.x; t < 14 * Lin; t += blockDim.x) {
            const int cch = t / Lin, j = t - cch * Lin;
            float v[4], hi[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * cch + e;
                if (c == 0) v[e] = (i == 0) ? eps[(size_t)r * L0 + j] : 0.f;
                else v[e] = (c <= NMA_C) ? cur[(c - 1) * lds + j] : 0.f;
                hi[e] = __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
            }
            const size_t o = ((size_t)cch * Q + qrow + j) * 4;
            *reinterpret_cast<float4*>(th + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(tl + o) = make_float4(v[0] - hi[0], v[1] - hi[1], v[2] - hi[2], v[3] - hi[3]);
        }
    }
}

int launch_feat_fwd(nma_handle_s* h, const float* params, const int64_t* idx, int p, bool save, cudaStream_t st);

static int feat_smem_bytes(const nma_handle_s* h) {
    int lp = 0;
    for (int i = 0; i < h->cfg.F; ++i) lp = lp > h->fd[i].LP ? lp : h->fd[i].LP;
    return (2 * NMA_C * lp + NMA_C * FEAT_WPITCH + 64) * 4;
}

// eps is passed through the handle-level wrapper (see nma_api.cu) via this file-scope pointer-free path
int launch_feat_fwd_eps(nma_handle_s* h, const float* params, const int64_t* idx, const float* eps, int p, bool save,
                        cudaStream_t st) {
    if (h->use_tc && h->use_tc_feat) return launch_feat_fwd_tc(h, params, idx, eps, p, save, st);
    FeatArgs fa;
    for (int i = 0; i < h->cfg.F; ++i) {
        for (int l = 0; l < 4; ++l) {
            fa.w[i][l] = params + h->po[i].featw[l];
            fa.b[i][l] = params + h->po[i].featb[l];
        }
        for (int l = 0; l < 5; ++l) fa.a[i][l] = h->ws[i].a[l];
        fa.Lin[i] = h->fd[i].Lin;
        fa.LP[i] = h->fd[i].LP;
        fa.tin_hi[i] = h->use_tc ? h->ws[i].tin_hi : nullptr;
        fa.tin_lo[i] = h->use_tc ? h->ws[i].tin_lo : nullptr;
        fa.tin_Q[i] = h->ws[i].tin_Q;
    }
    fa.x0 = h->ws[0].x;
    fa.XP0 = (h->fd[0].L + 3) & ~3;
    const int smem = feat_smem_bytes(h);
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_feat_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    SeriesView sv = nma_series_view(h);
    k_feat_fwd<<<dim3(p, h->cfg.F), FEAT_THREADS, smem, st>>>(fa, sv, idx, eps, h->L0, h->cfg.K, h->Cf_in,
                                                               h->feat_off, save ? 1 : 0);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// fused conv + theta-bias + ELU + hidden 1x1 layers + head + softplus + affine flow layer (A3-A5)
//
// Work decomposition: an "item" is one block of 10 consecutive output positions of one row; items of
// all rows are numbered consecutively and every CTA takes 32 of them (one per lane), so lanes stay
// full whatever N_i is.  A CTA therefore spans a few consecutive rows; each ring stage holds the
// current input channel of all of them.
// ---------------------------------------------------------------------------
#define CONVF_WARPS 5
#define CONVF_THREADS (CONVF_WARPS * 32)
#define ITEM_COLS (32 * CONV_TM)
#define ITEM_PITCH (ITEM_COLS + 4)   // 324: multiple of 4 and (pitch/4) odd

struct ConvFwdArgs {
    ConvSrc src;
    const float* wpk;
    const float* tb;         // [p][3][50]; slot 2 = theta bias + conv bias
    FlowEpiArgs e;
    int KP, rcmax, npb, row_pitch, p, cin;
    long long items_total;
    float* part;             // channel split (gridDim.y > 1): partial accumulators and tickets, conv_split_reduce
    unsigned* ticket;
};

__global__ void __launch_bounds__(CONVF_THREADS, 2) k_conv_fwd(ConvFwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t full_bar[CONV_STAGES];
    __shared__ unsigned char col_ok[ITEM_COLS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long item0 = (long long)blockIdx.x * 32;
    const int row_first = (int)(item0 / a.npb);
    long long last_item = item0 + 31;
    if (last_item >= a.items_total) last_item = a.items_total - 1;
    const int rc = (int)(last_item / a.npb) - row_first + 1;

    ConvRing rg;
    rg.cin = a.cin; rg.ngroups = 5; rg.KP = a.KP; rg.rc = rc; rg.row_pitch = a.row_pitch;
    rg.stage_floats = 5 * a.KP * CONV_WPAD + a.rcmax * a.row_pitch;

    // zero the ring once: the tail of every input row must read as 0 (finite) under the padded taps
    for (int t = tid; t < CONV_STAGES * rg.stage_floats; t += blockDim.x) smem[t] = 0.f;
    if (tid == 0) {
        for (int s = 0; s < CONV_STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    __syncthreads();

    const long long item = item0 + lane;
    const bool active = item < a.items_total;
    const int my_r = active ? (int)(item / a.npb) : row_first;
    const int pb = active ? (int)(item - (long long)my_r * a.npb) : 0;
    const int m0 = pb * CONV_TM;

    float2 acc[CONV_TM][5];
    {
        const int cper = (a.cin + (int)gridDim.y - 1) / (int)gridDim.y, c0 = (int)blockIdx.y * cper;
        conv_main_loop(acc, smem, full_bar, rg, a.src, a.wpk, row_first, my_r - row_first, m0, warp, active, true, c0,
                       min(a.cin, c0 + cper));
    }
    // (conv_main_loop ends with __syncthreads: the ring is free and is reused as the activation tile)
    if (!conv_split_reduce(acc, a.part, a.ticket)) return;

    float* tile = smem;                               // [50][ITEM_PITCH], column = lane*10 + j
    __shared__ int col_r[ITEM_COLS];
    __shared__ int col_m[ITEM_COLS];
    for (int col = tid; col < ITEM_COLS; col += blockDim.x) {
        const long long it = item0 + col / CONV_TM;
        bool ok = false;
        int r = 0, m = 0;
        if (it < a.items_total) {
            r = (int)(it / a.npb);
            m = (int)(it - (long long)r * a.npb) * CONV_TM + col % CONV_TM;
            ok = m < a.e.N;
        }
        col_ok[col] = ok ? 1 : 0;
        col_r[col] = r;
        col_m[col] = m;
    }
    // e_0 = elu(A + theta-bias + conv bias)  (AR.py:70-72)
    if (active) {
        const float* tbr = a.tb + ((size_t)my_r * 3 + 2) * NMA_C + warp * 10;
        float bias[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) bias[q] = tbr[q];
#pragma unroll
        for (int j = 0; j < CONV_TM; ++j) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float* t0 = tile + (size_t)(warp * 10 + 2 * q) * ITEM_PITCH + lane * CONV_TM + j;
                t0[0] = elu_f(acc[j][q].x + bias[2 * q]);
                t0[ITEM_PITCH] = elu_f(acc[j][q].y + bias[2 * q + 1]);
            }
        }
    }
    __syncthreads();
    flow_epilogue<ITEM_COLS>(a.e, tile, col_r, col_m, col_ok);
}

static int conv_rows_spanned(int npb) { return (npb - 1 + 31) / npb + 1; }

void fill_flow_epi_args(nma_handle_s* h, int i, const float* params, bool save, FlowEpiArgs& e) {
    const FlowDims& d = h->fd[i];
    e.XP = (d.L + 3) & ~3;
    e.XPn = (h->fd[i + 1].L + 3) & ~3;
    e.N = d.N; e.NP = d.NP; e.K = h->cfg.K; e.H = h->cfg.H; e.bn = h->cfg.bn; e.D = h->cfg.D;
    e.save = save ? 1 : 0;
    e.permute_out = (h->cfg.D == 2 && i < h->cfg.F - 1) ? 1 : 0;
    for (int l = 0; l < NMA_MAXH; ++l) {
        e.hidw[l] = l < h->cfg.H ? params + h->po[i].hidw[l] : nullptr;
        e.hidb[l] = l < h->cfg.H ? params + h->po[i].hidb[l] : nullptr;
        e.gam[l] = (l < h->cfg.H && h->cfg.bn) ? params + h->po[i].gam[l] : nullptr;
        e.bet[l] = (l < h->cfg.H && h->cfg.bn) ? params + h->po[i].bet[l] : nullptr;
    }
    for (int l = 0; l <= NMA_MAXH; ++l) e.h[l] = h->ws[i].h[l];
    e.headw = params + h->po[i].headw; e.headb = params + h->po[i].headb;
    e.x_in = h->ws[i].x; e.x_out = h->ws[i + 1].x; e.s = h->ws[i].s;
    const bool next_tc = h->use_tc && (i + 1 < h->cfg.F);
    e.nx_hi = next_tc ? h->ws[i + 1].tin_hi : nullptr;
    e.nx_lo = next_tc ? h->ws[i + 1].tin_lo : nullptr;
    e.nx_Q = next_tc ? h->ws[i + 1].tin_Q : 0;
    e.nx_Lin = next_tc ? h->fd[i + 1].Lin : 0;
}

int launch_conv_fwd(nma_handle_s* h, int i, const float* params, int p, bool save, cudaStream_t st) {
    if (h->use_tc) return launch_conv_fwd_tc(h, i, params, p, save, st);
    const FlowDims& d = h->fd[i];
    ConvFwdArgs a;
    const int npb = (d.N + CONV_TM - 1) / CONV_TM;
    a.rcmax = conv_rows_spanned(npb);
    a.npb = npb;
    a.items_total = (long long)p * npb;
    a.KP = h->KP; a.p = p;
    fill_flow_epi_args(h, i, params, save, a.e);
    int rp = npb * CONV_TM + h->KP + 4;
    if (rp < d.LP) rp = d.LP;
    a.row_pitch = (rp + 3) & ~3;
    a.src.chan0 = h->ws[i].x; a.src.row_stride0 = a.e.XP;
    // channels 1.. : the feature activations [50][LP]; Lotka-Volterra: the transposed 4th layer, [window][LP]
    a.src.rest = h->ws[i].a[4]; a.src.row_stride = (long long)(h->conv_cin - 1) * d.LP; a.src.chan_stride = d.LP;
    a.src.copy_floats = d.LP; a.src.dst_off = 0;
    a.wpk = h->ws[i].wpk;
    a.tb = h->ws[i].tb;
    a.cin = h->conv_cin;

    const size_t ring = (size_t)CONV_STAGES * (5 * h->KP * CONV_WPAD + a.rcmax * a.row_pitch);
    const size_t epi = flow_epi_smem_floats<ITEM_COLS>();
    const size_t smem = (ring > epi ? ring : epi) * 4;
    static size_t configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const long long grid = (a.items_total + 31) / 32;
    const int nsplit = conv_split_count((int)grid, a.cin, h->sm_count);
    a.part = h->split_part; a.ticket = h->split_ticket;
    k_conv_fwd<<<dim3((unsigned)grid, nsplit), CONVF_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// Lotka-Volterra feature MLP (lotka_volterra_partial_batch_fix_theta.py:71-76): every flow reads the WHOLE window
// (LW = L0 - 1 positions, no i*K offset, :343-344), runs 3 x dense(50, elu) and a 4th dense layer as wide as the
// flow's conv input (feat_dims = L_i - 1 units), and the [window position w][unit m] result is transposed: it is
// stored exactly like that, a4[r][w][m], which makes w the conv's input channel 1 + w and m its position.
// The script runs p = 1 (one series per iteration): the window positions are pointwise through all four layers, so a
// (row, flow) is cut into gridDim.z segments of `seg` window positions, one CTA each.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(FEAT_THREADS) k_lv_feat_fwd(FeatArgs fa, SeriesView sv, const int64_t* __restrict__ idx,
                                                              const float* __restrict__ eps, int L0, int Cf_in, int LW,
                                                              int LWP, int save, int seg) {
    extern __shared__ __align__(128) float smem[];
    const int r = blockIdx.x, i = blockIdx.y;
    const int w0 = (int)blockIdx.z * seg;                       // this CTA: window positions [w0, w0 + n)
    const int n = min(LW - w0, seg), cols = min(LWP - w0, seg); // cols: padded to a multiple of 4
    const int Fd = fa.Lin[i], LP = fa.LP[i];
    const int ld = seg;                                         // shared-memory pitch
    float* T0 = smem;
    float* T1 = T0 + NMA_C * ld;
    float* Wsm = T1 + NMA_C * ld;
    float* bsm = Wsm + NMA_C * FEAT_WPITCH;
    const long long win0 = (long long)sv.D * idx[r];
    if (i == 0 && blockIdx.z == 0) {
        for (int t = threadIdx.x; t < fa.XP0; t += blockDim.x)
            fa.x0[(size_t)r * fa.XP0 + t] = (t < L0) ? eps[(size_t)r * L0 + t] : 0.f;
    }
    float* ga0 = save ? fa.a[i][0] + (size_t)r * Cf_in * LWP + w0 : nullptr;
    for (int t = threadIdx.x; t < Cf_in * cols; t += blockDim.x) {
        const int c = t / cols, w = t - c * cols;
        const float v = (w < n) ? series_val(sv, c, win0 + w0 + w) : 0.f;
        T0[c * ld + w] = v;
        if (ga0) ga0[(size_t)c * LWP + w] = v;
    }
    float* cur = T0;
    float* nxt = T1;
    for (int l = 0; l < 3; ++l) {
        __syncthreads();
        const int nin = (l == 0) ? Cf_in : NMA_C;
        stage_dense_w(fa.w[i][l], fa.b[i][l], nin, Wsm, bsm);
        __syncthreads();
        float* gout = save ? fa.a[i][l + 1] + (size_t)r * NMA_C * LWP + w0 : nullptr;
        dense_tile_elu(cur, ld, nin, Wsm, bsm, nxt, ld, cols / 4, gout, LWP, n);
        float* t = cur; cur = nxt; nxt = t;
    }
    __syncthreads();
    // 4th layer: a4[w][m] = elu(b[m] + sum_f a3[f][w] W4[f][m]), lanes along the units m (kernel rows are read coalesced)
    const float* __restrict__ W4 = fa.w[i][3];
    const float* __restrict__ b4 = fa.b[i][3];
    float* out = fa.a[i][4] + ((size_t)r * LW + w0) * LP;
    for (int t = threadIdx.x; t < n * LP; t += blockDim.x) {
        const int w = t / LP, m = t - w * LP;
        float v = 0.f;
        if (m < Fd) {
            float acc = b4[m];
            for (int f = 0; f < NMA_C; ++f) acc = fmaf(cur[f * ld + w], __ldg(W4 + (size_t)f * Fd + m), acc);
            v = elu_f(acc);
        }
        out[t] = v;
    }
}

int launch_lv_feat_fwd(nma_handle_s* h, const float* params, const int64_t* idx, const float* eps, int p, bool save,
                       cudaStream_t st) {
    FeatArgs fa;
    memset(&fa, 0, sizeof(fa));
    for (int i = 0; i < h->cfg.F; ++i) {
        for (int l = 0; l < 4; ++l) {
            fa.w[i][l] = params + h->po[i].featw[l];
            fa.b[i][l] = params + h->po[i].featb[l];
        }
        for (int l = 0; l < 5; ++l) fa.a[i][l] = h->ws[i].a[l];
        fa.Lin[i] = h->fd[i].Lin;
        fa.LP[i] = h->fd[i].LP;
    }
    fa.x0 = h->ws[0].x;
    fa.XP0 = (h->fd[0].L + 3) & ~3;
    // segments of the window so that p x F x nseg CTAs fill the machine (>= 16 positions each)
    int seg = h->LWP, nseg = 1;
    {
        int want = (2 * h->sm_count) / (p * h->cfg.F);
        const int most = (h->LW + 15) / 16;
        if (want > most) want = most;
        if (want > 1) { seg = ((h->LW + want - 1) / want + 3) & ~3; nseg = (h->LW + seg - 1) / seg; }
    }
    const int smem = (2 * NMA_C * seg + NMA_C * FEAT_WPITCH + 64) * 4;
    if (smem > 227 * 1024) { nma_set_error("Lotka-Volterra window of %d positions does not fit in shared memory", h->LW); return -1; }
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_lv_feat_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    SeriesView sv = nma_series_view(h);
    k_lv_feat_fwd<<<dim3(p, h->cfg.F, nseg), FEAT_THREADS, smem, st>>>(fa, sv, idx, eps, h->L0, h->Cf_in, h->LW, h->LWP,
                                                                        save ? 1 : 0, seg);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}
