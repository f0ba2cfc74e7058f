// Hand-written backward kernels of the NMA ELBO step (A8 of SURVEY §8a: what TF autodiff derives for
// opt.compute_gradients(-loss), AR.py:228-229).  Gradients are of the SUM over rows.
//   k_epi_bwd    : affine flow layer, softplus head, hidden 1x1 layers (+BN affine), ELU  -> dA, head/hidden grads
//   k_conv_dgrad : data gradient of the K-tap conv (full correlation with the flipped kernel) -> df, dx
//   k_conv_wgrad : weight gradient of the K-tap conv, reduction over rows x positions
//   k_feat_bwd   : feature-MLP backward (4 dense+ELU layers)
//   k_theta_bwd  : theta-bias MLP backward -> d/dtheta and its weight gradients
#include <stdlib.h>
#include "nma_conv_core.cuh"

#include "nma_flow_epi.cuh"
#define BWD_THREADS 256

// ---------------------------------------------------------------------------
// shared helpers on [50][pitch] shared-memory tiles
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load_tile(float* tile, int pitch, const float* __restrict__ g, int gpitch, int nrows,
                                          int ncols = -1) {
    const int n4 = (ncols < 0 ? gpitch : ncols) / 4;
    for (int t = threadIdx.x; t < nrows * n4; t += blockDim.x) {
        const int f = t / n4, j4 = t - f * n4;
        *reinterpret_cast<float4*>(tile + f * pitch + 4 * j4) =
            __ldg(reinterpret_cast<const float4*>(g + (size_t)f * gpitch + 4 * j4));
    }
}

// out[f] = sum_g in[g] * Wsm[g][f] per column, in place (no bias / activation)
__device__ __forceinline__ void col_matvec_inplace(float* tile, int pitch, int npos, const float* Wsm) {
    for (int m = threadIdx.x; m < npos; m += blockDim.x) {
        float* col = tile + m;
        float acc[52];
#pragma unroll
        for (int f = 0; f < 52; ++f) acc[f] = 0.f;
        for (int g = 0; g < NMA_C; ++g) {
            const float xv = col[g * pitch];
            const float4* w4 = reinterpret_cast<const float4*>(Wsm + g * PW_WPITCH);
#pragma unroll
            for (int q = 0; q < 13; ++q) {
                const float4 w = w4[q];
                acc[4 * q + 0] = fmaf(xv, w.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(xv, w.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(xv, w.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(xv, w.w, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int f = 0; f < NMA_C; ++f) col[f * pitch] = acc[f];
    }
}

// thread (f, gg) accumulates acc[g] += sum_m X[f][m] * G[10gg+g][m]   (pad columns must be zero in X or G)
__device__ __forceinline__ void wgrad_accum(const float* X, const float* G, int pitch, int np4, int f, int gg,
                                            float (&acc)[10]) {
    const float4* x4 = reinterpret_cast<const float4*>(X + f * pitch);
    for (int j = 0; j < np4; ++j) {
        const float4 xv = x4[j];
#pragma unroll
        for (int g = 0; g < 10; ++g) {
            const float4 gv = reinterpret_cast<const float4*>(G + (gg * 10 + g) * pitch)[j];
            acc[g] = fmaf(xv.x, gv.x, acc[g]);
            acc[g] = fmaf(xv.y, gv.y, acc[g]);
            acc[g] = fmaf(xv.z, gv.z, acc[g]);
            acc[g] = fmaf(xv.w, gv.w, acc[g]);
        }
    }
}

// per-row sums over positions: s1[g] = sum_m G[g][m], s2[g] = sum_m G[g][m]*E[g][m] (E may be null)
__device__ __forceinline__ void row_sums(const float* G, const float* E, int pitch, int npos, float* s1, float* s2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int g = warp; g < NMA_C; g += nw) {
        float a = 0.f, b = 0.f;
        for (int m = lane; m < npos; m += 32) {
            const float gv = G[g * pitch + m];
            a += gv;
            if (E) b = fmaf(gv, E[g * pitch + m], b);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) { s1[g] = a; if (s2) s2[g] = b; }
    }
}

// ---------------------------------------------------------------------------
// k_epi_bwd
// ---------------------------------------------------------------------------
struct EpiBwdArgs {
    const float* hidw[NMA_MAXH];
    const float* gam[NMA_MAXH];
    const float* bet[NMA_MAXH];
    const float* headw;
    const float* h[NMA_MAXH + 1];  // saved e_0..e_H  [p][50][NP]
    const float* s;                // [p][NP]
    const float* x_in;             // [p][XP]
    const float* dx_next;          // [p][XPn]  d objective / d x^(i+1)
    float* dx;                     // [p][XP]   d objective / d x^(i): direct (affine) part written here
    float* dA;                     // [p][50][NP]
    float* dtb;                    // [p][50]
    float* dat_hi;                 // tensor-core layout of dA: [14][dat_Q][4] at q = K-1 + r*Lin + m (null: SIMT path)
    float* dat_lo;
    long long dat_Q;
    int Lin;
    // gradient blob sections
    float* g_hidw[NMA_MAXH]; float* g_hidb[NMA_MAXH]; float* g_gam[NMA_MAXH]; float* g_bet[NMA_MAXH];
    float* g_headw; float* g_headb; float* g_convb;
    int XP, XPn, L, N, NP, K, H, bn, D, S, p, permute_out, tile_pitch;
    int seg, nseg;                 // a row is processed in nseg segments of seg positions (multiple of 4): one work unit each
    float cq;                      // d objective / d logq
};

__global__ void __launch_bounds__(BWD_THREADS) k_epi_bwd(EpiBwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x;
    const int tp = a.tile_pitch, N = a.N;
    float* E = smem;                         // [50][tp]
    float* G = E + NMA_C * tp;               // [50][tp]
    float* Wsm = G + NMA_C * tp;             // [50][52]
    float* dmu = Wsm + NMA_C * PW_WPITCH;    // [tp]
    float* dsr = dmu + tp;                   // [tp]
    float* v1 = dsr + tp;                    // [64]
    float* v2 = v1 + 64;                     // [64]
    float* bns = v2 + 64;                    // [64]
    float* bno = bns + 64;                   // [64]
    float* hw = bno + 64;                    // [128] head weights

    const int f_own = tid % NMA_C, gg_own = tid / NMA_C;   // wgrad ownership (tid < 250)
    const bool w_owner = tid < 5 * NMA_C;
    float accW[NMA_MAXH][10];
#pragma unroll
    for (int l = 0; l < NMA_MAXH; ++l)
#pragma unroll
        for (int g = 0; g < 10; ++g) accW[l][g] = 0.f;
    float acc_b[NMA_MAXH] = {0.f, 0.f, 0.f, 0.f}, acc_gam[NMA_MAXH] = {0.f, 0.f, 0.f, 0.f},
          acc_bet[NMA_MAXH] = {0.f, 0.f, 0.f, 0.f};
    float acc_head = 0.f, acc_headb = 0.f, acc_convb = 0.f;
    const float rs = rsqrtf(1.f + 1e-3f);

    if (tid < 2 * NMA_C) hw[tid] = a.headw[tid];
    const int np4 = tp / 4;
    // the tile columns beyond NP are never loaded but are swept by wgrad_accum: they must hold finite zeros
    for (int t = tid; t < 2 * NMA_C * tp; t += blockDim.x) smem[t] = 0.f;

    if (a.dat_hi && blockIdx.x == 0) {
        // the tensor-core weight gradient sweeps whole 64-position stages: positions past the last row must read as zero
        // (they may hold rows of an earlier, larger batch)
        const long long q_lo = (long long)a.p * a.Lin + (a.K - 1);
        const int ntail = 128;
        for (int t = tid; t < 14 * ntail; t += blockDim.x) {
            const int fch = t / ntail, q = t - fch * ntail;
            if (q_lo + q < a.dat_Q) {
                const size_t o = ((size_t)fch * a.dat_Q + q_lo + q) * 4;
                *reinterpret_cast<float4*>(a.dat_hi + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(a.dat_lo + o) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    for (int u = blockIdx.x; u < a.p * a.nseg; u += gridDim.x) {
        const int r = u / a.nseg, m0 = (u - r * a.nseg) * a.seg;     // this unit: positions [m0, m0 + n) of row r
        const int n = min(N - m0, a.seg), cols = min(a.NP - m0, a.seg);
        __syncthreads();
        // E <- e_H ; zero G, dmu, dsr
        load_tile(E, tp, a.h[a.H] + (size_t)r * NMA_C * a.NP + m0, a.NP, NMA_C, cols);
        for (int t = tid; t < NMA_C * tp; t += blockDim.x) G[t] = 0.f;
        for (int t = tid; t < tp; t += blockDim.x) { dmu[t] = 0.f; dsr[t] = 0.f; }
        if (tid < NMA_C) {
            if (a.bn && a.H > 0) { bns[tid] = a.gam[a.H - 1][tid] * rs; bno[tid] = a.bet[a.H - 1][tid]; }
            else { bns[tid] = 1.f; bno[tid] = 0.f; }
        }
        if (m0 == 0)
            for (int j = tid; j < a.K && j < a.L; j += blockDim.x) a.dx[(size_t)r * a.XP + j] = 0.f;
        __syncthreads();
        // affine layer + softplus head (AR.py:83-88)
        for (int ml = tid; ml < n; ml += blockDim.x) {
            const int m = m0 + ml;
            const int mo = a.permute_out ? (m ^ 1) : m;
            const float dxo = a.dx_next[(size_t)r * a.XPn + mo];
            float dxin = dxo;
            if (a.D == 1 || (m & 1)) {
                const int mh = (a.D == 1) ? m : m - 1;
                const float sr = a.s[(size_t)r * a.NP + m];
                const float sigma = softplus_f(sr) + 1e-10f;
                const float xin = a.x_in[(size_t)r * a.XP + m + a.K];
                float dsig = dxo * xin;
                if (m >= N - a.S) dsig -= a.cq / sigma;     // logq -= log sigma over the last S slots
                dmu[mh - m0] = dxo;
                dsr[mh - m0] = dsig * sigmoid_f(sr);
                dxin = dxo * sigma;
            }
            a.dx[(size_t)r * a.XP + m + a.K] = dxin;
        }
        __syncthreads();
        // head weight gradients: o_H = BN(e_H)
        if (tid < 2 * NMA_C) {
            const int g = tid >> 1, which = tid & 1;
            const float* vec = which ? dsr : dmu;
            float acc = 0.f;
            for (int m = 0; m < n; ++m) acc = fmaf(fmaf(E[g * tp + m], bns[g], bno[g]), vec[m], acc);
            acc_head += acc;
        } else if (tid < 2 * NMA_C + 2) {
            const float* vec = (tid & 1) ? dsr : dmu;
            float acc = 0.f;
            for (int m = 0; m < n; ++m) acc += vec[m];
            acc_headb += acc;
        }
        // G = d objective / d o_H
        for (int t = tid; t < NMA_C * n; t += blockDim.x) {
            const int g = t / n, m = t - g * n;
            G[g * tp + m] = fmaf(dmu[m], hw[2 * g], dsr[m] * hw[2 * g + 1]);
        }
        __syncthreads();
        for (int l = a.H - 1; l >= 0; --l) {
            // E = e_{l+1} (raw), G = grad w.r.t. o_{l+1}
            if (a.bn) {
                row_sums(G, E, tp, n, v1, v2);
                __syncthreads();
                if (tid < NMA_C) {
                    acc_bet[l] += v1[tid];
                    acc_gam[l] += v2[tid] * rs;
                    bns[tid] = a.gam[l][tid] * rs;
                }
                __syncthreads();
                for (int t = tid; t < NMA_C * n; t += blockDim.x) {
                    const int g = t / n, m = t - g * n;
                    G[g * tp + m] *= bns[g];
                }
                __syncthreads();
            }
            for (int t = tid; t < NMA_C * n; t += blockDim.x) {
                const int g = t / n, m = t - g * n;
                G[g * tp + m] *= elu_grad_from_out(E[g * tp + m]);
            }
            __syncthreads();
            row_sums(G, nullptr, tp, n, v1, nullptr);
            // E <- input of layer l: e_l, BN_{l-1}-transformed when l > 0
            load_tile(E, tp, a.h[l] + (size_t)r * NMA_C * a.NP + m0, a.NP, NMA_C, cols);
            for (int t = tid; t < NMA_C * PW_WPITCH; t += blockDim.x) {   // Wsm[g][f] = W_l[f][g]
                const int g = t / PW_WPITCH, f = t - g * PW_WPITCH;
                Wsm[t] = (f < NMA_C) ? a.hidw[l][f * NMA_C + g] : 0.f;
            }
            __syncthreads();
            if (tid < NMA_C) acc_b[l] += v1[tid];
            const bool bn_in = a.bn && l > 0;
            if (bn_in) {
                if (tid < NMA_C) { bns[tid] = a.gam[l - 1][tid] * rs; bno[tid] = a.bet[l - 1][tid]; }
                __syncthreads();
                for (int t = tid; t < NMA_C * n; t += blockDim.x) {
                    const int g = t / n, m = t - g * n;
                    E[g * tp + m] = fmaf(E[g * tp + m], bns[g], bno[g]);
                }
                __syncthreads();
            }
            if (w_owner) wgrad_accum(E, G, tp, np4, f_own, gg_own, accW[l]);
            __syncthreads();
            col_matvec_inplace(G, tp, n, Wsm);
            if (bn_in) {
                __syncthreads();
                load_tile(E, tp, a.h[l] + (size_t)r * NMA_C * a.NP + m0, a.NP, NMA_C, cols);   // raw e_l again
            }
            __syncthreads();
        }
        // E = e_0, G = grad w.r.t. e_0:  dA = G * elu'(e_0)
        for (int t = tid; t < NMA_C * n; t += blockDim.x) {
            const int g = t / n, m = t - g * n;
            G[g * tp + m] *= elu_grad_from_out(E[g * tp + m]);
        }
        __syncthreads();
        row_sums(G, nullptr, tp, n, v1, nullptr);
        {
            const int n4 = cols / 4;
            for (int t = tid; t < NMA_C * n4; t += blockDim.x) {
                const int f = t / n4, j4 = t - f * n4;
                *reinterpret_cast<float4*>(a.dA + ((size_t)r * NMA_C + f) * a.NP + m0 + 4 * j4) =
                    *reinterpret_cast<const float4*>(G + f * tp + 4 * j4);
            }
        }
        if (a.dat_hi) {
            const long long qb = (long long)(a.K - 1) + (long long)r * a.Lin + m0;
            for (int t = tid; t < 14 * n; t += blockDim.x) {
                const int fch = t / n, m = t - fch * n;
                float v[4], hi[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int f = 4 * fch + e;
                    v[e] = f < NMA_C ? G[f * tp + m] : 0.f;
                    hi[e] = __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
                }
                const size_t o = ((size_t)fch * a.dat_Q + qb + m) * 4;
                *reinterpret_cast<float4*>(a.dat_hi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(a.dat_lo + o) = make_float4(v[0] - hi[0], v[1] - hi[1], v[2] - hi[2], v[3] - hi[3]);
            }
        }
        __syncthreads();
        if (tid < NMA_C) {
            if (a.nseg == 1) a.dtb[(size_t)r * NMA_C + tid] = v1[tid];
            else atomicAdd(a.dtb + (size_t)r * NMA_C + tid, v1[tid]);       // zeroed by the launcher
            acc_convb += v1[tid];
        }
    }
    // flush thread-owned accumulators
    if (w_owner) {
        for (int l = 0; l < a.H; ++l)
#pragma unroll
            for (int g = 0; g < 10; ++g) atomicAdd(a.g_hidw[l] + f_own * NMA_C + gg_own * 10 + g, accW[l][g]);
    }
    if (tid < NMA_C) {
        for (int l = 0; l < a.H; ++l) {
            atomicAdd(a.g_hidb[l] + tid, acc_b[l]);
            if (a.bn) { atomicAdd(a.g_gam[l] + tid, acc_gam[l]); atomicAdd(a.g_bet[l] + tid, acc_bet[l]); }
        }
        atomicAdd(a.g_convb + tid, acc_convb);
    }
    if (tid < 2 * NMA_C) atomicAdd(a.g_headw + tid, acc_head);
    else if (tid < 2 * NMA_C + 2) atomicAdd(a.g_headb + (tid & 1), acc_headb);
}

static int persistent_grid(const nma_handle_s* h, int p, int per_sm) {
    int g = h->sm_count * per_sm;
    return g < p ? g : p;
}
// segments of `seg` positions (a multiple of 4, >= 32): the p x nseg work units run on 2 x SMs resident CTAs in
// ceil(units / CTAs) rounds of 1 / nseg of a row each (+ a fixed share per unit for the weight staging), and nseg is the
// count that makes that product smallest; one segment = the whole padded row when there are rows enough
static void row_segments(const nma_handle_s* h, int p, int N, int NP, int* seg, int* nseg) {
    const int slots = 2 * h->sm_count;
    const int most = (N + 31) / 32;
    int best = 1;
    double best_cost = 1e30;
    for (int n = 1; n <= most; ++n) {
        const int s = ((N + n - 1) / n + 3) & ~3;
        const int real = (N + s - 1) / s;
        if (real != n) continue;
        const long long units = (long long)p * n;
        const long long rounds = (units + slots - 1) / slots;
        const double cost = (double)rounds * (1.0 / n + 0.15);   // 0.15: SV (p = 200) measured no gain from 4 segments
        if (cost < best_cost * 0.98) { best_cost = cost; best = n; }
    }
    if (best <= 1) { *seg = NP; *nseg = 1; return; }
    *seg = ((N + best - 1) / best + 3) & ~3;
    *nseg = (N + *seg - 1) / *seg;
}

int launch_epi_bwd(nma_handle_s* h, int i, const float* params, int p, int objective, float* gp, cudaStream_t st) {
    if (epi_bwd_tc_supported(h)) return launch_epi_bwd_tc(h, i, params, p, objective, gp, st);
    const FlowDims& d = h->fd[i];
    EpiBwdArgs a;
    for (int l = 0; l < NMA_MAXH; ++l) {
        const bool on = l < h->cfg.H;
        a.hidw[l] = on ? params + h->po[i].hidw[l] : nullptr;
        a.gam[l] = (on && h->cfg.bn) ? params + h->po[i].gam[l] : nullptr;
        a.bet[l] = (on && h->cfg.bn) ? params + h->po[i].bet[l] : nullptr;
        a.g_hidw[l] = on ? gp + h->po[i].hidw[l] : nullptr;
        a.g_hidb[l] = on ? gp + h->po[i].hidb[l] : nullptr;
        a.g_gam[l] = (on && h->cfg.bn) ? gp + h->po[i].gam[l] : nullptr;
        a.g_bet[l] = (on && h->cfg.bn) ? gp + h->po[i].bet[l] : nullptr;
    }
    for (int l = 0; l <= NMA_MAXH; ++l) a.h[l] = h->ws[i].h[l];
    a.headw = params + h->po[i].headw;
    a.g_headw = gp + h->po[i].headw; a.g_headb = gp + h->po[i].headb; a.g_convb = gp + h->po[i].convb;
    a.s = h->ws[i].s; a.x_in = h->ws[i].x; a.dx_next = h->ws[i + 1].dx; a.dx = h->ws[i].dx;
    a.dA = h->ws[i].dA; a.dtb = h->ws[i].dtb;
    a.dat_hi = h->use_tc ? h->ws[i].dat_hi : nullptr; a.dat_lo = h->use_tc ? h->ws[i].dat_lo : nullptr;
    a.dat_Q = h->ws[i].dat_Q; a.Lin = d.Lin;
    a.XP = (d.L + 3) & ~3; a.XPn = (h->fd[i + 1].L + 3) & ~3; a.L = d.L; a.N = d.N; a.NP = d.NP; a.K = h->cfg.K;
    a.H = h->cfg.H; a.bn = h->cfg.bn; a.D = h->cfg.D; a.S = h->S; a.p = p;
    a.permute_out = (h->cfg.D == 2 && i < h->cfg.F - 1) ? 1 : 0;
    // fewer rows than CTAs the machine holds (the scripts' own shapes): split the rows into position segments
    row_segments(h, p, d.N, d.NP, &a.seg, &a.nseg);
    a.tile_pitch = a.seg | 4;
    a.cq = (objective == NMA_OBJ_ELBO) ? (float)h->cfg.scale : 0.f;
    if (a.nseg > 1) NMA_CHECK_CUDA(cudaMemsetAsync(a.dtb, 0, (size_t)p * NMA_C * sizeof(float), st));
    const size_t smem = ((size_t)2 * NMA_C * a.tile_pitch + NMA_C * PW_WPITCH + 2 * a.tile_pitch + 4 * 64 + 128) * 4;
    static size_t configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_epi_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    k_epi_bwd<<<persistent_grid(h, p * a.nseg, 2), BWD_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// k_conv_dgrad: dinp[c][j] = sum_k sum_f dA[f][j-k] W[k][c][f]
//   c >= 1 -> df (gradient w.r.t. the feature channels); c == 0 -> added to dx^(i)[j], j < Lin
// Same item decomposition as k_conv_fwd (32 blocks of 10 positions per CTA, rows flattened).
// ---------------------------------------------------------------------------
struct ConvDgradArgs {
    ConvSrc src;
    const float* wdpk;
    float* df;           // [p][50][LP]
    float* dx;           // [p][XP]
    int XP, Lin, LP, KP, rcmax, npb, row_pitch, x_off, p, need_dx;
    long long items_total;
    float* part;         // channel split (gridDim.y > 1), conv_split_reduce
    unsigned* ticket;
};

#define CONVD_WARPS 6
#define DG_COLS (32 * CONV_TM)
#define DG_PITCH (DG_COLS + 4)
__global__ void __launch_bounds__(CONVD_WARPS * 32, 2) k_conv_dgrad(ConvDgradArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t full_bar[CONV_STAGES];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long item0 = (long long)blockIdx.x * 32;
    const int row_first = (int)(item0 / a.npb);
    long long last_item = item0 + 31;
    if (last_item >= a.items_total) last_item = a.items_total - 1;
    const int rc = (int)(last_item / a.npb) - row_first + 1;

    ConvRing rg;
    rg.cin = NMA_C; rg.ngroups = 6; rg.KP = a.KP; rg.rc = rc; rg.row_pitch = a.row_pitch;
    rg.stage_floats = 6 * a.KP * CONV_WPAD + a.rcmax * a.row_pitch;
    for (int t = tid; t < CONV_STAGES * rg.stage_floats; t += blockDim.x) smem[t] = 0.f;
    if (tid == 0) {
        for (int s = 0; s < CONV_STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    __syncthreads();

    const long long item = item0 + lane;
    const bool item_ok = item < a.items_total;
    const int my_r = item_ok ? (int)(item / a.npb) : row_first;
    const int pb = item_ok ? (int)(item - (long long)my_r * a.npb) : 0;
    const bool wide = warp < 5;
    const bool active = item_ok && (wide || a.need_dx);
    const int j0 = pb * CONV_TM;

    float2 acc[CONV_TM][5];
    {
        const int cper = (NMA_C + (int)gridDim.y - 1) / (int)gridDim.y, c0 = (int)blockIdx.y * cper;
        conv_main_loop(acc, smem, full_bar, rg, a.src, a.wdpk, row_first, my_r - row_first, j0 + a.x_off, warp, active, wide,
                       c0, min(NMA_C, c0 + cper));
    }
    if (!conv_split_reduce(acc, a.part, a.ticket)) return;

    // stage the result through shared memory for coalesced stores: tile[51][DG_PITCH], row 50 = x-channel
    float* tile = smem;
    if (active) {
#pragma unroll
        for (int j = 0; j < CONV_TM; ++j) {
            if (wide) {
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    float* t0 = tile + (size_t)(warp * 10 + 2 * q) * DG_PITCH + lane * CONV_TM + j;
                    t0[0] = acc[j][q].x;
                    t0[DG_PITCH] = acc[j][q].y;
                }
            } else {
                tile[(size_t)NMA_C * DG_PITCH + lane * CONV_TM + j] = acc[j][0].x;
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < NMA_C1 * DG_COLS; t += blockDim.x) {
        const int f = t / DG_COLS, col = t - f * DG_COLS;
        const long long it = item0 + col / CONV_TM;
        if (it >= a.items_total) continue;
        const int r = (int)(it / a.npb);
        const int jj = (int)(it - (long long)r * a.npb) * CONV_TM + col % CONV_TM;
        if (f < NMA_C) {
            if (jj < a.LP) a.df[((size_t)r * NMA_C + f) * a.LP + jj] = (jj < a.Lin) ? tile[(size_t)f * DG_PITCH + col] : 0.f;
        } else if (a.need_dx && jj < a.Lin) {
            a.dx[(size_t)r * a.XP + jj] += tile[(size_t)NMA_C * DG_PITCH + col];
        }
    }
}

int launch_conv_dgrad(nma_handle_s* h, int i, int p, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    ConvDgradArgs a;
    const int K = h->cfg.K;
    const int npb = (d.Lin + CONV_TM - 1) / CONV_TM;
    a.rcmax = (npb - 1 + 31) / npb + 1;
    a.npb = npb; a.p = p;
    a.items_total = (long long)p * npb;
    a.XP = (d.L + 3) & ~3; a.Lin = d.Lin; a.LP = d.LP; a.KP = h->KP;
    const int padl = (K - 1 + 3) & ~3;          // left zero pad, 16B aligned for the bulk copy
    a.x_off = padl - (K - 1);                   // smem offset of (output j = 0, tap k' = 0)
    int rp = padl + d.NP + 4;
    const int need = a.x_off + npb * CONV_TM + h->KP + 4;
    if (rp < need) rp = need;
    a.row_pitch = (rp + 3) & ~3;
    a.src.chan0 = nullptr; a.src.row_stride0 = 0;
    a.src.rest = h->ws[i].dA; a.src.row_stride = (long long)NMA_C * d.NP; a.src.chan_stride = d.NP;
    a.src.copy_floats = d.NP; a.src.dst_off = padl;
    a.wdpk = h->ws[i].wdpk;
    a.df = h->ws[i].df; a.dx = h->ws[i].dx;
    a.need_dx = i > 0 ? 1 : 0;
    const size_t ring = (size_t)CONV_STAGES * (6 * h->KP * CONV_WPAD + a.rcmax * a.row_pitch);
    const size_t epi = (size_t)NMA_C1 * DG_PITCH;
    const size_t smem = (ring > epi ? ring : epi) * 4;
    static size_t configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_dgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int item_ctas = (int)((a.items_total + 31) / 32);
    const int nsplit = conv_split_count(item_ctas, NMA_C, h->sm_count);
    a.part = h->split_part; a.ticket = h->split_ticket;
    k_conv_dgrad<<<dim3((unsigned)item_ctas, nsplit), CONVD_WARPS * 32, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// k_conv_wgrad: dW[k][c][f] = sum_r sum_m inp[r][c][m+k] dA[r][f][m]
//   thread item = (input channel c, block of 10 taps, group of 5 output channels): 50 accumulators,
//   sliding 13-float window over m, 4 positions per iteration.
// ---------------------------------------------------------------------------
#define WG_THREADS 320
struct ConvWgradArgs {
    const float* x;      // [p][XP]  channel 0
    const float* a4;     // [p][50][LP] channels 1..50
    const float* dA;     // [p][50][NP]
    float* gW;           // [K][51][50] gradient section
    int XP, LP, NP, N, K, ntb, nitems, p, rows_per_cta, ip, dp;
};

__global__ void __launch_bounds__(WG_THREADS) k_conv_wgrad(ConvWgradArgs a) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x;
    float* Dt = smem;                       // [50][dp]  dA tile (float4 access: keep it 16B aligned)
    float* I = Dt + NMA_C * a.dp;           // [51][ip]  conv input tile of one row (odd pitch, scalar access)
    // item decomposition: item = (fg*ntb + tb)*51 + c ; c fastest so a warp reads 32 distinct input rows
    const int item = blockIdx.x * WG_THREADS + tid;
    const bool active = item < a.nitems;
    const int c = item % NMA_C1;
    const int tb = (item / NMA_C1) % a.ntb;
    const int fg = item / (NMA_C1 * a.ntb);
    const int k0 = tb * 10, f0 = fg * 5;

    float acc[10][5];
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int f = 0; f < 5; ++f) acc[k][f] = 0.f;

    const int r_begin = blockIdx.y * a.rows_per_cta;
    const int r_end = min(a.p, r_begin + a.rows_per_cta);
    const int nq = (a.N + 3) / 4;
    // zero once: pad columns of both tiles must read as zero
    for (int t = tid; t < NMA_C1 * a.ip + NMA_C * a.dp; t += blockDim.x) smem[t] = 0.f;
    for (int r = r_begin; r < r_end; ++r) {
        __syncthreads();
        {   // load tiles (coalesced float4)
            const int n4 = a.LP / 4;
            for (int t = tid; t < NMA_C1 * n4; t += blockDim.x) {
                const int cc = t / n4, j4 = t - cc * n4;
                const float* src = (cc == 0) ? a.x + (size_t)r * a.XP + 4 * j4
                                             : a.a4 + ((size_t)r * NMA_C + (cc - 1)) * a.LP + 4 * j4;
                const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                float* dst = I + cc * a.ip + 4 * j4;      // odd pitch: scalar stores
                dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
            }
            const int m4 = a.NP / 4;
            for (int t = tid; t < NMA_C * m4; t += blockDim.x) {
                const int f = t / m4, j4 = t - f * m4;
                *reinterpret_cast<float4*>(Dt + f * a.dp + 4 * j4) =
                    __ldg(reinterpret_cast<const float4*>(a.dA + ((size_t)r * NMA_C + f) * a.NP + 4 * j4));
            }
        }
        __syncthreads();
        if (active) {
            const float* ir = I + c * a.ip + k0;
            float w[13];
#pragma unroll
            for (int j = 0; j < 9; ++j) w[j + 4] = ir[j];     // window for m = 0 sits in w[4..12] after the shift below
            for (int q = 0; q < nq; ++q) {
                const int m = 4 * q;
                // slide by 4: w[0..8] <- w[4..12], load 4 new
#pragma unroll
                for (int j = 0; j < 9; ++j) w[j] = w[j + 4];
                w[9] = ir[m + 9]; w[10] = ir[m + 10]; w[11] = ir[m + 11]; w[12] = ir[m + 12];
                float4 d[5];
#pragma unroll
                for (int f = 0; f < 5; ++f) d[f] = *reinterpret_cast<const float4*>(Dt + (f0 + f) * a.dp + m);
#pragma unroll
                for (int k = 0; k < 10; ++k)
#pragma unroll
                    for (int f = 0; f < 5; ++f) {
                        acc[k][f] = fmaf(w[k + 0], d[f].x, acc[k][f]);
                        acc[k][f] = fmaf(w[k + 1], d[f].y, acc[k][f]);
                        acc[k][f] = fmaf(w[k + 2], d[f].z, acc[k][f]);
                        acc[k][f] = fmaf(w[k + 3], d[f].w, acc[k][f]);
                    }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            if (k0 + k < a.K) {
#pragma unroll
                for (int f = 0; f < 5; ++f)
                    atomicAdd(a.gW + ((size_t)(k0 + k) * NMA_C1 + c) * NMA_C + f0 + f, acc[k][f]);
            }
        }
    }
}

int launch_conv_wgrad(nma_handle_s* h, int i, int p, float* gp, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    ConvWgradArgs a;
    a.x = h->ws[i].x; a.a4 = h->ws[i].a[4]; a.dA = h->ws[i].dA; a.gW = gp + h->po[i].convw;
    a.XP = (d.L + 3) & ~3; a.LP = d.LP; a.NP = d.NP; a.N = d.N; a.K = h->cfg.K; a.p = p;
    a.ntb = h->KP / 10;
    a.nitems = 10 * a.ntb * NMA_C1;
    // smem pitches: input rows must cover m + k0 + 12 for m up to roundup4(N)-4 (+KP), and be == 2 mod 4-ish odd-friendly
    int ip = ((d.N + 3) & ~3) + h->KP + 16;
    if (ip < d.LP) ip = d.LP;
    a.ip = ip | 1;                 // odd pitch: lanes walk distinct channels -> conflict-free scalar loads
    a.dp = (((d.N + 3) & ~3) + 4) | 4;
    if (a.dp < d.NP) a.dp = d.NP | 4;
    const int item_ctas = (a.nitems + WG_THREADS - 1) / WG_THREADS;
    int row_groups = (2 * h->sm_count + item_ctas - 1) / item_ctas;
    if (row_groups > p) row_groups = p;
    if (row_groups < 1) row_groups = 1;
    a.rows_per_cta = (p + row_groups - 1) / row_groups;
    row_groups = (p + a.rows_per_cta - 1) / a.rows_per_cta;
    const size_t smem = ((size_t)NMA_C1 * a.ip + (size_t)NMA_C * a.dp) * 4;
    static size_t configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    k_conv_wgrad<<<dim3(item_ctas, row_groups), WG_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// k_feat_bwd: backward of 4x dense(C, elu) over the window positions (AR.py:53-56 differentiated).
//
// Layer-outer, rows-inner: a CTA owns a fixed set of rows and sweeps them once per layer (l = 3..0), so the
// weight-gradient tile of ONE layer lives in registers for the whole sweep (5 inputs x 10 outputs per thread,
// 15 shared-memory float4 loads per 200 FMAs) and the transposed weights are staged once per layer instead
// of once per row.  The gradient w.r.t. the layer input overwrites the row's df slab in place (only this CTA
// ever touches that row), and elu' is applied while the next sweep loads it.
// ---------------------------------------------------------------------------
#define FB_WPITCH 64         // transposed weights [g][4 output groups][16]: group fg holds outputs f = 4*i + fg
struct FeatBwdArgs {
    const float* w[4];
    const float* act[5];     // a0 [p][Cf_in][LP], a1..a4 [p][50][LP]
    float* df;               // [p][50][LP]  in: d objective / d a4;  overwritten layer by layer
    float* gw[4]; float* gb[4];
    int Lin, LP, Cf_in, p, tile_pitch;
    int top;                 // last dense(50) layer: 3; Lotka-Volterra: 2 (its 4th layer is the wide one, nma_lv.cu)
    int seg, nseg;           // work unit = one segment of seg positions (multiple of 4) of one row (row_segments)
};

__global__ void __launch_bounds__(BWD_THREADS, 2) k_feat_bwd(FeatBwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x, tp = a.tile_pitch;
    float* X = smem;                          // [50][tp] layer input a_l
    float* G = X + NMA_C * tp;                // [50][tp] gradient w.r.t. the layer's pre-activation
    float* Wt = G + NMA_C * tp;               // [50][FB_WPITCH]
    float* v1 = Wt + NMA_C * FB_WPITCH;       // [64]
    float* wred = v1 + 64;                    // [50][50] cross-slice reduction of the weight-gradient tiles
    const int np4 = tp / 4;
    // weight-gradient ownership: 50 (5 x 10) tiles x 5 position slices
    const bool w_owner = tid < 250;
    const int slice = tid / 50, wt_tile = tid % 50;
    const int f0 = (wt_tile % 10) * 5, g0 = (wt_tile / 10) * 10;
    const int j_lo = (np4 * slice) / 5, j_hi = (np4 * (slice + 1)) / 5;
    // data-gradient work items: (4 columns mg) x (13 outputs f = 4*i + fg), 4 * np4 items per row; a window
    // wider than 64 float4 columns (SV: 302 slots) needs more items than the CTA has threads, hence the loop

    for (int t = tid; t < 2 * NMA_C * tp; t += blockDim.x) smem[t] = 0.f;     // pad columns stay zero
    for (int l = a.top; l >= 0; --l) {
        const int nin = (l == 0) ? a.Cf_in : NMA_C;
        __syncthreads();
        if (l > 0) {    // Wt[g][fg][i] = W_l[f = 4i + fg][g]
            for (int t = tid; t < NMA_C * FB_WPITCH; t += blockDim.x) {
                const int g = t / FB_WPITCH, q = t - g * FB_WPITCH;
                const int f = 4 * (q & 15) + (q >> 4);
                Wt[t] = ((q & 15) < 13 && f < NMA_C) ? a.w[l][f * NMA_C + g] : 0.f;
            }
        }
        for (int t = tid; t < NMA_C * NMA_C; t += blockDim.x) wred[t] = 0.f;
        float acc[5][10];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int g = 0; g < 10; ++g) acc[i][g] = 0.f;
        float acc_b = 0.f;

        for (int u = blockIdx.x; u < a.p * a.nseg; u += gridDim.x) {
            const int r = u / a.nseg, m0 = (u - r * a.nseg) * a.seg;          // positions [m0, m0 + n) of row r
            const int n = min(a.Lin - m0, a.seg), n4 = min(a.LP - m0, a.seg) / 4;
            __syncthreads();
            // G = df * elu'(a_{l+1}),  X = a_l   (coalesced float4; columns >= n are zero in G)
            {
                const float* gsrc = a.df + (size_t)r * NMA_C * a.LP + m0;
                const float* esrc = a.act[l + 1] + (size_t)r * NMA_C * a.LP + m0;
                for (int t = tid; t < NMA_C * np4; t += blockDim.x) {
                    const int f = t / np4, j4 = t - f * np4;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j4 < n4) {
                        // plain load: df is rewritten by this kernel between sweeps, the read-only path is not coherent
                        const float4 gv = *reinterpret_cast<const float4*>(gsrc + (size_t)f * a.LP + 4 * j4);
                        const float4 ev = __ldg(reinterpret_cast<const float4*>(esrc + (size_t)f * a.LP + 4 * j4));
                        o.x = (4 * j4 + 0 < n) ? gv.x * elu_grad_from_out(ev.x) : 0.f;
                        o.y = (4 * j4 + 1 < n) ? gv.y * elu_grad_from_out(ev.y) : 0.f;
                        o.z = (4 * j4 + 2 < n) ? gv.z * elu_grad_from_out(ev.z) : 0.f;
                        o.w = (4 * j4 + 3 < n) ? gv.w * elu_grad_from_out(ev.w) : 0.f;
                    }
                    *reinterpret_cast<float4*>(G + f * tp + 4 * j4) = o;
                }
                const float* xsrc = a.act[l] + (size_t)r * nin * a.LP + m0;
                for (int t = tid; t < nin * n4; t += blockDim.x) {
                    const int f = t / n4, j4 = t - f * n4;
                    *reinterpret_cast<float4*>(X + f * tp + 4 * j4) =
                        __ldg(reinterpret_cast<const float4*>(xsrc + (size_t)f * a.LP + 4 * j4));
                }
            }
            __syncthreads();
            row_sums(G, nullptr, tp, n, v1, nullptr);
            // weight gradient: acc[i][g] += sum_m X[f0+i][m] * G[g0+g][m] over this thread's slice of positions
            if (w_owner && f0 < nin) {
                for (int jj = j_lo; jj < j_hi; ++jj) {
                    float4 xv[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) xv[i] = *reinterpret_cast<const float4*>(X + (f0 + i) * tp + 4 * jj);
#pragma unroll
                    for (int g = 0; g < 10; ++g) {
                        const float4 gv = *reinterpret_cast<const float4*>(G + (g0 + g) * tp + 4 * jj);
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            acc[i][g] = fmaf(xv[i].x, gv.x, acc[i][g]);
                            acc[i][g] = fmaf(xv[i].y, gv.y, acc[i][g]);
                            acc[i][g] = fmaf(xv[i].z, gv.z, acc[i][g]);
                            acc[i][g] = fmaf(xv[i].w, gv.w, acc[i][g]);
                        }
                    }
                }
            }
            // data gradient (l > 0): d a_l[f][m] = sum_g G[g][m] * W_l[f][g]  -> overwrites df[r] in place
            if (l > 0) for (int it = tid; it < 4 * np4; it += blockDim.x) {
                const int mg = it % np4, fg = it / np4;
                float4 d[13];
#pragma unroll
                for (int i = 0; i < 13; ++i) d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int g = 0; g < NMA_C; ++g) {
                    const float4 gv = *reinterpret_cast<const float4*>(G + g * tp + 4 * mg);
                    const float4* w4 = reinterpret_cast<const float4*>(Wt + g * FB_WPITCH + fg * 16);
                    const float4 wa = w4[0], wb = w4[1], wc = w4[2], wd = w4[3];
                    const float w[13] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w, wd.x};
#pragma unroll
                    for (int i = 0; i < 13; ++i) {
                        d[i].x = fmaf(gv.x, w[i], d[i].x);
                        d[i].y = fmaf(gv.y, w[i], d[i].y);
                        d[i].z = fmaf(gv.z, w[i], d[i].z);
                        d[i].w = fmaf(gv.w, w[i], d[i].w);
                    }
                }
                if (mg < n4) {
                    float* dst = a.df + (size_t)r * NMA_C * a.LP + m0 + 4 * mg;
#pragma unroll
                    for (int i = 0; i < 13; ++i) {
                        const int f = 4 * i + fg;
                        if (f < NMA_C) *reinterpret_cast<float4*>(dst + (size_t)f * a.LP) = d[i];
                    }
                }
            }
            __syncthreads();
            if (tid < NMA_C) acc_b += v1[tid];
        }
        // combine the 5 position slices in shared memory, then one atomic per weight per CTA
        if (w_owner && f0 < nin) {
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int g = 0; g < 10; ++g) atomicAdd(wred + (f0 + i) * NMA_C + g0 + g, acc[i][g]);
        }
        __syncthreads();
        for (int t = tid; t < nin * NMA_C; t += blockDim.x) atomicAdd(a.gw[l] + t, wred[t]);
        if (tid < NMA_C) atomicAdd(a.gb[l] + tid, acc_b);
    }
}

int launch_feat_bwd(nma_handle_s* h, int i, const float* params, int p, float* gp, cudaStream_t st) {
    if (h->use_tc && h->use_tc_feat) return launch_feat_bwd_tc(h, i, params, p, gp, st);
    const FlowDims& d = h->fd[i];
    FeatBwdArgs a;
    for (int l = 0; l < 4; ++l) {
        a.w[l] = params + h->po[i].featw[l];
        a.gw[l] = gp + h->po[i].featw[l];
        a.gb[l] = gp + h->po[i].featb[l];
    }
    for (int l = 0; l < 5; ++l) a.act[l] = h->ws[i].a[l];
    a.df = h->ws[i].df; a.Lin = d.Lin; a.LP = d.LP; a.Cf_in = h->Cf_in; a.p = p;
    a.top = 3;
    if (h->is_lv) {          // three dense(50) layers over the whole window; df3 = d objective / d a3 (nma_lv.cu)
        a.df = h->ws[i].df3; a.Lin = h->LW; a.LP = h->LWP; a.top = 2;
    }
    row_segments(h, p, a.Lin, a.LP, &a.seg, &a.nseg);
    a.tile_pitch = a.seg | 4;
    const size_t smem = ((size_t)2 * NMA_C * a.tile_pitch + NMA_C * FB_WPITCH + 64 + NMA_C * NMA_C) * 4;
    static size_t configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_feat_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    k_feat_bwd<<<persistent_grid(h, p * a.nseg, 2), BWD_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// k_theta_bwd: b = W3(W2(W1 theta + b1) + b2) + b3 (AR.py:63-68); upstream = dtb[r][f] = sum_m dA
// ---------------------------------------------------------------------------
struct ThetaBwdArgs {
    const float* w[NMA_MAX_FLOWS][3];
    const float* tb[NMA_MAX_FLOWS];     // [p][3][50]: t1, t2, b
    const float* dtb[NMA_MAX_FLOWS];    // [p][50]
    float* gw[NMA_MAX_FLOWS][3];
    float* gb[NMA_MAX_FLOWS][3];
    const float* theta;
    float* grad_theta;                  // [p][dth], accumulated with atomics (already holds the ELBO part)
    int p, dth, rows_per_cta, flow0;
};

__global__ void __launch_bounds__(BWD_THREADS) k_theta_bwd(ThetaBwdArgs a) {
    const int i = a.flow0 + blockIdx.y, tid = threadIdx.x;
    __shared__ float W2[NMA_C * NMA_C], W3[NMA_C * NMA_C], W1[8 * NMA_C];
    __shared__ float d3[NMA_C], d2[NMA_C], d1[NMA_C], t1[NMA_C], t2[NMA_C], th[8];
    for (int t = tid; t < NMA_C * NMA_C; t += blockDim.x) { W2[t] = a.w[i][1][t]; W3[t] = a.w[i][2][t]; }
    for (int t = tid; t < a.dth * NMA_C; t += blockDim.x) W1[t] = a.w[i][0][t];
    const int f_own = tid % NMA_C, gg_own = tid / NMA_C;
    const bool owner = tid < 5 * NMA_C;
    float aW3[10], aW2[10], aW1[10];
#pragma unroll
    for (int g = 0; g < 10; ++g) aW3[g] = aW2[g] = aW1[g] = 0.f;
    float ab3 = 0.f, ab2 = 0.f, ab1 = 0.f;
    const int r0 = blockIdx.x * a.rows_per_cta, r1 = min(a.p, r0 + a.rows_per_cta);
    for (int r = r0; r < r1; ++r) {
        __syncthreads();
        if (tid < NMA_C) {
            d3[tid] = a.dtb[i][(size_t)r * NMA_C + tid];
            t1[tid] = a.tb[i][(size_t)r * 3 * NMA_C + tid];
            t2[tid] = a.tb[i][(size_t)r * 3 * NMA_C + NMA_C + tid];
        }
        if (tid < a.dth) th[tid] = a.theta[(size_t)r * a.dth + tid];
        __syncthreads();
        if (tid < NMA_C) {   // d2[a] = sum_b W3[a][b] d3[b]
            float v = 0.f;
            for (int b = 0; b < NMA_C; ++b) v = fmaf(W3[tid * NMA_C + b], d3[b], v);
            d2[tid] = v;
            ab3 += d3[tid];
        }
        __syncthreads();
        if (tid < NMA_C) {
            float v = 0.f;
            for (int b = 0; b < NMA_C; ++b) v = fmaf(W2[tid * NMA_C + b], d2[b], v);
            d1[tid] = v;
            ab2 += d2[tid];
        }
        __syncthreads();
        if (tid < NMA_C) ab1 += d1[tid];
        if (tid < a.dth) {
            float v = 0.f;
            for (int b = 0; b < NMA_C; ++b) v = fmaf(W1[tid * NMA_C + b], d1[b], v);
            atomicAdd(a.grad_theta + (size_t)r * a.dth + tid, v);
        }
        if (owner) {
#pragma unroll
            for (int g = 0; g < 10; ++g) {
                aW3[g] = fmaf(t2[f_own], d3[gg_own * 10 + g], aW3[g]);
                aW2[g] = fmaf(t1[f_own], d2[gg_own * 10 + g], aW2[g]);
                if (f_own < a.dth) aW1[g] = fmaf(th[f_own], d1[gg_own * 10 + g], aW1[g]);
            }
        }
    }
    if (owner) {
#pragma unroll
        for (int g = 0; g < 10; ++g) {
            atomicAdd(a.gw[i][2] + f_own * NMA_C + gg_own * 10 + g, aW3[g]);
            atomicAdd(a.gw[i][1] + f_own * NMA_C + gg_own * 10 + g, aW2[g]);
            if (f_own < a.dth) atomicAdd(a.gw[i][0] + f_own * NMA_C + gg_own * 10 + g, aW1[g]);
        }
    }
    if (tid < NMA_C) {
        atomicAdd(a.gb[i][2] + tid, ab3);
        atomicAdd(a.gb[i][1] + tid, ab2);
        atomicAdd(a.gb[i][0] + tid, ab1);
    }
}

int launch_theta_bwd(nma_handle_s* h, const float* params, const float* theta, int p, float* gp, float* grad_theta,
                     int flow, cudaStream_t st) {
    ThetaBwdArgs a;
    a.flow0 = flow < 0 ? 0 : flow;
    for (int i = 0; i < h->cfg.F; ++i) {
        for (int l = 0; l < 3; ++l) {
            a.w[i][l] = params + h->po[i].thw[l];
            a.gw[i][l] = gp + h->po[i].thw[l];
            a.gb[i][l] = gp + h->po[i].thb[l];
        }
        a.tb[i] = h->ws[i].tb;
        a.dtb[i] = h->ws[i].dtb;
    }
    a.theta = theta; a.grad_theta = grad_theta; a.p = p; a.dth = h->cfg.dtheta;
    int ctas = h->sm_count;
    if (ctas > p) ctas = p;
    a.rows_per_cta = (p + ctas - 1) / ctas;
    ctas = (p + a.rows_per_cta - 1) / a.rows_per_cta;
    k_theta_bwd<<<dim3(ctas, flow < 0 ? h->cfg.F : 1), BWD_THREADS, 0, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}
