// Backward kernels of the Lotka-Volterra instance of the NMA flow (lotka_volterra_partial_batch_fix_theta.py:71-82
// differentiated).  The script runs p_val = 1 - one 151-step series, 363 window positions, per iteration - so these
// launches are latency-bound whatever their inner loops look like; they are written for clarity and coalesced
// access (one thread per output element, reductions in registers), not for tensor-core throughput.
//   k_lv_conv_dgrad : d objective / d conv input: channel 0 -> dx^(i), channel 1 + w -> df[r][w][m]  (gradient of a4)
//   k_lv_conv_wgrad : gW[k][c][f] = sum_r sum_m inp[r][m + k][c] dA[r][f][m]
//   k_lv_feat4_*    : the wide 4th feature layer: weight / bias gradients and the gradient w.r.t. a3
#include "nma_common.cuh"

// dinp[r][j][c] = sum_k sum_f dA[r][f][j - k] W[k][c][f],  0 <= j - k < N
__global__ void k_lv_conv_dgrad(const float* __restrict__ dA, const float* __restrict__ W, float* __restrict__ df,
                                float* __restrict__ dx, int K, int cin, int N, int NP, int Lin, int LP, int XP, int LW,
                                int need_dx) {
    const int r = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cin * LP) return;
    const int c = t / LP, j = t - c * LP;              // lanes along the position: dA rows are read coalesced
    float acc = 0.f;
    if (j < Lin) {
        const float* dr = dA + (size_t)r * NMA_C * NP;
        for (int k = 0; k < K; ++k) {
            const int m = j - k;
            if (m < 0 || m >= N) continue;
            const float* wk = W + ((size_t)k * cin + c) * NMA_C;
            for (int f = 0; f < NMA_C; ++f) acc = fmaf(dr[(size_t)f * NP + m], __ldg(wk + f), acc);
        }
    }
    if (c == 0) {
        if (need_dx && j < Lin) dx[(size_t)r * XP + j] += acc;
    } else {
        df[((size_t)r * LW + (c - 1)) * LP + j] = acc;  // pad columns (j >= Lin) get 0
    }
}

int launch_lv_conv_dgrad(nma_handle_s* h, int i, const float* params, int p, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    const int n = h->conv_cin * d.LP;
    k_lv_conv_dgrad<<<dim3((n + 255) / 256, p), 256, 0, st>>>(h->ws[i].dA, params + h->po[i].convw, h->ws[i].df,
                                                               h->ws[i].dx, h->cfg.K, h->conv_cin, d.N, d.NP, d.Lin, d.LP,
                                                               (d.L + 3) & ~3, h->LW, i > 0 ? 1 : 0);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// one thread per weight; inp channel 0 = x^(i), channel 1 + w = a4[r][w][.]
__global__ void k_lv_conv_wgrad(const float* __restrict__ x, const float* __restrict__ a4, const float* __restrict__ dA,
                                float* __restrict__ gW, int K, int cin, int N, int NP, int LP, int XP, int LW, int p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= K * cin * NMA_C) return;
    const int f = t % NMA_C, c = (t / NMA_C) % cin, k = t / (NMA_C * cin);
    float acc = 0.f;
    for (int r = 0; r < p; ++r) {
        const float* in = (c == 0) ? x + (size_t)r * XP + k : a4 + ((size_t)r * LW + (c - 1)) * LP + k;
        const float* dr = dA + ((size_t)r * NMA_C + f) * NP;
        for (int m = 0; m < N; ++m) acc = fmaf(__ldg(in + m), __ldg(dr + m), acc);
    }
    gW[t] += acc;       // t == (k * cin + c) * 50 + f: the conv kernel's own [K][Cin][Cout] layout
}

// The same sums with dA[r] staged once per CTA in shared memory ([50][odd pitch]: lanes along f read distinct banks).
// CTA = a range of input channels; thread = (output channel f, group of LVW_TG taps): the input row slides through
// registers, one shared-memory load per position feeds LVW_TG FMAs.
#define LVW_TG 4
__global__ void __launch_bounds__(256) k_lv_conv_wgrad_sm(const float* __restrict__ x, const float* __restrict__ a4,
                                                          const float* __restrict__ dA, float* __restrict__ gW, int K,
                                                          int cin, int N, int NP, int LP, int XP, int LW, int p, int cper) {
    extern __shared__ __align__(16) float smem[];
    const int dp = NP | 1;                          // odd pitch
    float* dsm = smem;                              // [50][dp]
    float* ins = dsm + NMA_C * dp;                  // [N + K + LVW_TG] one input channel of one row
    const int tid = threadIdx.x;
    const int ntg = (K + LVW_TG - 1) / LVW_TG;
    const int f = tid % NMA_C, tg = tid / NMA_C;    // tg < ntg: taps [tg * LVW_TG, ...)
    const bool owner = tg < ntg;
    const int c0 = blockIdx.x * cper, c1 = min(cin, c0 + cper);
    for (int c = c0; c < c1; ++c) {
        float acc[LVW_TG];
#pragma unroll
        for (int u = 0; u < LVW_TG; ++u) acc[u] = 0.f;
        for (int r = 0; r < p; ++r) {
            __syncthreads();
            if (c == c0 || p > 1) {
                const float* dr = dA + (size_t)r * NMA_C * NP;
                for (int t = tid; t < NMA_C * NP; t += blockDim.x) {
                    const int ff = t / NP, m = t - ff * NP;
                    dsm[ff * dp + m] = (m < N) ? dr[t] : 0.f;
                }
            }
            const float* in = (c == 0) ? x + (size_t)r * XP : a4 + ((size_t)r * LW + (c - 1)) * LP;
            const int avail = (c == 0) ? XP : LP;   // floats of this row that exist
            for (int t = tid; t < N + K + LVW_TG; t += blockDim.x) ins[t] = (t < avail) ? __ldg(in + t) : 0.f;
            __syncthreads();
            if (owner) {
                const float* dm = dsm + f * dp;
                const float* iw = ins + tg * LVW_TG;
                float win[LVW_TG];
#pragma unroll
                for (int u = 0; u < LVW_TG - 1; ++u) win[u] = iw[u];
                for (int m = 0; m < N; ++m) {
                    win[LVW_TG - 1] = iw[m + LVW_TG - 1];
                    const float g = dm[m];
#pragma unroll
                    for (int u = 0; u < LVW_TG; ++u) acc[u] = fmaf(win[u], g, acc[u]);
#pragma unroll
                    for (int u = 0; u < LVW_TG - 1; ++u) win[u] = win[u + 1];
                }
            }
        }
        if (owner) {
#pragma unroll
            for (int u = 0; u < LVW_TG; ++u) {
                const int k = tg * LVW_TG + u;
                if (k < K) gW[((size_t)k * cin + c) * NMA_C + f] += acc[u];     // one CTA owns a channel: no atomics
            }
        }
    }
}

int launch_lv_conv_wgrad(nma_handle_s* h, int i, int p, float* gp, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    const int n = h->cfg.K * h->conv_cin * NMA_C;
    const int ntg = (h->cfg.K + LVW_TG - 1) / LVW_TG;
    const size_t smem = ((size_t)NMA_C * (d.NP | 1) + d.N + h->cfg.K + LVW_TG + 4) * 4;
    if (ntg * NMA_C <= 256 && smem <= 200 * 1024) {
        static size_t configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_lv_conv_wgrad_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        int ctas = 2 * h->sm_count;
        if (ctas > h->conv_cin) ctas = h->conv_cin;
        const int cper = (h->conv_cin + ctas - 1) / ctas;
        ctas = (h->conv_cin + cper - 1) / cper;
        k_lv_conv_wgrad_sm<<<ctas, 256, smem, st>>>(h->ws[i].x, h->ws[i].a[4], h->ws[i].dA, gp + h->po[i].convw, h->cfg.K,
                                                   h->conv_cin, d.N, d.NP, d.LP, (d.L + 3) & ~3, h->LW, p, cper);
        nma_count_launch(1);
        NMA_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    k_lv_conv_wgrad<<<(n + 255) / 256, 256, 0, st>>>(h->ws[i].x, h->ws[i].a[4], h->ws[i].dA, gp + h->po[i].convw,
                                                     h->cfg.K, h->conv_cin, d.N, d.NP, d.LP, (d.L + 3) & ~3, h->LW, p);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// 4th feature layer a4[w][m] = elu(b[m] + sum_f a3[f][w] W4[f][m]); G[w][m] = df[w][m] * elu'(a4[w][m]).
// weight gradient (f < 50) and bias gradient (f == 50): one thread per (f, m), lanes along m
__global__ void k_lv_feat4_wgrad(const float* __restrict__ df, const float* __restrict__ a4, const float* __restrict__ a3,
                                 float* __restrict__ gW4, float* __restrict__ gb4, int Fd, int LP, int LW, int LWP, int p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (NMA_C + 1) * Fd) return;
    const int f = t / Fd, m = t - f * Fd;
    // the window sum is cut into gridDim.y slices (the script's p = 1 leaves no other parallelism)
    const int wper = (LW + gridDim.y - 1) / gridDim.y, wa = blockIdx.y * wper, wb = min(LW, wa + wper);
    float acc = 0.f;
    for (int r = 0; r < p; ++r) {
        const float* dr = df + (size_t)r * LW * LP + m;
        const float* ar = a4 + (size_t)r * LW * LP + m;
        const float* xr = a3 + ((size_t)r * NMA_C + f) * LWP;
        for (int w = wa; w < wb; ++w) {
            const float g = dr[(size_t)w * LP] * elu_grad_from_out(ar[(size_t)w * LP]);
            acc = fmaf(f < NMA_C ? __ldg(xr + w) : 1.f, g, acc);
        }
    }
    if (f < NMA_C) atomicAdd(gW4 + (size_t)f * Fd + m, acc);
    else atomicAdd(gb4 + m, acc);
}
// gradient w.r.t. a3: df3[r][f][w] = sum_m G[w][m] W4[f][m]; one warp per (r, f, w), lanes along m
__global__ void k_lv_feat4_dgrad(const float* __restrict__ df, const float* __restrict__ a4, const float* __restrict__ W4,
                                 float* __restrict__ df3, int Fd, int LP, int LW, int LWP, int p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p * NMA_C * LWP) return;
    const int w = warp % LWP, f = (warp / LWP) % NMA_C, r = warp / (LWP * NMA_C);
    float acc = 0.f;
    if (w < LW) {
        const float* dr = df + ((size_t)r * LW + w) * LP;
        const float* ar = a4 + ((size_t)r * LW + w) * LP;
        const float* wr = W4 + (size_t)f * Fd;
        for (int m = lane; m < Fd; m += 32) acc = fmaf(dr[m] * elu_grad_from_out(ar[m]), __ldg(wr + m), acc);
        acc = warp_sum(acc);
    }
    if (lane == 0) df3[((size_t)r * NMA_C + f) * LWP + w] = acc;     // pad columns (w >= LW) get 0
}

int launch_lv_feat4_bwd(nma_handle_s* h, int i, const float* params, int p, float* gp, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    const int Fd = h->feat_out[i];
    const int n1 = (NMA_C + 1) * Fd;
    const int nb1 = (n1 + 127) / 128;
    int wsl = (2 * h->sm_count + nb1 - 1) / nb1;
    if (wsl > h->LW / 16) wsl = h->LW / 16;
    if (wsl < 1) wsl = 1;
    k_lv_feat4_wgrad<<<dim3(nb1, wsl), 128, 0, st>>>(h->ws[i].df, h->ws[i].a[4], h->ws[i].a[3], gp + h->po[i].featw[3],
                                                       gp + h->po[i].featb[3], Fd, d.LP, h->LW, h->LWP, p);
    const long long warps = (long long)p * NMA_C * h->LWP;
    k_lv_feat4_dgrad<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(h->ws[i].df, h->ws[i].a[4],
                                                                           params + h->po[i].featw[3], h->ws[i].df3, Fd,
                                                                           d.LP, h->LW, h->LWP, p);
    nma_count_launch(2);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}
