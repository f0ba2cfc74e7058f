// tcgen05 (5th-gen tensor core) form of the K-tap moving-average conv (A3, AR.py:61-62) and of its
// data gradient, each fused with its epilogue.  See nma_tc.cuh for the operand layouts.
//   k_tc_pack_w      : conv kernel [K][51][50] -> per-tap UMMA B tiles [K][14][64 hi rows | 64 lo rows][4] (fwd / flipped+transposed)
//   k_conv_fwd_tc    : conv + theta-bias + ELU (TMEM -> smem tile) + flow_epilogue (hidden 1x1, head, affine update)
//   k_conv_dgrad_tc  : full correlation of dA with the flipped kernel -> df (feature channels) and dx (channel 0)
//   nma_tc_conv_raw  : test hook, the bare contraction on caller-provided data
#include <stdlib.h>
#include "nma_tc.cuh"
#include "nma_flow_epi.cuh"

#define TC_THREADS 256

// main accumulator + correction accumulator (64 columns to the right), 32 columns of one warp's 32 lanes
__device__ __forceinline__ void tc_load_sum32(uint32_t taddr, float (&v)[32]) {
    float c[32];
    tmem_ld32(taddr, v);
    tmem_ld32(taddr + TC_N, c);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += c[i];
}

// ---------------------------------------------------------------------------
// weight packing.  mode 0 (forward): B[k][c][n] = W[k][c][n], c < 51 reduction, n < 50 outputs.
//                  mode 1 (dgrad):   B[k'][f][n] = W[K-1-k'][n][f], f < 50 reduction, n < 51 outputs.
// ---------------------------------------------------------------------------
__global__ void k_tc_pack_w(const float* __restrict__ W, int K, int mode, float* __restrict__ out) {
    const int n_half = K * TC_WHALF;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_half; t += gridDim.x * blockDim.x) {
        const int e = t & 3;
        const int n = (t >> 2) % TC_N;
        const int cch = (t / (4 * TC_N)) % TC_CCH;
        const int k = t / TC_WHALF;
        const int c = 4 * cch + e;
        float v = 0.f;
        if (mode == 0) {
            if (c < NMA_C1 && n < NMA_C) v = W[((size_t)k * NMA_C1 + c) * NMA_C + n];
        } else {
            if (c < NMA_C && n < NMA_C1) v = W[((size_t)(K - 1 - k) * NMA_C1 + n) * NMA_C + c];
        }
        const float hi = tf32_hi(v);
        const size_t o = (size_t)k * TC_WSTAGE + ((size_t)cch * TC_WROWS + n) * 4 + e;     // [k][cch][128 rows][4]
        out[o] = hi;
        out[o + TC_N * 4] = v - hi;                                                         // lo rows follow the hi rows
    }
}

// the same two packs in the 2-term bf16 split: [K][8][64 hi rows | 64 lo rows][8 x bf16]
// interleave = 1: hi and lo part of output n in rows 2n, 2n+1 (the tile is then the stacked A operand of the
// position-wide kernels, nma_tc_conv2.cu) instead of rows n, 64 + n
__global__ void k_tc_pack_w_bf(const float* __restrict__ W, int K, int mode, uint16_t* __restrict__ out, int interleave) {
    const int n_half = K * 8 * TC_N * 8;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_half; t += gridDim.x * blockDim.x) {
        const int e = t & 7;
        const int n = (t >> 3) & (TC_N - 1);
        const int cch = (t >> 9) & 7;
        const int k = t >> 12;
        const int c = 8 * cch + e;
        float v = 0.f;
        if (mode == 0) {
            if (c < NMA_C1 && n < NMA_C) v = W[((size_t)k * NMA_C1 + c) * NMA_C + n];
        } else {
            if (c < NMA_C && n < NMA_C1) v = W[((size_t)(K - 1 - k) * NMA_C1 + n) * NMA_C + c];
        }
        uint32_t hi, lo;
        bf_split(v, hi, lo);
        const size_t tile = (size_t)k * (8 * TC_WROWS * 8) + (size_t)cch * TC_WROWS * 8 + e;
        out[tile + (size_t)(interleave ? 2 * n : n) * 8] = (uint16_t)hi;
        out[tile + (size_t)(interleave ? 2 * n + 1 : n + TC_N) * 8] = (uint16_t)lo;
    }
}

// tap pairs (nma_tc_conv2.cu): one tile per pair, chunks [tap a: 0-5][tap b: 0-5][a: 6][b: 6]; a tap >= K packs zeros
__global__ void k_tc_pack_w_bf_pair(const float* __restrict__ W, int K, int mode, uint16_t* __restrict__ out, int interleave) {
    const int npairs = (K + 1) / 2;
    const int n_half = npairs * 14 * TC_N * 8;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_half; t += gridDim.x * blockDim.x) {
        const int e = t & 7;
        const int n = (t >> 3) & (TC_N - 1);
        const int u = (t >> 9) % 14;
        const int pr = t / (14 * TC_N * 8);
        const int k = 2 * pr + (u < 6 ? 0 : u < 12 ? 1 : u - 12);
        const int cch = u < 6 ? u : u < 12 ? u - 6 : 6;
        const int c = 8 * cch + e;
        float v = 0.f;
        if (k < K) {
            if (mode == 0) {
                if (c < NMA_C1 && n < NMA_C) v = W[((size_t)k * NMA_C1 + c) * NMA_C + n];
            } else {
                if (c < NMA_C && n < NMA_C1) v = W[((size_t)(K - 1 - k) * NMA_C1 + n) * NMA_C + c];
            }
        }
        uint32_t hi, lo;
        bf_split(v, hi, lo);
        const size_t tile = (size_t)pr * (14 * TC_WROWS * 8) + (size_t)u * TC_WROWS * 8 + e;
        out[tile + (size_t)(interleave ? 2 * n : n) * 8] = (uint16_t)hi;
        out[tile + (size_t)(interleave ? 2 * n + 1 : n + TC_N) * 8] = (uint16_t)lo;
    }
}

int launch_pack_weights_tc(nma_handle_s* h, const float* params, bool need_bwd, cudaStream_t st, int which) {
    for (int i = 0; (which & 1) && i < h->cfg.F; ++i) {
        if (h->use_bf16 && h->tap_pairs) {
            k_tc_pack_w_bf_pair<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, 0, (uint16_t*)h->ws[i].wtc_f, 0);
            if (need_bwd) k_tc_pack_w_bf_pair<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, 1, (uint16_t*)h->ws[i].wtc_d, 1);
            nma_count_launch(need_bwd ? 2 : 1);
            continue;
        }
        if (h->use_bf16) {
            k_tc_pack_w_bf<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, 0, (uint16_t*)h->ws[i].wtc_f, 0);
            if (need_bwd) k_tc_pack_w_bf<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, 1, (uint16_t*)h->ws[i].wtc_d, h->dgrad_wide);
            nma_count_launch(need_bwd ? 2 : 1);
            continue;
        }
        k_tc_pack_w<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, 0, h->ws[i].wtc_f);
        nma_count_launch(1);
        if (need_bwd) {
            k_tc_pack_w<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, 1, h->ws[i].wtc_d);
            nma_count_launch(1);
        }
    }
    NMA_CHECK_CUDA(cudaGetLastError());
    if (which & 1) {
        // the hidden 1x1 kernels of the tensor-core epilogues (forward: conv_fwd_tcp; backward: epi_bwd_tc)
        int rc;
        for (int i = 0; i < h->cfg.F; ++i) {
            if (conv_fwd_tcp_supported(h) && (rc = launch_pack_w1x1_fwd(h, i, params, st))) return rc;
            if (need_bwd && epi_bwd_tc_supported(h) && (rc = launch_pack_w1x1_t(h, i, params, st))) return rc;
        }
    }
    if ((which & 2) && h->use_tc_feat) return launch_pack_feat_tc(h, params, need_bwd, st);
    return 0;
}


// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
struct ConvFwdTcArgs {
    TcConvSrc src;
    const float* tb;         // [p][3][50]; slot 2 = theta bias + conv bias
    FlowEpiArgs e;
    int Lin, p, npos;
};

template <int NACC>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_fwd_tc(ConvFwdTcArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bars[TC_NBARS];
    __shared__ uint32_t tmem_slot;
    constexpr int NCOLS = NACC * TC_M;
    constexpr int PITCH = NCOLS + 4;
    __shared__ int col_r[NCOLS];
    __shared__ int col_m[NCOLS];
    __shared__ unsigned char col_ok[NCOLS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TcConvSmem s = tc_conv_carve(smem, a.npos, bars, &tmem_slot);
    const uint32_t tmem = tc_conv_setup(s, NACC * 2 * TC_N);
    const long long q0 = (long long)blockIdx.x * NCOLS;

    for (int col = tid; col < NCOLS; col += blockDim.x) {
        const long long q = q0 + col;
        const int r = (int)(q / a.Lin);
        const int m = (int)(q - (long long)r * a.Lin);
        col_r[col] = r < a.p ? r : 0;
        col_m[col] = m;
        col_ok[col] = (r < a.p && m < a.e.N) ? 1 : 0;
    }

    tc_conv_mainloop<NACC>(s, a.src, q0, a.npos, tmem);
    __syncthreads();

    // e_0 = elu(A + theta-bias + conv bias)  (AR.py:70-72): TMEM lane = position, column = output channel
    float* tile = smem;                               // [50][PITCH], aliases the (now idle) operand buffers
    if (warp < 4 * NACC) {
        const int acc = warp >> 2, quarter = warp & 3;
        const int col = acc * TC_M + quarter * 32 + lane;
        const bool ok = col_ok[col] != 0;
        const float* tbr = a.tb + ((size_t)col_r[col] * 3 + 2) * NMA_C;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tc_load_sum32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * TC_N + half * 32), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int f = half * 32 + i;
                if (f < NMA_C) tile[(size_t)f * PITCH + col] = ok ? elu_f(v[i] + tbr[f]) : 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    flow_epilogue<NCOLS>(a.e, tile, col_r, col_m, col_ok);
    tc_conv_teardown(tmem, NACC * 2 * TC_N);
}

static int tc_smem_bytes(int nacc, int K, size_t epi_floats) {
    size_t f = tc_conv_smem_floats(nacc, K);
    if (f < epi_floats) f = epi_floats;
    return (int)(f * 4);
}

int launch_conv_fwd_tc(nma_handle_s* h, int i, const float* params, int p, bool save, cudaStream_t st) {
    if (conv_fwd_tcp_supported(h)) return launch_conv_fwd_tcp(h, i, params, p, save, st);
    const FlowDims& d = h->fd[i];
    ConvFwdTcArgs a;
    const int nacc = h->tc_nacc;
    a.src.a_hi = h->ws[i].tin_hi; a.src.a_lo = h->ws[i].tin_lo; a.src.Qalloc = h->ws[i].tin_Q;
    a.src.wt = h->ws[i].wtc_f; a.src.K = h->cfg.K; a.src.diag = 0;
    a.tb = h->ws[i].tb;
    fill_flow_epi_args(h, i, params, save, a.e);
    a.Lin = d.Lin; a.p = p; a.npos = tc_conv_npos(nacc, h->cfg.K);
    const long long qtot = (long long)p * d.Lin;
    const int ncols = nacc * TC_M;
    const unsigned grid = (unsigned)((qtot + ncols - 1) / ncols);
    if (nacc == 2) {
        const int smem = tc_smem_bytes(2, h->cfg.K, flow_epi_smem_floats<2 * TC_M>());
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_fwd_tc<2><<<grid, TC_THREADS, smem, st>>>(a);
    } else {
        const int smem = tc_smem_bytes(1, h->cfg.K, flow_epi_smem_floats<TC_M>());
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_fwd_tc<1><<<grid, TC_THREADS, smem, st>>>(a);
    }
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// data gradient: dinp[q][c] = sum_k' sum_f dA_flat[q + k'][f] * W[K-1-k'][c][f]   (dA_flat has a K-1 lead pad and
// exact zeros in the K-1 slots between rows, so the flattened causal conv never mixes rows)
//   c >= 1 -> df[r][c-1][j];  c == 0 -> dx[r][j] += (flows > 0)
// ---------------------------------------------------------------------------
struct ConvDgradTcArgs {
    TcConvSrc src;
    float* df;           // [p][50][LP]
    float* dx;           // [p][XP]
    int Lin, LP, XP, p, npos, need_dx;
};

template <int NACC>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_dgrad_tc(ConvDgradTcArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bars[TC_NBARS];
    __shared__ uint32_t tmem_slot;
    constexpr int NCOLS = NACC * TC_M;
    constexpr int PITCH = NCOLS + 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TcConvSmem s = tc_conv_carve(smem, a.npos, bars, &tmem_slot);
    const uint32_t tmem = tc_conv_setup(s, NACC * 2 * TC_N);
    const long long q0 = (long long)blockIdx.x * NCOLS;

    tc_conv_mainloop<NACC>(s, a.src, q0, a.npos, tmem);
    __syncthreads();

    float* tile = smem;                               // [51][PITCH]: row n = input channel n
    if (warp < 4 * NACC) {
        const int acc = warp >> 2, quarter = warp & 3;
        const int col = acc * TC_M + quarter * 32 + lane;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tc_load_sum32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * TC_N + half * 32), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = half * 32 + i;
                if (n < NMA_C1) tile[(size_t)n * PITCH + col] = v[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    // column -> (row, slot) once per column (no 64-bit division in the store loop)
    int* col_r = reinterpret_cast<int*>(tile + (size_t)NMA_C1 * PITCH);
    int* col_j = col_r + NCOLS;
    const long long qtot = (long long)a.p * a.Lin;
    for (int col = tid; col < NCOLS; col += blockDim.x) {
        const long long q = q0 + col;
        const int r = (int)(q / a.Lin);
        col_r[col] = q < qtot ? r : -1;
        col_j[col] = (int)(q - (long long)r * a.Lin);
    }
    __syncthreads();
    for (int t = tid; t < NMA_C1 * NCOLS; t += blockDim.x) {
        const int n = t / NCOLS, col = t - n * NCOLS;
        const int r = col_r[col], j = col_j[col];
        if (r < 0) continue;
        const float v = tile[(size_t)n * PITCH + col];
        if (n >= 1) a.df[((size_t)r * NMA_C + (n - 1)) * a.LP + j] = v;
        else if (a.need_dx) a.dx[(size_t)r * a.XP + j] += v;
    }
    tc_conv_teardown(tmem, NACC * 2 * TC_N);
}

int launch_conv_dgrad_tc(nma_handle_s* h, int i, int p, cudaStream_t st) {
    if (h->use_tc_persist && h->tc_nacc == 2) return launch_conv_dgrad_tcp(h, i, p, st);
    const FlowDims& d = h->fd[i];
    ConvDgradTcArgs a;
    const int nacc = h->tc_nacc;
    a.src.a_hi = h->ws[i].dat_hi; a.src.a_lo = h->ws[i].dat_lo; a.src.Qalloc = h->ws[i].dat_Q;
    a.src.wt = h->ws[i].wtc_d; a.src.K = h->cfg.K; a.src.diag = 0;
    a.df = h->ws[i].df; a.dx = h->ws[i].dx;
    a.Lin = d.Lin; a.LP = d.LP; a.XP = (d.L + 3) & ~3; a.p = p; a.npos = tc_conv_npos(nacc, h->cfg.K);
    a.need_dx = i > 0 ? 1 : 0;
    const long long qtot = (long long)p * d.Lin;
    const int ncols = nacc * TC_M;
    const unsigned grid = (unsigned)((qtot + ncols - 1) / ncols);
    if (nacc == 2) {
        const int smem = tc_smem_bytes(2, h->cfg.K, (size_t)NMA_C1 * (2 * TC_M + 4) + 4 * TC_M);
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_dgrad_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_dgrad_tc<2><<<grid, TC_THREADS, smem, st>>>(a);
    } else {
        const int smem = tc_smem_bytes(1, h->cfg.K, (size_t)NMA_C1 * (TC_M + 4) + 2 * TC_M);
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_dgrad_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_dgrad_tc<1><<<grid, TC_THREADS, smem, st>>>(a);
    }
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// test hook: the bare contraction.  d_in [Q][56] fp32 (channel-last), d_w [K][51][50] (TF conv1d layout),
// d_out [Q][64]:  mode 0: out[q][n] = sum_k sum_c in[q+k][c] W[k][c][n]
//                 mode 1: out[q][n] = sum_k sum_f in[q+k][f] W[K-1-k][n][f]
// rows q >= Q-K+1 read the zero tail.  Allocates its own scratch (not a product path).
// ---------------------------------------------------------------------------
__global__ void k_tc_split_in(const float* __restrict__ in, long long Q, long long Qalloc, float* __restrict__ hi,
                              float* __restrict__ lo) {
    const long long n = Q * TC_CCH;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long q = t / TC_CCH;
        const int cch = (int)(t - q * TC_CCH);
        const float4 v = *reinterpret_cast<const float4*>(in + q * 56 + 4 * cch);
        const float4 h4 = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const size_t o = ((size_t)cch * Qalloc + q) * 4;
        *reinterpret_cast<float4*>(hi + o) = h4;
        *reinterpret_cast<float4*>(lo + o) = make_float4(v.x - h4.x, v.y - h4.y, v.z - h4.z, v.w - h4.w);
    }
}

// channel-last fp32 [Q][56] -> the bf16 split operand [8][Qalloc][8 x bf16] (hi and lo)
__global__ void k_tc_split_in_bf(const float* __restrict__ in, long long Q, long long Qalloc, uint4* __restrict__ hi,
                                 uint4* __restrict__ lo) {
    const long long n = Q * 7;                  // chunk 7 (channels 56..63) stays zero
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long q = t / 7;
        const int cch = (int)(t - q * 7);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = in[q * 56 + 8 * cch + e];
        uint4 h4, l4;
        bf_split8(v, h4, l4);
        hi[(size_t)cch * Qalloc + q] = h4;
        lo[(size_t)cch * Qalloc + q] = l4;
    }
}

template <int NACC, bool BF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_conv_raw(TcConvSrc src, int npos, long long Q, float* __restrict__ out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bars[TC_NBARS];
    __shared__ uint32_t tmem_slot;
    constexpr int NCOLS = NACC * TC_M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TcConvSmem s = tc_conv_carve(smem, npos, bars, &tmem_slot, TcP<BF>::CCH);
    const uint32_t tmem = tc_conv_setup(s, NACC * 2 * TC_N);
    const long long q0 = (long long)blockIdx.x * NCOLS;
    tc_conv_mainloop<NACC, BF>(s, src, q0, npos, tmem);
    if (warp < 4 * NACC) {
        const int acc = warp >> 2, quarter = warp & 3;
        const long long q = q0 + acc * TC_M + quarter * 32 + lane;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tc_load_sum32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * TC_N + half * 32), v);
            if (q < Q)
#pragma unroll
                for (int i = 0; i < 32; ++i) out[q * TC_N + half * 32 + i] = v[i];
        }
    }
    tc_conv_teardown(tmem, NACC * 2 * TC_N);
}

template <int NACC, bool BF>
static cudaError_t raw_launch(const TcConvSrc& src, int npos, long long Q, float* d_out, int smem, unsigned grid, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_conv_raw<NACC, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) k_tc_conv_raw<NACC, BF><<<grid, TC_THREADS, smem, st>>>(src, npos, Q, d_out);
    return e;
}

// mode bit 0: 0 forward / 1 data-gradient orientation of the kernel; mode bit 1: operand format, 0 3xTF32 / 1 bf16 split
extern "C" int nma_tc_conv_raw(const float* d_in, const float* d_w, int32_t mode, int32_t nacc, float* d_out, int64_t Q,
                               int32_t K, void* stream) {
    if (!d_in || !d_w || !d_out || Q < 1 || K < 1 || (nacc != 1 && nacc != 2) || mode < 0 || mode > 3) {
        nma_set_error("nma_tc_conv_raw: bad argument");
        return -1;
    }
    const bool bf = (mode & 2) != 0;
    mode &= 1;
    const int cch = bf ? TcP<true>::CCH : TcP<false>::CCH;
    cudaStream_t st = (cudaStream_t)stream;
    const int npos = tc_conv_npos(nacc, K);
    const int ncols = nacc * TC_M;
    const long long Qalloc = (Q + ncols - 1) / ncols * ncols + npos;
    float *hi = nullptr, *lo = nullptr, *wt = nullptr;
    const size_t abytes = (size_t)cch * Qalloc * 16;
    NMA_CHECK_CUDA(cudaMalloc(&hi, abytes));
    NMA_CHECK_CUDA(cudaMalloc(&lo, abytes));
    NMA_CHECK_CUDA(cudaMalloc(&wt, (size_t)K * TC_WSTAGE * 4));
    NMA_CHECK_CUDA(cudaMemsetAsync(hi, 0, abytes, st));
    NMA_CHECK_CUDA(cudaMemsetAsync(lo, 0, abytes, st));
    if (bf) {
        k_tc_split_in_bf<<<296, 256, 0, st>>>(d_in, Q, Qalloc, (uint4*)hi, (uint4*)lo);
        k_tc_pack_w_bf<<<148, 256, 0, st>>>(d_w, K, mode, (uint16_t*)wt, 0);
    } else {
        k_tc_split_in<<<296, 256, 0, st>>>(d_in, Q, Qalloc, hi, lo);
        k_tc_pack_w<<<148, 256, 0, st>>>(d_w, K, mode, wt);
    }
    TcConvSrc src;
    src.a_hi = hi; src.a_lo = lo; src.Qalloc = Qalloc; src.wt = wt; src.K = K; src.diag = 0;
    const int smem = (int)(tc_conv_smem_floats(nacc, K, cch) * 4);
    const unsigned grid = (unsigned)((Q + ncols - 1) / ncols);
    cudaError_t e;
    if (nacc == 2) e = bf ? raw_launch<2, true>(src, npos, Q, d_out, smem, grid, st) : raw_launch<2, false>(src, npos, Q, d_out, smem, grid, st);
    else e = bf ? raw_launch<1, true>(src, npos, Q, d_out, smem, grid, st) : raw_launch<1, false>(src, npos, Q, d_out, smem, grid, st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(hi); cudaFree(lo); cudaFree(wt);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        nma_set_error("nma_tc_conv_raw: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return -2;
    }
    return 0;
}


// ---------------------------------------------------------------------------
// weight gradient on the tensor cores:  dW[k][c][f] = sum_q inp_flat[q + k][c] * dA_flat[q][f]
//
// GEMM per pair of taps (t, t+1):  D[(j, c), f] += A[(j, c), q] * B[q, f]  with M = 2 taps x 64 channel slots,
// N = 64 and the flattened positions as the reduction.  tf32 operands must be K-major here (the no-swizzle MN-major
// form is not available for 32-bit types), i.e. 4 reduction indices per 16-byte unit.  The reduction order is free,
// so a unit holds 4 positions SPACED by 8:  unit(w)[c] = { inp[P + w + 8e][c] : e = 0..3 }.  With units stored
// [w][64 channel rows][16 B], tap t is again a start-address shift (w -> w + t, +1152 B), and the 128 M rows of a
// tap pair are simply two consecutive units.  Worker warps build the units from the same [channel/4][position][4]
// hi/lo tiles the forward pass reads (4x4 register transposes, conflict-free with an 8-row group stride of 144 B),
// so no extra layout lives in HBM.
//
// 3xTF32 as in the forward pass: B rows are [64 hi | 64 lo], one N=128 MMA with a_hi gives the main and the first
// correction sum, one N=64 MMA with a_lo the second.  The tensor core accumulates with truncation and this
// reduction is millions long, so every WGT_FLUSH stages (256 positions, a 32-MMA chain) the accumulators are drained
// TMEM -> registers and summed there in round-to-nearest fp32.
// CTA = (group of <= 4 tap pairs, range of stages); warps 0-7 transform + drain, warp 8 TMA, warp 9 MMA issue.
// ---------------------------------------------------------------------------
#define WGT_KT 32                    // positions per stage
#define WGT_AUNITS 16                // units of the A tile: 4 k-steps + 4 (second K chunk) + 7 (tap offsets) + 1
#define WGT_BUNITS 8
#define WGT_SBO 144                  // bytes between 8-row groups (128 + 16 pad: conflict-free transposes)
#define WGT_AUNIT_F (8 * WGT_SBO / 4)     // floats per A unit (64 rows)  = 288
#define WGT_BUNIT_F (16 * WGT_SBO / 4)    // floats per B unit (128 rows) = 576
#define WGT_ASRC 41                  // source positions of the A tile (40 needed; odd pitch)
#define WGT_BSRC 33
#define WGT_FLUSH 8
#define WGT_MAXPAIRS 4
#define WGT_THREADS 320
#define WGT_SRCA_F (TC_CCH * WGT_ASRC * 4)
#define WGT_SRCB_F (TC_CCH * WGT_BSRC * 4)
#define WGT_UA_F (WGT_AUNITS * WGT_AUNIT_F)
#define WGT_UB_F (WGT_BUNITS * WGT_BUNIT_F)
#define WGT_STAGE_F (2 * WGT_SRCA_F + 2 * WGT_SRCB_F + 2 * WGT_UA_F + WGT_UB_F)

struct ConvWgradTcArgs {
    const float* in_hi; const float* in_lo; long long in_Q;      // [14][in_Q][4]
    const float* da_hi; const float* da_lo; long long da_Q;      // [14][da_Q][4], dA(q) at index q + K - 1
    float* gW;                                                   // [K][51][50]
    int K, npairs, ngroups, nstages_total, nq;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(WGT_THREADS, 1) k_conv_wgrad_tc(ConvWgradTcArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t src_full[2], src_empty[2], unit_full[2], unit_empty[2], acc_full, acc_free;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;
    const int base = a.npairs / a.ngroups, rem = a.npairs % a.ngroups;
    const int np = base + (g < rem ? 1 : 0);                    // tap pairs of this CTA
    const int k0 = 2 * (g * base + (g < rem ? g : rem));        // first tap
    const int s_begin = (int)((long long)a.nstages_total * blockIdx.x / a.nq);
    const int s_end = (int)((long long)a.nstages_total * (blockIdx.x + 1) / a.nq);
    const int nst = s_end - s_begin;
    const int nchunks = (nst + WGT_FLUSH - 1) / WGT_FLUSH;

    // unit rows that no transform ever writes (channel slots 56..63) must hold finite values: zero everything once
    for (int t = tid; t < 2 * WGT_STAGE_F; t += blockDim.x) smem[t] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&src_full[i], 1); mbar_init(&src_empty[i], 8);
            mbar_init(&unit_full[i], 8); mbar_init(&unit_empty[i], 1);
        }
        mbar_init(&acc_full, 1); mbar_init(&acc_free, 8);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 8) {
        // ===== TMA producer =====
        for (int si = 0; si < nst; ++si) {
            const int st = si & 1;
            if (si >= 2) mbar_wait_lane0(&src_empty[st], (uint32_t)(((si >> 1) - 1) & 1));
            if (elect_one()) {
                float* sb = smem + (size_t)st * WGT_STAGE_F;
                const long long q0 = (long long)(s_begin + si) * WGT_KT;
                mbar_expect_tx(&src_full[st], (uint32_t)(2 * WGT_SRCA_F + 2 * WGT_SRCB_F) * 4u);
                for (int c = 0; c < TC_CCH; ++c) {
                    const size_t sa = ((size_t)c * a.in_Q + q0 + k0) * 4;
                    bulk_g2s(sb + (size_t)c * WGT_ASRC * 4, a.in_hi + sa, WGT_ASRC * 16u, &src_full[st]);
                    bulk_g2s(sb + WGT_SRCA_F + (size_t)c * WGT_ASRC * 4, a.in_lo + sa, WGT_ASRC * 16u, &src_full[st]);
                    const size_t sd = ((size_t)c * a.da_Q + q0 + (a.K - 1)) * 4;
                    bulk_g2s(sb + 2 * WGT_SRCA_F + (size_t)c * WGT_BSRC * 4, a.da_hi + sd, WGT_BSRC * 16u, &src_full[st]);
                    bulk_g2s(sb + 2 * WGT_SRCA_F + WGT_SRCB_F + (size_t)c * WGT_BSRC * 4, a.da_lo + sd, WGT_BSRC * 16u,
                             &src_full[st]);
                }
            }
            __syncwarp();
        }
    } else if (warp == 9) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc_n64 = umma_idesc_tf32(TC_M, TC_N, 0, 0);
        constexpr uint32_t idesc_n128 = umma_idesc_tf32(TC_M, 2 * TC_N, 0, 0);
        const uint32_t sbase = smem_u32(smem);
        const uint32_t hi32 = desc_hi(WGT_SBO);
        for (int si = 0; si < nst; ++si) {
            const int st = si & 1;
            const int ci = si / WGT_FLUSH;
            const bool chunk_first = (si % WGT_FLUSH) == 0;
            const bool chunk_last = ((si % WGT_FLUSH) == WGT_FLUSH - 1) || (si == nst - 1);
            if (chunk_first && ci > 0) mbar_wait_lane0(&acc_free, (uint32_t)((ci - 1) & 1));
            mbar_wait_lane0(&unit_full[st], (uint32_t)((si >> 1) & 1));
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ua_hi = sbase + (uint32_t)(st * WGT_STAGE_F + 2 * WGT_SRCA_F + 2 * WGT_SRCB_F) * 4u;
                const uint32_t ua_lo = ua_hi + WGT_UA_F * 4u;
                const uint32_t ub = ua_lo + WGT_UA_F * 4u;
                const uint32_t ah0 = desc_lo(ua_hi, 4u * WGT_AUNIT_F * 4u), al0 = desc_lo(ua_lo, 4u * WGT_AUNIT_F * 4u);
                const uint32_t b0 = desc_lo(ub, 4u * WGT_BUNIT_F * 4u);
                for (int pr = 0; pr < np; ++pr) {
                    const uint32_t d = tmem + (uint32_t)(pr * 2 * TC_N);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t aoff = (uint32_t)(u + 2 * pr) * (WGT_AUNIT_F * 4u / 16u);
                        const uint64_t ah = desc_pack(ah0 + aoff, hi32), al = desc_pack(al0 + aoff, hi32);
                        const uint64_t bw = desc_pack(b0 + (uint32_t)u * (WGT_BUNIT_F * 4u / 16u), hi32);
                        umma_tf32(d, ah, bw, idesc_n128, (chunk_first && u == 0) ? 0u : 1u);
                        umma_tf32(d + TC_N, al, bw, idesc_n64, 1u);
                    }
                }
                tc_commit(&unit_empty[st]);
                if (chunk_last) tc_commit(&acc_full);
            }
            __syncwarp();
        }
    } else if (warp < 8) {
        // ===== worker warps: build units, drain accumulators =====
        const int quarter = warp & 3, colhalf = warp >> 2;
        float acc[WGT_MAXPAIRS][32];
#pragma unroll
        for (int pr = 0; pr < WGT_MAXPAIRS; ++pr)
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[pr][i] = 0.f;

        auto drain = [&](int ci) {
            mbar_wait_lane0(&acc_full, (uint32_t)(ci & 1));
            tc_fence_after();
#pragma unroll
            for (int pr = 0; pr < WGT_MAXPAIRS; ++pr) {
                if (pr < np) {
                    float v[32], c2[32];
                    const uint32_t ta = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(pr * 2 * TC_N + colhalf * 32);
                    tmem_ld32(ta, v);
                    tmem_ld32(ta + TC_N, c2);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[pr][i] += v[i] + c2[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free);
        };

        const int oct = tid >> 3, l8 = tid & 7;           // 32 octets of 8 lanes; 96 octet-items per stage
        for (int si = 0; si < nst; ++si) {
            const int st = si & 1;
            mbar_wait_lane0(&src_full[st], (uint32_t)((si >> 1) & 1));
            if (si >= 2) mbar_wait_lane0(&unit_empty[st], (uint32_t)(((si >> 1) - 1) & 1));
            float* sb = smem + (size_t)st * WGT_STAGE_F;
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int o = oct + 32 * it;                 // 0..63: A (hl, w, grp); 64..95: B (hl, u, grp)
                const bool isA = o < 64;
                const int oo = isA ? o : o - 64;
                const int grp = oo & 1;
                const int w = isA ? ((oo >> 1) & 15) : ((oo >> 1) & 7);
                const int hl = isA ? (oo >> 5) : (oo >> 4);
                const int c4 = grp * 8 + l8;
                if (c4 < TC_CCH) {
                    const int spitch = isA ? WGT_ASRC : WGT_BSRC;
                    const float* src = sb + (isA ? hl * WGT_SRCA_F : 2 * WGT_SRCA_F + hl * WGT_SRCB_F) + ((size_t)c4 * spitch + w) * 4;
                    const float4 v0 = *reinterpret_cast<const float4*>(src);
                    const float4 v1 = *reinterpret_cast<const float4*>(src + 8 * 4);
                    const float4 v2 = *reinterpret_cast<const float4*>(src + 16 * 4);
                    const float4 v3 = *reinterpret_cast<const float4*>(src + 24 * 4);
                    // rows 4*c4 .. 4*c4+3 of the unit (B: + 64 for the lo half); group = row / 8, 36 floats per group
                    const int row0 = 4 * c4 + (isA ? 0 : hl * 64);
                    float* dst = sb + 2 * WGT_SRCA_F + 2 * WGT_SRCB_F +
                                 (isA ? hl * WGT_UA_F + w * WGT_AUNIT_F : 2 * WGT_UA_F + w * WGT_BUNIT_F) +
                                 (row0 >> 3) * (WGT_SBO / 4) + (row0 & 7) * 4;
                    *reinterpret_cast<float4*>(dst + 0) = make_float4(v0.x, v1.x, v2.x, v3.x);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(v0.y, v1.y, v2.y, v3.y);
                    *reinterpret_cast<float4*>(dst + 8) = make_float4(v0.z, v1.z, v2.z, v3.z);
                    *reinterpret_cast<float4*>(dst + 12) = make_float4(v0.w, v1.w, v2.w, v3.w);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&unit_full[st]); mbar_arrive(&src_empty[st]); }
            if (si > 0 && (si % WGT_FLUSH) == 0) drain(si / WGT_FLUSH - 1);
        }
        drain(nchunks - 1);

        // TMEM lane = M row = j*64 + channel slot
        const int M = quarter * 32 + lane, j = M >> 6, c = M & 63;
#pragma unroll
        for (int pr = 0; pr < WGT_MAXPAIRS; ++pr) {
            const int tap = k0 + 2 * pr + j;
            if (pr < np && tap < a.K && c < NMA_C1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int f = colhalf * 32 + i;
                    if (f < NMA_C) atomicAdd(a.gW + ((size_t)tap * NMA_C1 + c) * NMA_C + f, acc[pr][i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static void wgrad_tc_geometry(ConvWgradTcArgs& a, long long qtot, int sm_count) {
    a.npairs = (a.K + 1) / 2;
    a.ngroups = (a.npairs + WGT_MAXPAIRS - 1) / WGT_MAXPAIRS;
    a.nstages_total = (int)((qtot + WGT_KT - 1) / WGT_KT);
    int nq = (3 * sm_count) / a.ngroups;
    if (nq < 1) nq = 1;
    if (nq > a.nstages_total) nq = a.nstages_total;
    a.nq = nq;
}

int launch_conv_wgrad_tc(nma_handle_s* h, int i, int p, float* gp, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    ConvWgradTcArgs a;
    a.in_hi = h->ws[i].tin_hi; a.in_lo = h->ws[i].tin_lo; a.in_Q = h->ws[i].tin_Q;
    a.da_hi = h->ws[i].dat_hi; a.da_lo = h->ws[i].dat_lo; a.da_Q = h->ws[i].dat_Q;
    a.gW = gp + h->po[i].convw;
    a.K = h->cfg.K;
    wgrad_tc_geometry(a, (long long)p * d.Lin, h->sm_count);
    const int smem = 2 * WGT_STAGE_F * 4;
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    k_conv_wgrad_tc<<<dim3(a.nq, a.ngroups), WGT_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// test hook: bare tensor-core weight gradient.  d_in, d_da [Q][56] (channel-last fp32), d_gw [K][51][50]
// (accumulated into; the caller zeroes it):  gw[k][c][f] += sum_{q <= Q-K} in[q+k][c] * da[q][f]
extern "C" int nma_tc_wgrad_raw(const float* d_in, const float* d_da, float* d_gw, int64_t Q, int32_t K, void* stream) {
    if (!d_in || !d_da || !d_gw || Q < K || K < 1) { nma_set_error("nma_tc_wgrad_raw: bad argument"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    const long long Qalloc = (Q + 255) / 256 * 256 + 640 + K;
    const size_t abytes = (size_t)TC_CCH * Qalloc * 16;
    float *ih = nullptr, *il = nullptr, *dh = nullptr, *dl = nullptr;
    NMA_CHECK_CUDA(cudaMalloc(&ih, abytes)); NMA_CHECK_CUDA(cudaMalloc(&il, abytes));
    NMA_CHECK_CUDA(cudaMalloc(&dh, abytes)); NMA_CHECK_CUDA(cudaMalloc(&dl, abytes));
    NMA_CHECK_CUDA(cudaMemsetAsync(ih, 0, abytes, st)); NMA_CHECK_CUDA(cudaMemsetAsync(il, 0, abytes, st));
    NMA_CHECK_CUDA(cudaMemsetAsync(dh, 0, abytes, st)); NMA_CHECK_CUDA(cudaMemsetAsync(dl, 0, abytes, st));
    k_tc_split_in<<<296, 256, 0, st>>>(d_in, Q, Qalloc, ih, il);
    // dA(q) lives at index q + K - 1; only q <= Q-K contribute
    k_tc_split_in<<<296, 256, 0, st>>>(d_da, Q - K + 1, Qalloc, dh + (size_t)(K - 1) * 4, dl + (size_t)(K - 1) * 4);
    ConvWgradTcArgs a;
    a.in_hi = ih; a.in_lo = il; a.in_Q = Qalloc; a.da_hi = dh; a.da_lo = dl; a.da_Q = Qalloc;
    a.gW = d_gw; a.K = K;
    wgrad_tc_geometry(a, Q, 4);
    const int smem = 2 * WGT_STAGE_F * 4;
    cudaError_t e = cudaFuncSetAttribute(k_conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) k_conv_wgrad_tc<<<dim3(a.nq, a.ngroups), WGT_THREADS, smem, st>>>(a);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(ih); cudaFree(il); cudaFree(dh); cudaFree(dl);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        nma_set_error("nma_tc_wgrad_raw: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return -2;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// weight gradient in the bf16 split:  dW[k][c][f] = sum_q inp_flat[q + k][c] * dA_flat[q][f]   on kind::f16,
// with dA as a TMEM-RESIDENT A operand (the "TS" form of tcgen05.mma).
//
// 16-bit operands may be MN-major, and the conv operand layout [channel/8][position][8 x bf16] IS the no-swizzle
// MN-major canonical layout of a (channels x positions) matrix with the positions as the reduction: 8 consecutive
// positions of one channel chunk are one 128-byte core matrix, the next 8 positions follow at +128 B (leading byte
// offset), the next channel chunk at the slab stride (stride byte offset).  So the `in` operand is fed to the tensor
// core exactly as the forward pass wrote it - the TMA engine copies slabs - and tap t is a +16t-byte start address.
// dA is the operand every tap of a CTA shares, so it is written ONCE per stage into tensor memory - transposed to the
// K-major form the A-from-TMEM path requires, lane = channel (64 hi rows over 64 lo rows), column = two consecutive
// positions - and every MMA then fetches only its 2 KB B operand from shared memory and runs at the pipe's own rate
// (32 cycles for M = 128, N = 64, K = 16; 48 with A in shared memory: tools/micro/mma_rate.cu):
//     D[tap][(hl, f), c] += dA_t[(hl, f), q] * in[q + tap][c]          M = 128, N = 64, K = 16 positions
// B is the `in` slab of ONE tap (hi part, then lo part: two instructions into the same accumulator), MN-major straight
// from the conv operand layout as before; no second shifted copy of the slabs, 30 KB instead of 53 KB per stage.  All
// four products of the split are accumulated (hi.hi + hi.lo in the hi lanes, lo.hi + lo.lo in the lo lanes; the
// epilogue adds the two lanes of a channel through the atomics).
// CTA = (group of <= 4 taps, range of stages); warps 0-7 transpose dA into TMEM AND drain the accumulators (they run
// WS_LAG stages ahead of the drain), warp 8 TMA, warp 9 MMA.  Accumulators are drained tap by tap (own barrier per
// tap), so the drain of tap t overlaps the MMAs of the other taps.
// ---------------------------------------------------------------------------
#define WS_KT 64
#define WS_TAPS 4
#define WS_APOS (WS_KT + 8)                 // `in` positions per stage: tap offsets 0..3, padded to a multiple of 8
#define WS_STAGES 6
#define WS_LAG 3
#define WS_FLUSH 32                         // stages per drain: 32 x 4 k-steps x 2 instructions = 256-MMA chains (raw error 5.5e-6 of the largest entry, 3.5e-6 at 8: profiles/r02_wgrad_ts.md)
#define WS_IN_UNITS (8 * WS_APOS)           // 16-byte units of in_hi (or in_lo): 8 chunk slabs
#define WS_DA_SLAB (WS_KT + 1)              // +1 unit: the transposing 2-byte reads of a warp hit 16 distinct banks
#define WS_DA_UNITS (16 * WS_DA_SLAB)
#define WS_STAGE_UNITS (2 * WS_IN_UNITS + WS_DA_UNITS)
#define WS_NCOPY 28
#define WS_DCOLS (WS_TAPS * TC_N)           // accumulator columns; the A buffers follow
#define WS_MMA_WARPS 2                      // issuing warps (taps split between them): one warp's instruction stream is
                                            // the limit otherwise (tools/micro/mma_rate.cu: the pipe itself takes an N = 64
                                            // A-from-TMEM instruction every 32 cycles)
#define WS_THREADS (32 * (10 + WS_MMA_WARPS))   // warps 0-7 transpose + drain, 8 and 11 TMA, 9-10 MMA

struct ConvWgradTsArgs {
    const uint4* in_hi; const uint4* in_lo; long long in_Q;
    const uint4* da_hi; const uint4* da_lo; long long da_Q;
    float* gW;
    int K, ngroups, nstages_total, nq, spr, flush;
    long long Lin, Nv;
    int diag;      // NMA_DIAG timing experiments (results invalid): 1 no TMA copies, 2 no transposition, 4 no MMAs, 8 no TMEM drain loads
};

// Barrier helpers on precomputed 32-bit shared addresses.  Taking `&bar[i]` of a static __shared__ array inside a loop makes
// the compiler rebuild the address from SR_CgaCtaId every time (S2UR + ULEA, ~100 cycles of latency each): with ~6 barrier
// operations per stage and warp that alone paced k_conv_wgrad_ts at ~760 cycles per stage with every other piece of work
// switched off (profiles/r02_wgrad_ts.md).  The addresses are formed once per kernel and hidden from rematerialisation.
__device__ __forceinline__ uint32_t smem_addr_once(const void* p) {
    uint32_t a = smem_u32(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_commit_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
// mbarrier wait without the clock reads of mbar_wait_backoff (hot per-stage waits); the watchdog is a poll counter:
// a lost arrival must fail loudly, not hang the device
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    // every lane polls: one polling lane + __syncwarp was measured 1.9x SLOWER on the whole kernel (round 2; round 1 saw
    // the same on the forward kernel).  The suspend-time hint lets a waiting warp sleep in the barrier unit.
    uint32_t done = 0, polls = 0;
    while (!done) {
        if (++polls > (1u << 27)) __trap();
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
    }
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(WS_THREADS, 1) k_conv_wgrad_ts(ConvWgradTsArgs a) {
    extern __shared__ __align__(128) uint4 smem_u[];
    __shared__ uint64_t full[WS_STAGES], empty[WS_STAGES], a_ready[WS_STAGES], acc_full[WS_TAPS], acc_free[WS_TAPS];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // tap group fastest in launch order: the groups of one range of stages run side by side and share its operands
    // through L2 (with the range fastest every wave of CTAs swept the operands again: 2.4x the operand bytes from DRAM)
    const int g = blockIdx.x, rng = blockIdx.y;
    const int base = a.K / a.ngroups, rem = a.K % a.ngroups;
    const int nt = base + (g < rem ? 1 : 0);                    // taps of this CTA
    const int k0 = g * base + (g < rem ? g : rem);              // first tap
    const int s_begin = (int)((long long)a.nstages_total * rng / a.nq);
    const int s_end = (int)((long long)a.nstages_total * (rng + 1) / a.nq);
    const int nst = s_end - s_begin;

    // slabs of chunk 7 (channel slots 56..63) are never loaded: zero everything once
    for (int t = tid; t < WS_STAGES * WS_STAGE_UNITS; t += blockDim.x) smem_u[t] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int i = 0; i < WS_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&a_ready[i], 4); }
        for (int i = 0; i < WS_TAPS; ++i) { mbar_init(&acc_full[i], WS_MMA_WARPS); mbar_init(&acc_free[i], 8); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp < 8) {
        // every MMA accumulates (the two issuing warps are not ordered against each other, so none of them may be the one
        // that overwrites): the tiles start at zero and the drain leaves them at zero
        uint32_t z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0u;
#pragma unroll
        for (int t = 0; t < WS_TAPS; ++t)
            tmem_st32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * TC_N + (warp >> 2) * 32), z);
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    const uint32_t full_a = smem_addr_once(full), empty_a = smem_addr_once(empty), aready_a = smem_addr_once(a_ready);
    const uint32_t accfull_a = smem_addr_once(acc_full), accfree_a = smem_addr_once(acc_free);
    const uint32_t sbase = smem_addr_once(smem_u);

    if (warp == 8 || warp == 9 + WS_MMA_WARPS) {
        // ===== TMA producers: two warps, each takes every other stage (a warp's chain of barrier wait, address arithmetic and
        // 28 copy instructions per stage is ~400 cycles long: one warp alone could not keep up with the 850-cycle stages) =====
        const int pw = warp == 8 ? 0 : 1;
        long long row = (s_begin + pw) / a.spr;
        int blk = (int)((s_begin + pw) - row * a.spr), st = pw;
        uint32_t ph = 0;
        for (int si = pw; si < nst; si += 2) {
            if (si >= WS_STAGES) mbar_wait_spin(empty_a + 8u * st, ph ^ 1u);
            const uint32_t sb = sbase + (uint32_t)st * (WS_STAGE_UNITS * 16u), fb = full_a + 8u * st;
            const long long q0 = row * a.Lin + (long long)blk * WS_KT;
            const long long left = a.Nv - (long long)blk * WS_KT;     // the last stage of a row copies only what its k-steps read
            const uint32_t npos_b = left >= WS_KT ? (uint32_t)WS_KT : (uint32_t)((left + 15) / 16) * 16u;
            const uint32_t npos_a = npos_b + 8u;
            if (a.diag & 1) {
                if (elect_one()) mbar_expect_tx_a(fb, 0u);
            } else if (elect_one()) {
                mbar_expect_tx_a(fb, 14u * npos_a * 16u + 14u * npos_b * 16u);
                const uint4* ih = a.in_hi + q0 + k0;
                const uint4* il = a.in_lo + q0 + k0;
                const uint4* dh = a.da_hi + q0 + (a.K - 1);
                const uint4* dl = a.da_lo + q0 + (a.K - 1);
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    bulk_g2s_a(sb + (2 * WS_IN_UNITS + c * WS_DA_SLAB) * 16u, dh + (size_t)c * a.da_Q, npos_b * 16u, fb);
                    bulk_g2s_a(sb + (2 * WS_IN_UNITS + (8 + c) * WS_DA_SLAB) * 16u, dl + (size_t)c * a.da_Q, npos_b * 16u, fb);
                }
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    bulk_g2s_a(sb + (c * WS_APOS) * 16u, ih + (size_t)c * a.in_Q, npos_a * 16u, fb);
                    bulk_g2s_a(sb + (WS_IN_UNITS + c * WS_APOS) * 16u, il + (size_t)c * a.in_Q, npos_a * 16u, fb);
                }
            }
            __syncwarp();
            blk += 2;
            while (blk >= a.spr) { blk -= a.spr; ++row; }
            st += 2;
            if (st >= WS_STAGES) { st -= WS_STAGES; ph ^= 1u; }
        }
    } else if (warp >= 9 && warp < 9 + WS_MMA_WARPS) {
        // ===== MMA issuers: warp 9 + w issues ALL taps of the stages si = w (mod WS_MMA_WARPS) =====
        // ncu (profiles/r02_wgrad_ts.md): an issuing warp never waits for data (1.07 polls per barrier wait) and is busy
        // ~85 % of the time - its own serial chain per stage (two barrier checks, elect, descriptor set-up, 16-32
        // UTCHMMA, commits: ~1700 cycles) is what paced the kernel, not the tensor pipe (1024 cycles per stage).  Split
        // by STAGE, each warp has two stage times for that chain.  The two warps are not ordered against each other, so
        // no MMA overwrites: the tiles start at zero and the drain re-zeroes them.
        // Taps and k-steps are unrolled with compile-time descriptor offsets (a rolled loop with runtime offsets cost ~7
        // uniform-datapath instructions per MMA), and the stage's position inside its row is tracked incrementally.
        constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N, 0, 1);        // A from TMEM (K-major), B MN-major
        const uint32_t b_hi32 = desc_hi(WS_APOS * 16u);                        // stride offset: next channel chunk
        const int mw = warp - 9;
        static_assert(WS_FLUSH % WS_MMA_WARPS == 0 && WS_STAGES % WS_MMA_WARPS == 0, "chunks and ring passes hold whole rounds of the issuing warps");
        int blk = (int)(((long long)s_begin + mw) % a.spr);
        const int nks_last = (int)((a.Nv - (long long)(a.spr - 1) * WS_KT + 15) / 16);
        const int per_chunk = a.flush / WS_MMA_WARPS;                        // this warp's stages per drain chunk
        const int nchunks = (nst + a.flush - 1) / a.flush;
        int st = mw, in_chunk = 0, ci = 0;
        uint32_t ph = 0;                                                     // parity of full / a_ready for this ring pass
        for (int si = mw; si < nst; si += WS_MMA_WARPS) {
            const bool chunk_first = in_chunk == 0;
            const bool chunk_last = (in_chunk == per_chunk - 1) || (si + WS_MMA_WARPS >= nst);
            mbar_wait_spin(full_a + 8u * st, ph);
            mbar_wait_spin(aready_a + 8u * st, ph);
            tc_fence_after();
            const int nks = (blk == a.spr - 1) ? nks_last : WS_KT / 16;
            if (chunk_first && ci > 0) {            // the previous chunk's sums have been drained (and zeroed) from these tiles
#pragma unroll
                for (int t = 0; t < WS_TAPS; ++t)
                    if (t < nt) mbar_wait_spin(accfree_a + 8u * t, (uint32_t)((ci - 1) & 1));
                tc_fence_after();
            }
            if (elect_one()) {
                const uint32_t ub_hi = sbase + (uint32_t)(st * WS_STAGE_UNITS) * 16u;
                const uint32_t bh0 = desc_lo(ub_hi, 128u);                    // leading offset: next 8 positions
                const uint32_t bl0 = bh0 + WS_IN_UNITS;                       // lo slabs follow the hi slabs (16-byte units)
                const uint32_t ta = tmem + (uint32_t)(WS_DCOLS + st * (WS_KT / 2));
                // k-step-major, taps innermost: consecutive instructions accumulate into DIFFERENT tiles (an instruction
                // that accumulates into the tile of its predecessor waits for it: ~80 cycles, whatever the shape)
#pragma unroll
                for (int ks = 0; ks < WS_KT / 16; ++ks) {
                    if (ks < nks && !(a.diag & 4)) {
#pragma unroll
                        for (int t = 0; t < WS_TAPS; ++t)
                            if (t < nt)
                                umma_bf16_ts(tmem + (uint32_t)(t * TC_N), ta + (uint32_t)(8 * ks),
                                             desc_pack(bh0 + (uint32_t)(t + 16 * ks), b_hi32), idesc, 1u);
#pragma unroll
                        for (int t = 0; t < WS_TAPS; ++t)
                            if (t < nt)
                                umma_bf16_ts(tmem + (uint32_t)(t * TC_N), ta + (uint32_t)(8 * ks),
                                             desc_pack(bl0 + (uint32_t)(t + 16 * ks), b_hi32), idesc, 1u);
                    }
                }
                if (chunk_last) {
#pragma unroll
                    for (int t = 0; t < WS_TAPS; ++t)
                        if (t < nt) tc_commit_a(accfull_a + 8u * t);
                }
                tc_commit_a(empty_a + 8u * st);
            }
            __syncwarp();
            blk += WS_MMA_WARPS;
            while (blk >= a.spr) blk -= a.spr;
            st += WS_MMA_WARPS;
            if (st >= WS_STAGES) { st -= WS_STAGES; ph ^= 1u; }
            if (chunk_last) { in_chunk = 0; ++ci; } else ++in_chunk;
        }
        // a last chunk shorter than the number of issuing warps: the drain still expects this warp's arrival
        if (ci < nchunks) {
            if (ci > 0) {
#pragma unroll
                for (int t = 0; t < WS_TAPS; ++t)
                    if (t < nt) mbar_wait_spin(accfree_a + 8u * t, (uint32_t)((ci - 1) & 1));
            }
            if (elect_one()) {
#pragma unroll
                for (int t = 0; t < WS_TAPS; ++t)
                    if (t < nt) tc_commit_a(accfull_a + 8u * t);
            }
            __syncwarp();
        }
    } else {
        // ===== warps 0-7: dA -> TMEM (transposed), WS_LAG stages ahead of the accumulator drain =====
        // (ring slot, parity and the position inside the drain chunk are tracked incrementally: these warps' own
        // instruction stream - ~150 instructions per stage with runtime modulos - paced the kernel before)
        const int quarter = warp & 3, half = warp >> 2;
        const int m = quarter * 32 + lane, part = m >> 6, f = m & 63;
        float acc[WS_TAPS][32];
#pragma unroll
        for (int t = 0; t < WS_TAPS; ++t)
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[t][i] = 0.f;
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
        // group `half` (4 warps = all 128 lanes) transposes the stages of its parity, all 64 positions of them
        const uint16_t* src0 = reinterpret_cast<const uint16_t*>(smem_u + 2 * WS_IN_UNITS + (part * 8 + (f >> 3)) * WS_DA_SLAB) + (f & 7);
        int st = half, in_chunk = 0, ci = 0;
        uint32_t ph = 0;
        for (int si = 0; si < nst + WS_LAG; ++si) {
            if (si < nst && (si & 1) == half) {
                // the A buffer of this stage slot was read by the MMAs of stage si - WS_STAGES
                if (si >= WS_STAGES) mbar_wait_spin(empty_a + 8u * st, ph ^ 1u);
                mbar_wait_spin(full_a + 8u * st, ph);
                tc_fence_after();
                if (!(a.diag & 2)) {
                    const uint16_t* src = src0 + (size_t)st * (WS_STAGE_UNITS * 8);
                    uint32_t r[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        r[j] = (uint32_t)src[(2 * j) * 8] | ((uint32_t)src[(2 * j + 1) * 8] << 16);
                    tmem_st32(lane_addr + (uint32_t)(WS_DCOLS + st * (WS_KT / 2)), r);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(aready_a + 8u * st);
                st += 2;
                if (st >= WS_STAGES) { st -= WS_STAGES; ph ^= 1u; }
            }
            if (si >= WS_LAG) {
                const bool last = (in_chunk == a.flush - 1) || (si - WS_LAG == nst - 1);
                if (last) {
#pragma unroll
                    for (int t = 0; t < WS_TAPS; ++t) {
                        if (t < nt) {
                            mbar_wait_spin(accfull_a + 8u * t, (uint32_t)(ci & 1));
                            tc_fence_after();
                            if (!(a.diag & 8)) {
                                float v[32];
                                tmem_ld32(lane_addr + (uint32_t)(t * TC_N + half * 32), v);
#pragma unroll
                                for (int i = 0; i < 32; ++i) acc[t][i] += v[i];
                            }
                            {   // the next chunk accumulates onto zero
                                uint32_t z[32];
#pragma unroll
                                for (int i = 0; i < 32; ++i) z[i] = 0u;
                                tmem_st32(lane_addr + (uint32_t)(t * TC_N + half * 32), z);
                            }
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_a(accfree_a + 8u * t);
                        }
                    }
                    in_chunk = 0; ++ci;
                } else {
                    ++in_chunk;
                }
            }
        }
        // TMEM lane = (part, f): both parts of a channel add into the same weight gradient entry
#pragma unroll
        for (int t = 0; t < WS_TAPS; ++t) {
            const int tap = k0 + t;
            if (t < nt && tap < a.K && f < NMA_C) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = half * 32 + i;
                    if (c < NMA_C1) atomicAdd(a.gW + ((size_t)tap * NMA_C1 + c) * NMA_C + f, acc[t][i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static void wgrad_ts_geometry(ConvWgradTsArgs& a, long long rows, long long Lin, long long Nv, int sm_count) {
    a.ngroups = (a.K + WS_TAPS - 1) / WS_TAPS;
    // the last k-step of a row reads up to 15 positions past Nv: they must fall into the zero gap of K-1 slots
    if (rows > 1 && ((16 - Nv % 16) % 16) > a.K - 1) { Nv = Lin = rows * Lin; rows = 1; }
    a.Lin = Lin; a.Nv = Nv;
    a.spr = (int)((Nv + WS_KT - 1) / WS_KT);
    a.nstages_total = (int)(rows * a.spr);
    int nq = (3 * sm_count) / a.ngroups;
    if (nq < 1) nq = 1;
    if (nq > a.nstages_total) nq = a.nstages_total;
    a.nq = nq;
    a.flush = WS_FLUSH;
    { const char* e = getenv("NMA_WS_FLUSH"); if (e && atoi(e) >= 2) a.flush = atoi(e) & ~1; }   // timing experiment
    a.diag = nma_diag_bits();
}

static int launch_wgrad_ts(ConvWgradTsArgs& a, cudaStream_t st) {
    const int smem = WS_STAGES * WS_STAGE_UNITS * 16;
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_wgrad_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    k_conv_wgrad_ts<<<dim3(a.ngroups, a.nq), WS_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_conv_wgrad_bf(nma_handle_s* h, int i, int p, float* gp, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    ConvWgradTsArgs t;
    t.in_hi = (const uint4*)h->ws[i].tin_hi; t.in_lo = (const uint4*)h->ws[i].tin_lo; t.in_Q = h->ws[i].tin_Q;
    t.da_hi = (const uint4*)h->ws[i].dat_hi; t.da_lo = (const uint4*)h->ws[i].dat_lo; t.da_Q = h->ws[i].dat_Q;
    t.gW = gp + h->po[i].convw;
    t.K = h->cfg.K;
    wgrad_ts_geometry(t, p, d.Lin, d.N, h->sm_count);
    return launch_wgrad_ts(t, st);
}

// test hook: the bare bf16-split weight gradient, same contract as nma_tc_wgrad_raw
extern "C" int nma_tc_wgrad_raw_bf(const float* d_in, const float* d_da, float* d_gw, int64_t Q, int32_t K, void* stream) {
    if (!d_in || !d_da || !d_gw || Q < K || K < 1) { nma_set_error("nma_tc_wgrad_raw_bf: bad argument"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    const long long Qalloc = (Q + 255) / 256 * 256 + 640 + K;
    const size_t abytes = (size_t)8 * Qalloc * 16;
    uint4 *ih = nullptr, *il = nullptr, *dh = nullptr, *dl = nullptr;
    NMA_CHECK_CUDA(cudaMalloc(&ih, abytes)); NMA_CHECK_CUDA(cudaMalloc(&il, abytes));
    NMA_CHECK_CUDA(cudaMalloc(&dh, abytes)); NMA_CHECK_CUDA(cudaMalloc(&dl, abytes));
    NMA_CHECK_CUDA(cudaMemsetAsync(ih, 0, abytes, st)); NMA_CHECK_CUDA(cudaMemsetAsync(il, 0, abytes, st));
    NMA_CHECK_CUDA(cudaMemsetAsync(dh, 0, abytes, st)); NMA_CHECK_CUDA(cudaMemsetAsync(dl, 0, abytes, st));
    k_tc_split_in_bf<<<296, 256, 0, st>>>(d_in, Q, Qalloc, ih, il);
    // dA(q) lives at unit q + K - 1; only q <= Q-K contribute
    k_tc_split_in_bf<<<296, 256, 0, st>>>(d_da, Q - K + 1, Qalloc, dh + (K - 1), dl + (K - 1));
    cudaError_t e = cudaSuccess;
    ConvWgradTsArgs t;
    t.in_hi = ih; t.in_lo = il; t.in_Q = Qalloc; t.da_hi = dh; t.da_lo = dl; t.da_Q = Qalloc; t.gW = d_gw; t.K = K;
    wgrad_ts_geometry(t, 1, Q, Q, 4);
    if (launch_wgrad_ts(t, st)) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(ih); cudaFree(il); cudaFree(dh); cudaFree(dl);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        nma_set_error("nma_tc_wgrad_raw_bf: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return -2;
    }
    return 0;
}
