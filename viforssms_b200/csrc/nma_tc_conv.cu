// tcgen05 (5th-gen tensor core) form of the K-tap moving-average conv (A3, AR.py:61-62) and of its
// data gradient, each fused with its epilogue.  See nma_tc.cuh for the operand layouts.
//   k_tc_pack_w      : conv kernel [K][51][50] -> per-tap UMMA B tiles [K][14][64 hi rows | 64 lo rows][4] (fwd / flipped+transposed)
//   k_conv_fwd_tc    : conv + theta-bias + ELU (TMEM -> smem tile) + flow_epilogue (hidden 1x1, head, affine update)
//   k_conv_dgrad_tc  : full correlation of dA with the flipped kernel -> df (feature channels) and dx (channel 0)
//   nma_tc_conv_raw  : test hook, the bare contraction on caller-provided data
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

int launch_pack_weights_tc(nma_handle_s* h, const float* params, bool need_bwd, cudaStream_t st) {
    for (int i = 0; i < h->cfg.F; ++i) {
        k_tc_pack_w<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, 0, h->ws[i].wtc_f);
        nma_count_launch(1);
        if (need_bwd) {
            k_tc_pack_w<<<148, 256, 0, st>>>(params + h->po[i].convw, h->cfg.K, 1, h->ws[i].wtc_d);
            nma_count_launch(1);
        }
    }
    NMA_CHECK_CUDA(cudaGetLastError());
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
    const FlowDims& d = h->fd[i];
    ConvFwdTcArgs a;
    const int nacc = h->tc_nacc;
    a.src.a_hi = h->ws[i].tin_hi; a.src.a_lo = h->ws[i].tin_lo; a.src.Qalloc = h->ws[i].tin_Q;
    a.src.wt = h->ws[i].wtc_f; a.src.K = h->cfg.K;
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
    const FlowDims& d = h->fd[i];
    ConvDgradTcArgs a;
    const int nacc = h->tc_nacc;
    a.src.a_hi = h->ws[i].dat_hi; a.src.a_lo = h->ws[i].dat_lo; a.src.Qalloc = h->ws[i].dat_Q;
    a.src.wt = h->ws[i].wtc_d; a.src.K = h->cfg.K;
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

template <int NACC>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_conv_raw(TcConvSrc src, int npos, long long Q, float* __restrict__ out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bars[TC_NBARS];
    __shared__ uint32_t tmem_slot;
    constexpr int NCOLS = NACC * TC_M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TcConvSmem s = tc_conv_carve(smem, npos, bars, &tmem_slot);
    const uint32_t tmem = tc_conv_setup(s, NACC * 2 * TC_N);
    const long long q0 = (long long)blockIdx.x * NCOLS;
    tc_conv_mainloop<NACC>(s, src, q0, npos, tmem);
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

extern "C" int nma_tc_conv_raw(const float* d_in, const float* d_w, int32_t mode, int32_t nacc, float* d_out, int64_t Q,
                               int32_t K, void* stream) {
    if (!d_in || !d_w || !d_out || Q < 1 || K < 1 || (nacc != 1 && nacc != 2) || (mode != 0 && mode != 1)) {
        nma_set_error("nma_tc_conv_raw: bad argument");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int npos = tc_conv_npos(nacc, K);
    const int ncols = nacc * TC_M;
    const long long Qalloc = (Q + ncols - 1) / ncols * ncols + npos;
    float *hi = nullptr, *lo = nullptr, *wt = nullptr;
    const size_t abytes = (size_t)TC_CCH * Qalloc * 16;
    NMA_CHECK_CUDA(cudaMalloc(&hi, abytes));
    NMA_CHECK_CUDA(cudaMalloc(&lo, abytes));
    NMA_CHECK_CUDA(cudaMalloc(&wt, (size_t)K * TC_WSTAGE * 4));
    NMA_CHECK_CUDA(cudaMemsetAsync(hi, 0, abytes, st));
    NMA_CHECK_CUDA(cudaMemsetAsync(lo, 0, abytes, st));
    k_tc_split_in<<<296, 256, 0, st>>>(d_in, Q, Qalloc, hi, lo);
    k_tc_pack_w<<<148, 256, 0, st>>>(d_w, K, mode, wt);
    TcConvSrc src;
    src.a_hi = hi; src.a_lo = lo; src.Qalloc = Qalloc; src.wt = wt; src.K = K;
    const int smem = (int)(tc_conv_smem_floats(nacc, K) * 4);
    const unsigned grid = (unsigned)((Q + ncols - 1) / ncols);
    cudaError_t e;
    if (nacc == 2) {
        e = cudaFuncSetAttribute(k_tc_conv_raw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) k_tc_conv_raw<2><<<grid, TC_THREADS, smem, st>>>(src, npos, Q, d_out);
    } else {
        e = cudaFuncSetAttribute(k_tc_conv_raw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) k_tc_conv_raw<1><<<grid, TC_THREADS, smem, st>>>(src, npos, Q, d_out);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(hi); cudaFree(lo); cudaFree(wt);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        nma_set_error("nma_tc_conv_raw: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return -2;
    }
    return 0;
}

