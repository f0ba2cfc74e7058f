// Epilogue of one NMA flow layer, shared by the FP32 SIMT conv kernel and the tcgen05 conv kernel:
// given e_0 = elu(conv + theta-bias) of a tile of positions in shared memory, runs the hidden 1x1 layers
// (AR.py:74-76) [+ BN-affine, fitz_nag_NVP.py:93], the 2-unit head (AR.py:77-78; stride 2 when D == 2,
// fitz_nag_NVP.py:95-96), softplus, and the locally-affine update x <- x[:, K:] * sigma + mu (AR.py:83-85),
// saving what the backward pass needs.  A tile column is one conv output position of one row; the caller
// provides the column -> (row, position) map.
#pragma once
#include "nma_common.cuh"

#define PW_WPITCH 52            // pointwise-layer weights [50][52] in smem

struct FlowEpiArgs {
    const float* hidw[NMA_MAXH];
    const float* hidb[NMA_MAXH];
    const float* gam[NMA_MAXH];
    const float* bet[NMA_MAXH];
    const float* headw;      // [50][2]
    const float* headb;      // [2]
    const float* x_in;       // [p][XP]   input sample of this flow
    float* x_out;            // [p][XPn]  next flow's input sample (after the pair swap when D==2)
    float* h[NMA_MAXH + 1];  // [p][50][NP]
    float* s;                // [p][NP]
    // tensor-core layout of the NEXT flow's conv input, channel 0 (may be null): [14][nx_Q][4] hi / lo
    float* nx_hi;
    float* nx_lo;
    long long nx_Q;
    int nx_Lin;
    int XP, XPn, N, NP, K, H, bn, D, save, permute_out;
};

// shared-memory floats the epilogue needs for a tile of NCOLS columns
template <int NCOLS>
__host__ __device__ constexpr size_t flow_epi_smem_floats() {
    return (size_t)NMA_C * (NCOLS + 4) + NMA_C * PW_WPITCH + 3 * 64;
}

// in-place per-position dense layer on the [50][NCOLS+4] tile: a thread owns whole columns.
// out[g] = elu(b[g] + sum_f W[f][g] * in[f]) with optional BN-affine applied to the INPUT.
template <int NCOLS>
__device__ __forceinline__ void col_dense_inplace(float* tile, const unsigned char* col_ok, const float* Wsm,
                                                  const float* bsm, const float* in_scale, const float* in_shift) {
    constexpr int PITCH = NCOLS + 4;
    for (int col = threadIdx.x; col < NCOLS; col += blockDim.x) {
        if (!col_ok[col]) continue;
        float* cp = tile + col;
        float acc[52];
#pragma unroll
        for (int g = 0; g < 52; ++g) acc[g] = (g < NMA_C) ? bsm[g] : 0.f;
        for (int f = 0; f < NMA_C; ++f) {
            float xv = cp[f * PITCH];
            if (in_scale) xv = fmaf(xv, in_scale[f], in_shift[f]);
            const float4* w4 = reinterpret_cast<const float4*>(Wsm + f * PW_WPITCH);
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
        for (int g = 0; g < NMA_C; ++g) cp[g * PITCH] = elu_f(acc[g]);
    }
}

// tile: [50][NCOLS+4] holding e_0; scratch: NMA_C*PW_WPITCH + 3*64 floats right behind it.
// col_r / col_m: row and conv position of every column; col_ok: column is a real output.
// D == 2 requires that the column left of an odd position is the even position of the same row.
template <int NCOLS>
__device__ __forceinline__ void flow_epilogue(const FlowEpiArgs& a, float* tile, const int* col_r, const int* col_m,
                                              const unsigned char* col_ok) {
    constexpr int PITCH = NCOLS + 4;
    const int tid = threadIdx.x;
    float* Wsm = tile + NMA_C * PITCH;                // [50][52]
    float* bsm = Wsm + NMA_C * PW_WPITCH;             // [64]
    float* bns = bsm + 64;                            // [64] BN scale
    float* bno = bns + 64;                            // [64] BN shift

    auto store_tile = [&](float* gdst) {   // tile -> global [p][50][NP]; consecutive columns are consecutive m
        for (int t = tid; t < NMA_C * NCOLS; t += blockDim.x) {
            const int f = t / NCOLS, col = t - f * NCOLS;
            if (!col_ok[col]) continue;
            gdst[((size_t)col_r[col] * NMA_C + f) * a.NP + col_m[col]] = tile[(size_t)f * PITCH + col];
        }
    };
    if (a.save) store_tile(a.h[0]);

    // hidden 1x1 layers (AR.py:74-76) [+ BN-affine, fitz_nag_NVP.py:93]
    for (int l = 0; l < a.H; ++l) {
        __syncthreads();
        for (int t = tid; t < NMA_C * PW_WPITCH; t += blockDim.x) {
            const int f = t / PW_WPITCH, g = t - f * PW_WPITCH;
            Wsm[t] = (g < NMA_C) ? a.hidw[l][f * NMA_C + g] : 0.f;
        }
        if (tid < NMA_C) {
            bsm[tid] = a.hidb[l][tid];
            if (a.bn && l > 0) {   // input of layer l is BN_{l-1}(e_l)
                bns[tid] = a.gam[l - 1][tid] * rsqrtf(1.f + 1e-3f);
                bno[tid] = a.bet[l - 1][tid];
            }
        }
        __syncthreads();
        col_dense_inplace<NCOLS>(tile, col_ok, Wsm, bsm, (a.bn && l > 0) ? bns : nullptr, bno);
        __syncthreads();
        if (a.save) store_tile(a.h[l + 1]);
    }
    __syncthreads();
    // head: (mu, s) = conv1x1 -> 2 (AR.py:77-78); stride 2 when D == 2 (fitz_nag_NVP.py:95-96)
    if (tid < NMA_C) {
        if (a.bn && a.H > 0) {
            bns[tid] = a.gam[a.H - 1][tid] * rsqrtf(1.f + 1e-3f);
            bno[tid] = a.bet[a.H - 1][tid];
        } else {
            bns[tid] = 1.f;
            bno[tid] = 0.f;
        }
        Wsm[2 * tid] = a.headw[2 * tid];
        Wsm[2 * tid + 1] = a.headw[2 * tid + 1];
    }
    __syncthreads();
    const float hb0 = a.headb[0], hb1 = a.headb[1];
    for (int col = tid; col < NCOLS; col += blockDim.x) {
        if (!col_ok[col]) continue;
        const int r = col_r[col], m = col_m[col];
        const float xin = a.x_in[(size_t)r * a.XP + m + a.K];
        float xo;
        if (a.D == 1 || (m & 1)) {
            // D==2: odd output slot m uses the head evaluated at the even conv position m-1 (the column to the left)
            const float* cp = tile + ((a.D == 1) ? col : col - 1);
            float mu = hb0, sr = hb1;
            for (int g = 0; g < NMA_C; ++g) {
                const float v = fmaf(cp[g * PITCH], bns[g], bno[g]);
                mu = fmaf(v, Wsm[2 * g], mu);
                sr = fmaf(v, Wsm[2 * g + 1], sr);
            }
            const float sigma = softplus_f(sr) + 1e-10f;   // AR.py:83
            xo = fmaf(xin, sigma, mu);                      // AR.py:85
            a.s[(size_t)r * a.NP + m] = sr;
        } else {
            xo = xin;   // identity slot of the coupling layer (fitz_nag_NVP.py:99-102)
        }
        const int mo = a.permute_out ? (m ^ 1) : m;        // Permute = swap adjacent pairs (fitz_nag_NVP.py:205-211)
        a.x_out[(size_t)r * a.XPn + mo] = xo;
        if (a.nx_hi && mo < a.nx_Lin) {                    // channel 0 of the next flow's tensor-core input
            const size_t qn = ((size_t)r * a.nx_Lin + mo) * 4;
            const float hi = __uint_as_float(__float_as_uint(xo) & 0xffffe000u);
            a.nx_hi[qn] = hi;
            a.nx_lo[qn] = xo - hi;
        }
    }
}
