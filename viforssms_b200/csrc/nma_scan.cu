// A12/A13: data-side scans for 10^8-step synthetic series (AR_dat_gen.py:11-31), float64 like the reference.
//   nma_scan_ar1  : x[i] = a x[i-1] + b + c z[i-1] as a prefix scan over affine maps (compose-then-apply)
//   nma_time_till : hold-fill, observation indicator and count-down to the next observation
// Both are HBM-bound streaming kernels.
#include "nma_common.cuh"

#define SC_THREADS 256
#define SC_ITEMS 16
#define SC_CHUNK (SC_THREADS * SC_ITEMS)

struct Aff { double A, D; };   // x -> A x + D
__device__ __forceinline__ Aff compose(const Aff& first, const Aff& second) {   // apply `first`, then `second`
    Aff r;
    r.A = second.A * first.A;
    r.D = fma(second.A, first.D, second.D);
    return r;
}

// block-wide inclusive scan of per-thread composites (Hillis-Steele over shared memory)
__device__ __forceinline__ Aff block_scan(Aff v, Aff* sh) {
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < SC_THREADS; o <<= 1) {
        Aff prev = sh[t];
        if (t >= o) prev = compose(sh[t - o], sh[t]);
        __syncthreads();
        sh[t] = prev;
        __syncthreads();
    }
    return sh[t];
}

__device__ __forceinline__ Aff thread_composite(const double* __restrict__ z, int64_t n, int64_t base, double a, double b,
                                                double c) {
    Aff acc; acc.A = 1.0; acc.D = 0.0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int64_t i = base + k;
        if (i < n) { Aff e; e.A = a; e.D = fma(c, z[i], b); acc = compose(acc, e); }
    }
    return acc;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_reduce(const double* __restrict__ z, int64_t n, double a, double b,
                                                            double c, Aff* __restrict__ agg) {
    __shared__ Aff sh[SC_THREADS];
    const int64_t base = (int64_t)blockIdx.x * SC_CHUNK + (int64_t)threadIdx.x * SC_ITEMS;
    Aff v = thread_composite(z, n, base, a, b, c);
    v = block_scan(v, sh);
    if (threadIdx.x == SC_THREADS - 1) agg[blockIdx.x] = v;
}

// sequential carry over the (n / 4096) block aggregates: start value of every block
__global__ void k_scan_carry(Aff* __restrict__ agg, int64_t nblocks, double x0) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double x = x0;
    for (int64_t b = 0; b < nblocks; ++b) {
        const Aff g = agg[b];
        agg[b].D = x;                 // reuse the slot: D <- value entering block b
        x = fma(g.A, x, g.D);
    }
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_apply(const double* __restrict__ z, double* __restrict__ x, int64_t n,
                                                           double x0, double a, double b, double c,
                                                           const Aff* __restrict__ agg) {
    __shared__ Aff sh[SC_THREADS];
    const int64_t base = (int64_t)blockIdx.x * SC_CHUNK + (int64_t)threadIdx.x * SC_ITEMS;
    Aff v = thread_composite(z, n, base, a, b, c);
    Aff inc = block_scan(v, sh);
    // exclusive prefix of this thread = inclusive of the previous thread
    __syncthreads();
    sh[threadIdx.x] = inc;
    __syncthreads();
    double xin = agg[blockIdx.x].D;
    if (threadIdx.x > 0) { const Aff pre = sh[threadIdx.x - 1]; xin = fma(pre.A, xin, pre.D); }
    if (blockIdx.x == 0 && threadIdx.x == 0) x[0] = x0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int64_t i = base + k;
        if (i < n) { xin = fma(a, xin, fma(c, z[i], b)); x[i + 1] = xin; }
    }
}

extern "C" int nma_scan_ar1(const double* d_z, double* d_x, int64_t n, double x0, double a, double b, double c,
                            void* d_scratch, int64_t scratch_bytes, void* stream) {
    if (!d_z || !d_x || n < 1 || !d_scratch) { nma_set_error("nma_scan_ar1: bad argument"); return -1; }
    const int64_t nblocks = (n + SC_CHUNK - 1) / SC_CHUNK;
    if (scratch_bytes < nblocks * (int64_t)sizeof(Aff)) {
        nma_set_error("nma_scan_ar1: scratch needs %lld bytes", (long long)(nblocks * sizeof(Aff)));
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Aff* agg = (Aff*)d_scratch;
    k_scan_reduce<<<(unsigned)nblocks, SC_THREADS, 0, st>>>(d_z, n, a, b, c, agg);
    nma_count_launch(1);
    k_scan_carry<<<1, 32, 0, st>>>(agg, nblocks, x0);
    nma_count_launch(1);
    k_scan_apply<<<(unsigned)nblocks, SC_THREADS, 0, st>>>(d_z, d_x, n, x0, a, b, c, agg);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// AR_dat_gen.py:17-31.  obs has n+1 entries; kept items are obs[impute], obs[2*impute], ...
__global__ void k_time_till(const double* __restrict__ obs, int64_t m_out, int impute, double* __restrict__ fill,
                            double* __restrict__ binary, double* __restrict__ till) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m_out; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / impute;
        const int j = (int)(i - q * impute);
        const double item = obs[(q + 1) * impute];
        const double partial = (j == impute - 1) ? item : 0.0;
        const bool seen = partial != 0.0;                       // AR_dat_gen.py:21 (an exact 0.0 counts as missing)
        if (fill) fill[i] = item;
        if (binary) binary[i] = seen ? 1.0 : 0.0;
        double tt = 0.0;
        if (!seen) {
            // distance back to the last observed slot (or to the start): the reference's running `count`
            int64_t k = i - 1;
            while (k >= 0) {
                const int64_t qk = k / impute;
                const bool obs_slot = (k - qk * impute) == impute - 1;
                if (obs_slot && obs[(qk + 1) * impute] != 0.0) break;
                --k;
            }
            tt = (double)(i - k);
        }
        if (till) till[i] = -(tt - (double)impute);             // AR_dat_gen.py:31
    }
}

extern "C" int nma_time_till(const double* d_obs, int64_t n, int32_t impute, double* d_fill, double* d_binary,
                             double* d_till, void* stream) {
    if (!d_obs || n < 1 || impute < 1 || n < impute) { nma_set_error("nma_time_till: bad argument"); return -1; }
    const int64_t m_out = ((n - impute) / impute + 1) * impute;
    int64_t blocks = (m_out + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_time_till<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_obs, m_out, impute, d_fill, d_binary, d_till);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// A14 - rolling variances of the stochastic-volatility features (SV_dense.py:159-170):
//     var_store[i] = np.var(obs[i : i + K])      on the script's float32 series.
// Bit-exact with numpy: np.var is sum -> /K -> (x - mean)^2 -> sum -> /K, every step in float32, and numpy's float32
// sum is the pairwise scheme of loops_utils.h (8 interleaved accumulators over blocks of 8, combined as a balanced
// tree, remainder added sequentially; halves of at most 128 elements above that).  One thread per window; windows
// overlap, so the loads hit L1.  No FMA contraction (explicit _rn intrinsics): a fused (x - m)^2 + acc would round
// differently.
// ---------------------------------------------------------------------------
template <typename F>
__device__ float np_pairwise_sum(F elem, int64_t lo, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, elem(lo + i));
        return res;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = elem(lo + j);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], elem(lo + i + j));
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, elem(lo + i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum(elem, lo, n2), np_pairwise_sum(elem, lo + n2, n - n2));
}

__global__ void k_rolling_var(const float* __restrict__ x, int64_t nout, int K, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nout; i += (int64_t)gridDim.x * blockDim.x) {
        const float s = np_pairwise_sum([&](int64_t j) { return __ldg(x + j); }, i, K);
        const float mean = (float)((double)s / (double)K);
        const float s2 = np_pairwise_sum([&](int64_t j) { const float d = __fsub_rn(__ldg(x + j), mean); return __fmul_rn(d, d); }, i, K);
        out[i] = (float)((double)s2 / (double)K);
    }
}

extern "C" int nma_rolling_var(const float* d_x, int64_t n, int32_t K, float* d_var, void* stream) {
    if (!d_x || !d_var || K < 1 || n <= K) { nma_set_error("nma_rolling_var: bad argument"); return -1; }
    const int64_t nout = n - K;
    int64_t blocks = (nout + 127) / 128;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_rolling_var<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(d_x, nout, K, d_var);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}
