// A12/A13: data-side scans for 10^8-step synthetic series (AR_dat_gen.py:11-31), float64 like the reference.
//   nma_scan_ar1    : x[i] = a x[i-1] + b + c z[i-1] as a prefix scan over affine maps
//   nma_scan_affine : x[i] = A[i-1] x[i-1] + D[i-1] (per-element maps: the stochastic-volatility generator, SV_dense.py:211-223)
//   nma_time_till   : hold-fill, observation indicator and count-down to the next observation
// All HBM-bound streaming kernels.
//
// The scan is SINGLE-PASS with decoupled look-back (Merrill & Garland 2016): every tile of 4096 elements is read once
// and written once - 16 B per element for the AR(1) form (8 B noise in, 8 B state out), which is the algorithmic
// minimum - instead of the reduce / carry / apply passes (24 B per element and a serial carry over 24 k tile aggregates)
// this file used to have.  Tiles take their index from an atomic ticket (so a tile's predecessors are always resident
// or finished), scan themselves with warp shuffles (thread composite -> warp inclusive scan -> 8 warp aggregates), publish
// their aggregate, and warp 0 walks back over up to 32 predecessor descriptors at a time until it meets one that already
// carries an inclusive prefix.  An affine map is 16 bytes, so a descriptor is (flag, A, D) with the payload written
// before the flag and fenced (release) / read after it (acquire).  Composition is associative but not commutative:
// every reduction below keeps the older map on the inside.
#include "nma_common.cuh"

#define SC_THREADS 256
#define SC_ITEMS 16
#define SC_CHUNK (SC_THREADS * SC_ITEMS)
#define SC_PADDED (SC_CHUNK + SC_CHUNK / 16)
#define SC_PAD(i) ((i) + ((i) >> 4))            // one pad slot per 16 doubles: a thread's 16 items at stride 17 (no bank conflicts)

struct Aff { double A, D; };   // x -> A x + D
__device__ __forceinline__ Aff compose(const Aff& first, const Aff& second) {   // apply `first`, then `second`
    Aff r;
    r.A = second.A * first.A;
    r.D = fma(second.A, first.D, second.D);
    return r;
}
__device__ __forceinline__ Aff aff_shfl_up(const Aff& v, int o) {
    Aff r; r.A = __shfl_up_sync(0xffffffffu, v.A, o); r.D = __shfl_up_sync(0xffffffffu, v.D, o); return r;
}
__device__ __forceinline__ Aff aff_shfl_down(const Aff& v, int o) {
    Aff r; r.A = __shfl_down_sync(0xffffffffu, v.A, o); r.D = __shfl_down_sync(0xffffffffu, v.D, o); return r;
}

// tile descriptor: flag 0 = not ready, 1 = aggregate of this tile alone, 2 = inclusive prefix of everything up to it
struct TileDesc { int flag; int pad; Aff agg; Aff incl; };
struct ScanScratch { unsigned int ticket; unsigned int pad[3]; };     // followed by TileDesc[ntiles]

struct ScanArgs {
    const double* z;        // AR(1) form: noise; affine form: D
    const double* Aarr;     // affine form: per-element multiplier (null: the constant a)
    double* x;              // [n + 1]
    long long n;
    double x0, a, b, c;
    ScanScratch* scratch;
};

template <bool GENERIC>
__global__ void __launch_bounds__(SC_THREADS, 5) k_scan_lookback(ScanArgs g) {
    extern __shared__ double sc_smem[];                 // [SC_PADDED] D (then the outputs); generic form: + [SC_PADDED] A
    double* sD = sc_smem;
    double* sA = sc_smem + SC_PADDED;
    __shared__ Aff s_warp[SC_THREADS / 32];
    __shared__ Aff s_tile_excl;
    __shared__ unsigned int s_tile;
    TileDesc* desc = reinterpret_cast<TileDesc*>(g.scratch + 1);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(&g.scratch->ticket, 1u);
    __syncthreads();
    const long long tile = s_tile;
    const long long base = tile * SC_CHUNK;
    // ---- coalesced load of the tile (element i of the tile -> padded slot) ----
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int i = k * SC_THREADS + t;
        const long long gi = base + i;
        double d = 0.0, a = 1.0;                       // identity map beyond the end
        if (gi < g.n) {
            if (GENERIC) { d = __ldg(g.z + gi); a = __ldg(g.Aarr + gi); }
            else { d = fma(g.c, __ldg(g.z + gi), g.b); a = g.a; }
        }
        sD[SC_PAD(i)] = d;
        if (GENERIC) sA[SC_PAD(i)] = a;
    }
    __syncthreads();
    // ---- thread composite over its 16 consecutive elements ----
    // (the elements stay in shared memory and are read again in the apply phase: 5 tiles per SM hide the look-back
    // latency, which a 128-register kernel with 2 tiles per SM did not - 28 % of the HBM rate)
    Aff mine; mine.A = 1.0; mine.D = 0.0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int i = t * SC_ITEMS + k;
        const double eD = sD[SC_PAD(i)];
        const double eA = GENERIC ? sA[SC_PAD(i)] : ((base + i < g.n) ? g.a : 1.0);
        mine.D = fma(eA, mine.D, eD);
        mine.A = eA * mine.A;
    }
    // ---- warp inclusive scan (shuffles), warp aggregates, scan of the 8 aggregates ----
    Aff inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Aff prev = aff_shfl_up(inc, o);
        if (lane >= o) inc = compose(prev, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    Aff warp_excl; warp_excl.A = 1.0; warp_excl.D = 0.0;
    Aff tile_agg; tile_agg.A = 1.0; tile_agg.D = 0.0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; ++w) {
        if (w == warp) warp_excl = tile_agg;
        tile_agg = compose(tile_agg, s_warp[w]);
    }
    // exclusive prefix of this thread within the tile
    Aff up = aff_shfl_up(inc, 1);
    if (lane == 0) { up.A = 1.0; up.D = 0.0; }
    const Aff thr_excl = compose(warp_excl, up);
    // ---- publish the aggregate, look back (warp 0) ----
    if (warp == 0) {
        TileDesc* me = desc + tile;
        if (tile == 0) {
            if (lane == 0) {
                me->incl = tile_agg;
                __threadfence();
                *reinterpret_cast<volatile int*>(&me->flag) = 2;
                s_tile_excl.A = 1.0; s_tile_excl.D = 0.0;
            }
        } else {
            if (lane == 0) {
                me->agg = tile_agg;
                __threadfence();
                *reinterpret_cast<volatile int*>(&me->flag) = 1;
            }
            Aff excl; excl.A = 1.0; excl.D = 0.0;       // composite of the predecessors gathered so far (newer side)
            long long look = tile - 1;                  // newest tile of the current window
            while (true) {
                const long long j = look - lane;        // lane l looks at the l-th predecessor of the window
                int flag = 2;
                Aff v; v.A = 1.0; v.D = 0.0;
                if (j >= 0) {
                    const volatile int* fp = reinterpret_cast<const volatile int*>(&desc[j].flag);
                    do { flag = *fp; } while (flag == 0);
                    __threadfence();
                    const volatile double* src = reinterpret_cast<const volatile double*>(flag == 2 ? &desc[j].incl : &desc[j].agg);
                    v.A = src[0]; v.D = src[1];
                }
                // first lane (newest-to-oldest) that carries an inclusive prefix (lanes past tile 0 count as inclusive identity)
                const unsigned incl_mask = __ballot_sync(0xffffffffu, flag == 2);
                const int cut = incl_mask ? (__ffs(incl_mask) - 1) : 31;      // lanes 0..cut take part
                if (lane > cut) { v.A = 1.0; v.D = 0.0; }
                // ordered reduction: lane l <- compose(older lanes ..., lane l); older = higher lane
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const Aff older = aff_shfl_down(v, o);
                    if (lane + o < 32) v = compose(older, v);
                }
                const Aff window = {__shfl_sync(0xffffffffu, v.A, 0), __shfl_sync(0xffffffffu, v.D, 0)};
                excl = compose(window, excl);
                if (incl_mask) break;
                look -= 32;
            }
            if (lane == 0) {
                me->incl = compose(excl, tile_agg);
                __threadfence();
                *reinterpret_cast<volatile int*>(&me->flag) = 2;
                s_tile_excl = excl;
            }
        }
    }
    __syncthreads();
    // ---- apply: value entering this thread's first element, then the recursion itself ----
    const Aff pre = compose(s_tile_excl, thr_excl);
    double xin = fma(pre.A, g.x0, pre.D);
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int i = t * SC_ITEMS + k;
        const double eA = GENERIC ? sA[SC_PAD(i)] : ((base + i < g.n) ? g.a : 1.0);
        xin = fma(eA, xin, sD[SC_PAD(i)]);
        sD[SC_PAD(i)] = xin;
    }
    __syncthreads();
    if (tile == 0 && t == 0) g.x[0] = g.x0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int i = k * SC_THREADS + t;
        const long long gi = base + i;
        if (gi < g.n) g.x[gi + 1] = sD[SC_PAD(i)];
    }
}

static int64_t scan_scratch_bytes(int64_t n) {
    const int64_t ntiles = (n + SC_CHUNK - 1) / SC_CHUNK;
    return (int64_t)sizeof(ScanScratch) + ntiles * (int64_t)sizeof(TileDesc);
}
extern "C" int64_t nma_scan_scratch_bytes(int64_t n) { return n < 1 ? 0 : scan_scratch_bytes(n); }

static int scan_launch(const double* d_z, const double* d_A, double* d_x, int64_t n, double x0, double a, double b, double c,
                       void* d_scratch, int64_t scratch_bytes, cudaStream_t st) {
    const int64_t need = scan_scratch_bytes(n);
    if (scratch_bytes < need) { nma_set_error("nma_scan: scratch needs %lld bytes (nma_scan_scratch_bytes)", (long long)need); return -1; }
    if (((uintptr_t)d_scratch & 15) != 0) { nma_set_error("nma_scan: scratch must be 16-byte aligned"); return -1; }
    const int64_t ntiles = (n + SC_CHUNK - 1) / SC_CHUNK;
    NMA_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, (size_t)need, st));      // ticket and every flag back to 0
    ScanArgs g;
    g.z = d_z; g.Aarr = d_A; g.x = d_x; g.n = n; g.x0 = x0; g.a = a; g.b = b; g.c = c; g.scratch = (ScanScratch*)d_scratch;
    const size_t smem1 = (size_t)SC_PADDED * sizeof(double);
    if (d_A) {
        static bool attr_set = false;
        if (!attr_set) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_scan_lookback<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem1)));
            attr_set = true;
        }
        k_scan_lookback<true><<<(unsigned)ntiles, SC_THREADS, 2 * smem1, st>>>(g);
    } else {
        k_scan_lookback<false><<<(unsigned)ntiles, SC_THREADS, smem1, st>>>(g);
    }
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int nma_scan_ar1(const double* d_z, double* d_x, int64_t n, double x0, double a, double b, double c,
                            void* d_scratch, int64_t scratch_bytes, void* stream) {
    if (!d_z || !d_x || n < 1 || !d_scratch) { nma_set_error("nma_scan_ar1: bad argument"); return -1; }
    return scan_launch(d_z, nullptr, d_x, n, x0, a, b, c, d_scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" int nma_scan_affine(const double* d_A, const double* d_D, double* d_x, int64_t n, double x0, void* d_scratch,
                               int64_t scratch_bytes, void* stream) {
    if (!d_A || !d_D || !d_x || n < 1 || !d_scratch) { nma_set_error("nma_scan_affine: bad argument"); return -1; }
    return scan_launch(d_D, d_A, d_x, n, x0, 1.0, 0.0, 1.0, d_scratch, scratch_bytes, (cudaStream_t)stream);
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
