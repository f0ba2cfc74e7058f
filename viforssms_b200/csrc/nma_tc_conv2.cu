// Persistent, warp-specialised form of the tcgen05 K-tap conv (nma_tc.cuh) for the forward pass and the data
// gradient.
//
// ncu on the one-tile-per-CTA kernels (profiles/r01_tc_path.md): a CTA spends ~40 us in its MMA mainloop and ~20 us
// outside it - launch, barrier / TMEM set-up, the 140 KB operand-tile load, the TMEM -> global epilogue - with ONE CTA
// per SM (226 KB of shared memory), so the tensor pipe idles a third of the time.  Here one CTA per SM walks the
// tiles; three roles run concurrently:
//   warp 8      TMA producer: operand tile of the next tile as soon as the last MMA of the current one has retired
//               (a_free), taps through the shared-memory ring without regard to tile boundaries
//   warp 9      MMA issuer: accumulator set (it & 1) of TMEM (2 x 256 columns), so the MMAs of tile it+1 run while
//   warps 0-7   the epilogue warps drain tile it out of the other set, registers -> global (coalesced per channel)
// All mbarrier phases are derived from running counters (tile iteration `it`, global tap index `g`).
#include <stdlib.h>
#include "nma_tc.cuh"
#include "nma_flow_epi.cuh"

#define P_THREADS 320
#define P_EPI_WARPS 8

struct PBars {
    uint64_t full[TC_STAGES], empty[TC_STAGES];
    uint64_t a_full, a_free;
    uint64_t acc_full[2], acc_free[2];
    uint64_t hid;
};

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// producer and MMA-issuer roles, shared by both kernels.  NST = ring stages in use (<= TC_STAGES).
template <int NST, bool BF>
__device__ __forceinline__ void p_producer(PBars& b, float* a_hi, float* a_lo, float* wring, const TcConvSrc& src,
                                           int npos, long long ntiles) {
    constexpr int CCH = TcP<BF>::CCH;
    constexpr int WSTAGE = TcP<BF>::WSTAGE;
    const uint32_t slab_bytes = (uint32_t)npos * 16u;
    uint32_t g = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long q0 = tile * (2 * TC_M);
        if (it > 0) mbar_wait_backoff(&b.a_free, (uint32_t)((it - 1) & 1));
        if (elect_one()) {
            mbar_expect_tx(&b.a_full, 2u * CCH * slab_bytes);
            for (int c = 0; c < CCH; ++c) {
                bulk_g2s(a_hi + (size_t)c * npos * 4, src.a_hi + ((size_t)c * src.Qalloc + q0) * 4, slab_bytes, &b.a_full);
                bulk_g2s(a_lo + (size_t)c * npos * 4, src.a_lo + ((size_t)c * src.Qalloc + q0) * 4, slab_bytes, &b.a_full);
            }
        }
        __syncwarp();
        for (int k = 0; k < src.K; ++k, ++g) {
            const uint32_t st = g % NST;
            if (g >= NST) mbar_wait_backoff(&b.empty[st], ((g / NST) - 1u) & 1u);
            if (elect_one()) {
                if (src.diag && it > 0) {          // NMA_DIAG=4 timing experiment (results invalid): taps streamed for the first tile only
                    mbar_expect_tx(&b.full[st], 0u);
                } else {
                    mbar_expect_tx(&b.full[st], WSTAGE * 4u);
                    bulk_g2s(wring + (size_t)st * WSTAGE, src.wt + (size_t)k * WSTAGE, WSTAGE * 4u, &b.full[st]);
                }
            }
            __syncwarp();
        }
    }
}

template <int NST, bool BF>
__device__ __forceinline__ void p_mma(PBars& b, const float* a_hi, const float* a_lo, const float* wring, int K, int npos,
                                      uint32_t tmem_base, long long ntiles) {
    constexpr int CCH = TcP<BF>::CCH;
    constexpr int WSTAGE = TcP<BF>::WSTAGE;
    constexpr uint32_t idesc = umma_idesc<BF>(TC_M, TC_N, 0, 0);
    constexpr uint32_t idesc_wide = umma_idesc<BF>(TC_M, 2 * TC_N, 0, 0);
    const uint32_t a_lbo = (uint32_t)npos * 16u;
    const uint32_t ah_lo0 = desc_lo(smem_u32(a_hi), a_lbo), al_lo0 = desc_lo(smem_u32(a_lo), a_lbo);
    const uint32_t hi32 = desc_hi(128u);
    const uint32_t w_lo0 = desc_lo(smem_u32(wring), TC_WROWS * 16u);
    const uint32_t ks_step_a = 2u * (uint32_t)npos;
    constexpr uint32_t ks_step_b = 2u * TC_WROWS;
    uint32_t g = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t set = (uint32_t)(it & 1);
        mbar_wait_backoff(&b.a_full, (uint32_t)(it & 1));
        if (it >= 2) mbar_wait_backoff(&b.acc_free[set], (uint32_t)(((it >> 1) - 1) & 1));
        for (int k = 0; k < K; ++k, ++g) {
            const uint32_t st = g % NST;
            mbar_wait_backoff(&b.full[st], (g / NST) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t wb = w_lo0 + st * (WSTAGE * 4u / 16u);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const uint32_t row = (uint32_t)(a * TC_M + k);
                    const uint32_t d = tmem_base + set * (4u * TC_N) + (uint32_t)(a * 2 * TC_N);
#pragma unroll
                    for (int ks = 0; ks < CCH / 2; ++ks) {
                        const uint64_t ah = desc_pack(ah_lo0 + row + (uint32_t)ks * ks_step_a, hi32);
                        const uint64_t al = desc_pack(al_lo0 + row + (uint32_t)ks * ks_step_a, hi32);
                        const uint64_t bw = desc_pack(wb + (uint32_t)ks * ks_step_b, hi32);
                        umma<BF>(d, ah, bw, idesc_wide, (k | ks) ? 1u : 0u);
                        umma<BF>(d + TC_N, al, bw, idesc, 1u);
                    }
                }
                tc_commit(&b.empty[st]);
                if (k == K - 1) {
                    tc_commit(&b.acc_full[set]);
                    tc_commit(&b.a_free);
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------
// bf16 split, taps in PAIRS (NMA_TAP_PAIRS=1).  51 (50) reduction channels fill 6.4 (6.25) of the 8-channel chunks, and a
// K = 16 instruction eats two chunks: four instructions per tap, the last one for 3 (2) channels.  The 7th chunk of tap
// k can instead share an instruction with the 7th chunk of tap k+1: on the input side the two K-chunks of that
// instruction are the same channel chunk one position apart - a leading byte offset of 16 in the descriptor - and the
// tap kernels are packed to match ([tap a: chunks 0-5][tap b: chunks 0-5][a: chunk 6][b: chunk 6], 14 x 128 rows x 16 B
// = one ring stage per pair).  Seven instructions per pair instead of eight, and the all-zero 8th chunk of the operand
// tile is no longer loaded.  An odd last tap is paired with a zero kernel.
// ---------------------------------------------------------------------------
#define PW_CH 14
#define PW_STAGE (PW_CH * TC_WROWS * 4)          // floats of one pair stage: 28672 B

template <int NST>
__device__ __forceinline__ void pp_producer(PBars& b, float* a_hi, float* a_lo, float* wring, const TcConvSrc& src,
                                            int npos, long long ntiles) {
    const uint32_t slab_bytes = (uint32_t)npos * 16u;
    const int npairs = (src.K + 1) / 2;
    uint32_t g = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long q0 = tile * (2 * TC_M);
        if (it > 0) mbar_wait_backoff(&b.a_free, (uint32_t)((it - 1) & 1));
        if (elect_one()) {
            mbar_expect_tx(&b.a_full, 2u * 7u * slab_bytes);
            for (int c = 0; c < 7; ++c) {
                bulk_g2s(a_hi + (size_t)c * npos * 4, src.a_hi + ((size_t)c * src.Qalloc + q0) * 4, slab_bytes, &b.a_full);
                bulk_g2s(a_lo + (size_t)c * npos * 4, src.a_lo + ((size_t)c * src.Qalloc + q0) * 4, slab_bytes, &b.a_full);
            }
        }
        __syncwarp();
        for (int pr = 0; pr < npairs; ++pr, ++g) {
            const uint32_t st = g % NST;
            if (g >= NST) mbar_wait_backoff(&b.empty[st], ((g / NST) - 1u) & 1u);
            if (elect_one()) {
                mbar_expect_tx(&b.full[st], PW_STAGE * 4u);
                bulk_g2s(wring + (size_t)st * PW_STAGE, src.wt + (size_t)pr * PW_STAGE, PW_STAGE * 4u, &b.full[st]);
            }
            __syncwarp();
        }
    }
}

// M = 128 positions form (forward): per pair and accumulator 7 x (N=128 with a_hi, N=64 with a_lo)
template <int NST>
__device__ __forceinline__ void pp_mma(PBars& b, const float* a_hi, const float* a_lo, const float* wring, int K, int npos,
                                       uint32_t tmem_base, long long ntiles) {
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N, 0, 0);
    constexpr uint32_t idesc_wide = umma_idesc_bf16(TC_M, 2 * TC_N, 0, 0);
    const uint32_t a_lbo = (uint32_t)npos * 16u;
    const uint32_t ah_lo0 = desc_lo(smem_u32(a_hi), a_lbo), al_lo0 = desc_lo(smem_u32(a_lo), a_lbo);
    // the shared instruction: chunk 6 at tap a, then one position further on (tap b)
    const uint32_t ah_lo6 = desc_lo(smem_u32(a_hi) + 6u * a_lbo, 16u), al_lo6 = desc_lo(smem_u32(a_lo) + 6u * a_lbo, 16u);
    const uint32_t hi32 = desc_hi(128u);
    const uint32_t w_lo0 = desc_lo(smem_u32(wring), TC_WROWS * 16u);
    const uint32_t ks_step_a = 2u * (uint32_t)npos;
    constexpr uint32_t ks_step_b = 2u * TC_WROWS;
    const int npairs = (K + 1) / 2;
    uint32_t g = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t set = (uint32_t)(it & 1);
        mbar_wait_backoff(&b.a_full, (uint32_t)(it & 1));
        if (it >= 2) mbar_wait_backoff(&b.acc_free[set], (uint32_t)(((it >> 1) - 1) & 1));
        for (int pr = 0; pr < npairs; ++pr, ++g) {
            const uint32_t st = g % NST;
            mbar_wait_backoff(&b.full[st], (g / NST) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t wb = w_lo0 + st * (PW_STAGE * 4u / 16u);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const uint32_t row = (uint32_t)(a * TC_M + 2 * pr);
                    const uint32_t d = tmem_base + set * (4u * TC_N) + (uint32_t)(a * 2 * TC_N);
#pragma unroll
                    for (int ks = 0; ks < 7; ++ks) {
                        uint64_t ah, al;
                        if (ks < 6) {
                            const uint32_t off = row + (uint32_t)(ks >= 3 ? 1 : 0) + (uint32_t)(ks % 3) * ks_step_a;
                            ah = desc_pack(ah_lo0 + off, hi32);
                            al = desc_pack(al_lo0 + off, hi32);
                        } else {
                            ah = desc_pack(ah_lo6 + row, hi32);
                            al = desc_pack(al_lo6 + row, hi32);
                        }
                        const uint64_t bw = desc_pack(wb + (uint32_t)ks * ks_step_b, hi32);
                        umma_bf16(d, ah, bw, idesc_wide, (pr | ks) ? 1u : 0u);
                        umma_bf16(d + TC_N, al, bw, idesc, 1u);
                    }
                }
                tc_commit(&b.empty[st]);
                if (pr == npairs - 1) {
                    tc_commit(&b.acc_full[set]);
                    tc_commit(&b.a_free);
                }
            }
            __syncwarp();
        }
    }
}

// position-wide form (data gradient): per pair 7 x (B = dA_hi, B = dA_lo), N = 256
template <int NST>
__device__ __forceinline__ void qq_mma(PBars& b, const float* a_hi, const float* a_lo, const float* wring, int K, int npos,
                                       uint32_t tmem_base, long long ntiles) {
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, 2 * TC_M, 0, 0);
    const uint32_t in_lbo = (uint32_t)npos * 16u;
    const uint32_t bh_lo0 = desc_lo(smem_u32(a_hi), in_lbo), bl_lo0 = desc_lo(smem_u32(a_lo), in_lbo);
    const uint32_t bh_lo6 = desc_lo(smem_u32(a_hi) + 6u * in_lbo, 16u), bl_lo6 = desc_lo(smem_u32(a_lo) + 6u * in_lbo, 16u);
    const uint32_t hi32 = desc_hi(128u);
    const uint32_t w_lo0 = desc_lo(smem_u32(wring), TC_WROWS * 16u);
    const uint32_t ks_step_in = 2u * (uint32_t)npos;
    constexpr uint32_t ks_step_w = 2u * TC_WROWS;
    const int npairs = (K + 1) / 2;
    uint32_t g = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t set = (uint32_t)(it & 1);
        mbar_wait_backoff(&b.a_full, (uint32_t)(it & 1));
        if (it >= 2) mbar_wait_backoff(&b.acc_free[set], (uint32_t)(((it >> 1) - 1) & 1));
        for (int pr = 0; pr < npairs; ++pr, ++g) {
            const uint32_t st = g % NST;
            mbar_wait_backoff(&b.full[st], (g / NST) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t wb = w_lo0 + st * (PW_STAGE * 4u / 16u);
                const uint32_t d = tmem_base + set * (2u * TC_M);
                const uint32_t row = (uint32_t)(2 * pr);
#pragma unroll
                for (int ks = 0; ks < 7; ++ks) {
                    uint64_t bh, bl;
                    if (ks < 6) {
                        const uint32_t off = row + (uint32_t)(ks >= 3 ? 1 : 0) + (uint32_t)(ks % 3) * ks_step_in;
                        bh = desc_pack(bh_lo0 + off, hi32);
                        bl = desc_pack(bl_lo0 + off, hi32);
                    } else {
                        bh = desc_pack(bh_lo6 + row, hi32);
                        bl = desc_pack(bl_lo6 + row, hi32);
                    }
                    const uint64_t aw = desc_pack(wb + (uint32_t)ks * ks_step_w, hi32);
                    umma_bf16(d, aw, bh, idesc, (pr | ks) ? 1u : 0u);
                    umma_bf16(d, aw, bl, idesc, 1u);
                }
                tc_commit(&b.empty[st]);
                if (pr == npairs - 1) {
                    tc_commit(&b.acc_full[set]);
                    tc_commit(&b.a_free);
                }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ uint32_t p_setup(PBars& b, uint32_t* tmem_slot, int a_free_count) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&b.full[i], 1); mbar_init(&b.empty[i], 1); }
        mbar_init(&b.a_full, 1);
        mbar_init(&b.a_free, a_free_count);
        for (int i = 0; i < 2; ++i) { mbar_init(&b.acc_full[i], 1); mbar_init(&b.acc_free[i], P_EPI_WARPS); }
        mbar_init(&b.hid, 1);
        fence_barrier_init();
    }
    if ((threadIdx.x >> 5) == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *tmem_slot;
}

// main + correction accumulator for 32 columns
__device__ __forceinline__ void p_load_sum32(uint32_t taddr, float (&v)[32]) {
    float c[32];
    tmem_ld32(taddr, v);
    tmem_ld32(taddr + TC_N, c);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += c[i];
}

// ---------------------------------------------------------------------------
// data gradient: dinp[q][c] = sum_k' sum_f dA_flat[q + k'][f] * W[K-1-k'][c][f];  c >= 1 -> df[r][c-1][j], c == 0 -> dx[r][j] +=
// ---------------------------------------------------------------------------
struct ConvDgradP {
    TcConvSrc src;
    float* df;           // [p][50][LP]
    float* dx;           // [p][XP]
    int Lin, LP, XP, p, npos, need_dx;
};

template <bool BF>
__global__ void __launch_bounds__(P_THREADS, 1) k_conv_dgrad_tcp(ConvDgradP a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ PBars bars;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* a_hi = smem;
    float* a_lo = a_hi + (size_t)TcP<BF>::CCH * a.npos * 4;
    float* wring = a_lo + (size_t)TcP<BF>::CCH * a.npos * 4;
    const long long qtot = (long long)a.p * a.Lin;
    const long long ntiles = (qtot + 2 * TC_M - 1) / (2 * TC_M);
    const uint32_t tmem = p_setup(bars, &tmem_slot, 1);

    if (warp == 8) {
        p_producer<TC_STAGES, BF>(bars, a_hi, a_lo, wring, a.src, a.npos, ntiles);
    } else if (warp == 9) {
        p_mma<TC_STAGES, BF>(bars, a_hi, a_lo, wring, a.src.K, a.npos, tmem, ntiles);
    } else {
        const int acc = warp >> 2, quarter = warp & 3;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t set = (uint32_t)(it & 1);
            mbar_wait_backoff(&bars.acc_full[set], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const long long q = tile * (2 * TC_M) + acc * TC_M + quarter * 32 + lane;
            const bool valid = q < qtot;
            const int r = valid ? (int)(q / a.Lin) : 0;
            const int j = valid ? (int)(q - (long long)r * a.Lin) : 0;
            const uint32_t ta = tmem + ((uint32_t)(quarter * 32) << 16) + set * (4u * TC_N) + (uint32_t)(acc * 2 * TC_N);
            float* dfp = a.df + (size_t)r * NMA_C * a.LP + j;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[32];
                p_load_sum32(ta + (uint32_t)(half * 32), v);
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int n = half * 32 + i;
                        if (n == 0) { if (a.need_dx) a.dx[(size_t)r * a.XP + j] += v[i]; }
                        else if (n < NMA_C1) dfp[(size_t)(n - 1) * a.LP] = v[i];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(&bars.acc_free[set]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------
// data gradient, position-wide form (bf16 split):  D^T[c][q] = sum_k' sum_f Wd[k'][c][f] * dA_flat[q + k'][f]
//
// Measured (profiles/r01_bf16.md): an M = 128 MMA re-fetches its 128-row A operand from shared memory per instruction
// and that fetch, not the math, sets its duration (~116 cycles at N = 128, ~100 at N = 64, against 64 / 32 of math).
// Here the SMALL matrix is A - the tap kernel, hi and lo part of output channel c stacked in rows 2c, 2c+1 (M = 128) -
// and the 256 positions of the tile are N, the tap being a start-address shift of the B descriptor: two N = 256
// instructions per k-step (B = dA_hi, B = dA_lo) give all FOUR products of the split for 256 positions, against four
// instructions (two of them N = 64) for three products before.  The accumulator arrives channel-major - TMEM lane =
// 2c + {hi, lo} row, column = position - which is the layout of df [row][channel][slot]: one shuffle adds the two
// lanes of a channel, then every even lane stores consecutive slots of its channel.
// ---------------------------------------------------------------------------
template <int NST>
__device__ __forceinline__ void q_mma(PBars& b, const float* a_hi, const float* a_lo, const float* wring, int K, int npos,
                                      uint32_t tmem_base, long long ntiles) {
    constexpr int WSTAGE = TcP<true>::WSTAGE;
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, 2 * TC_M, 0, 0);       // M = 128 stacked kernel rows, N = 256 positions
    const uint32_t in_lbo = (uint32_t)npos * 16u;                           // between the two 8-channel chunks of a k-step
    const uint32_t bh_lo0 = desc_lo(smem_u32(a_hi), in_lbo), bl_lo0 = desc_lo(smem_u32(a_lo), in_lbo);
    const uint32_t hi32 = desc_hi(128u);
    const uint32_t w_lo0 = desc_lo(smem_u32(wring), TC_WROWS * 16u);
    const uint32_t ks_step_in = 2u * (uint32_t)npos;
    constexpr uint32_t ks_step_w = 2u * TC_WROWS;
    uint32_t g = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t set = (uint32_t)(it & 1);
        mbar_wait_backoff(&b.a_full, (uint32_t)(it & 1));
        if (it >= 2) mbar_wait_backoff(&b.acc_free[set], (uint32_t)(((it >> 1) - 1) & 1));
        for (int k = 0; k < K; ++k, ++g) {
            const uint32_t st = g % NST;
            mbar_wait_backoff(&b.full[st], (g / NST) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t wb = w_lo0 + st * (WSTAGE * 4u / 16u);
                const uint32_t d = tmem_base + set * (2u * TC_M);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t aw = desc_pack(wb + (uint32_t)ks * ks_step_w, hi32);
                    const uint64_t bh = desc_pack(bh_lo0 + (uint32_t)k + (uint32_t)ks * ks_step_in, hi32);
                    const uint64_t bl = desc_pack(bl_lo0 + (uint32_t)k + (uint32_t)ks * ks_step_in, hi32);
                    umma_bf16(d, aw, bh, idesc, (k | ks) ? 1u : 0u);
                    umma_bf16(d, aw, bl, idesc, 1u);
                }
                tc_commit(&b.empty[st]);
                if (k == K - 1) {
                    tc_commit(&b.acc_full[set]);
                    tc_commit(&b.a_free);
                }
            }
            __syncwarp();
        }
    }
}

template <bool PAIR>
__global__ void __launch_bounds__(P_THREADS, 1) k_conv_dgrad_tcq(ConvDgradP a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ PBars bars;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CCHU = PAIR ? 7 : TcP<true>::CCH;           // chunk slabs of the operand tile actually staged
    float* a_hi = smem;
    float* a_lo = a_hi + (size_t)CCHU * a.npos * 4;
    float* wring = a_lo + (size_t)CCHU * a.npos * 4;
    const long long qtot = (long long)a.p * a.Lin;
    const long long ntiles = (qtot + 2 * TC_M - 1) / (2 * TC_M);
    const uint32_t tmem = p_setup(bars, &tmem_slot, 1);

    if (warp == 8) {
        if (PAIR) pp_producer<TC_STAGES>(bars, a_hi, a_lo, wring, a.src, a.npos, ntiles);
        else p_producer<TC_STAGES, true>(bars, a_hi, a_lo, wring, a.src, a.npos, ntiles);
    } else if (warp == 9) {
        if (PAIR) qq_mma<TC_STAGES>(bars, a_hi, a_lo, wring, a.src.K, a.npos, tmem, ntiles);
        else q_mma<TC_STAGES>(bars, a_hi, a_lo, wring, a.src.K, a.npos, tmem, ntiles);
    } else {
        const int quarter = warp & 3, colhalf = warp >> 2;
        const int c = (quarter * 32 + lane) >> 1;                  // output channel of this TMEM lane (rows 2c, 2c+1)
        const bool writer = ((lane & 1) == 0) && c < NMA_C1;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t set = (uint32_t)(it & 1);
            mbar_wait_backoff(&bars.acc_full[set], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t ta = tmem + ((uint32_t)(quarter * 32) << 16) + set * (2u * TC_M) + (uint32_t)(colhalf * TC_M);
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
                float v[32];
                tmem_ld32(ta + (uint32_t)(blk * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
                const long long q = tile * (2 * TC_M) + colhalf * TC_M + blk * 32;
                if (writer && q < qtot) {
                    int r = (int)(q / a.Lin);
                    int j = (int)(q - (long long)r * a.Lin);
                    const int nvalid = (qtot - q < 32) ? (int)(qtot - q) : 32;
                    if (c >= 1) {
                        float* dst = a.df + ((size_t)r * NMA_C + (c - 1)) * a.LP;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < nvalid) dst[j] = v[i];
                            if (++j == a.Lin) { j = 0; dst += (size_t)NMA_C * a.LP; }
                        }
                    } else if (a.need_dx) {
                        float* dst = a.dx + (size_t)r * a.XP;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < nvalid) dst[j] += v[i];
                            if (++j == a.Lin) { j = 0; dst += a.XP; }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(&bars.acc_free[set]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// operand tile rows of the pair kernels: 256 positions + taps 0..K (an odd last tap is paired with a zero kernel at tap K)
static int pair_npos(int K) { return (2 * TC_M + K + 7) & ~7; }

int launch_conv_dgrad_tcp(nma_handle_s* h, int i, int p, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    ConvDgradP a;
    a.src.a_hi = h->ws[i].dat_hi; a.src.a_lo = h->ws[i].dat_lo; a.src.Qalloc = h->ws[i].dat_Q;
    a.src.wt = h->ws[i].wtc_d; a.src.K = h->cfg.K;
    a.src.diag = (nma_diag_bits() & 4) ? 1 : 0;
    a.df = h->ws[i].df; a.dx = h->ws[i].dx;
    a.Lin = d.Lin; a.LP = d.LP; a.XP = (d.L + 3) & ~3; a.p = p; a.npos = tc_conv_npos(2, h->cfg.K);
    a.need_dx = i > 0 ? 1 : 0;
    const long long ntiles = ((long long)p * d.Lin + 2 * TC_M - 1) / (2 * TC_M);
    const int grid = (int)(ntiles < h->sm_count ? ntiles : h->sm_count);
    if (h->use_bf16 && h->dgrad_wide && h->tap_pairs) {
        a.npos = pair_npos(h->cfg.K);
        const int smem = 2 * 7 * a.npos * 16 + TC_STAGES * PW_STAGE * 4;
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_dgrad_tcq<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_dgrad_tcq<true><<<grid, P_THREADS, smem, st>>>(a);
        nma_count_launch(1);
        NMA_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (h->use_bf16 && h->dgrad_wide) {
        const int smem = (int)(tc_conv_smem_floats(2, h->cfg.K, TcP<true>::CCH) * 4);
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_dgrad_tcq<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_dgrad_tcq<false><<<grid, P_THREADS, smem, st>>>(a);
        nma_count_launch(1);
        NMA_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (h->use_bf16) {
        const int smem = (int)(tc_conv_smem_floats(2, h->cfg.K, TcP<true>::CCH) * 4);
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_dgrad_tcp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_dgrad_tcp<true><<<grid, P_THREADS, smem, st>>>(a);
        nma_count_launch(1);
        NMA_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    const int smem = (int)(tc_conv_smem_floats(2, h->cfg.K) * 4);
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_dgrad_tcp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    k_conv_dgrad_tcp<false><<<grid, P_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// forward: conv + theta-bias + ELU, hidden 1x1 layer ON THE TENSOR CORES, head, softplus, affine flow update
// (AR.py:58-89) for the configuration of the AR scripts (one hidden layer, no batch-norm, flow_dims = 1).
//
// Epilogue of a tile (warps 0-7, thread = one of the 256 positions): the conv accumulators are read out of TMEM,
// e_0 = elu(. + theta-bias) is saved and written - hi/lo split - as the A operand of the hidden layer into the operand
// buffer the mainloop has just released (E_hi | E_lo | packed hidden kernel, exactly the 140 KB of a 320-position tile);
// one elected thread issues the 2 x 7 x 2 MMAs of the hidden layer into the same TMEM columns; e_1 = elu(. + b) comes
// back to registers, where the 2-unit head, softplus and the affine update are thread-local.  Nothing is staged
// through shared memory except the MMA operand.  The next tile's operand load waits for this epilogue (a_free counts the
// 8 epilogue warps as well); the weight ring keeps streaming the next tile's taps meanwhile.
// ---------------------------------------------------------------------------
struct ConvFwdP {
    TcConvSrc src;
    const float* tb;         // [p][3][50]; slot 2 = theta bias + conv bias
    const float* whid;       // packed hidden kernel [14][64 hi | 64 lo rows][4] (forward orientation)
    const float* hidb;       // [50]
    const float* headw;      // [50][2]
    const float* headb;      // [2]
    const float* x_in;       // [p][XP]
    float* x_out;            // [p][XPn]
    float* h0;               // [p][50][NP]
    float* h1;
    float* s;                // [p][NP]
    float* nx_hi;            // channel 0 of the next flow's conv operand (may be null)
    float* nx_lo;
    int nx_Lin;
    int Lin, p, npos, N, NP, XP, XPn, K, save, ring_off, e_off;
};

#define PF_E_F (TC_CCH * 2 * TC_M * 4)          // floats of E_hi (or E_lo): [14][256][4]

template <bool BF, bool PAIR>
__global__ void __launch_bounds__(P_THREADS, 1) k_conv_fwd_tcp(ConvFwdP a) {
    static_assert(BF || !PAIR, "tap pairs exist in the bf16 split only");
    extern __shared__ __align__(128) float smem[];
    __shared__ PBars bars;
    __shared__ uint64_t wh_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float hb_sm[64], hw_sm[128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // operand region: the conv operand tile (hi | lo) during the mainloop, E_hi | E_lo | hidden kernel (3xTF32, 140 KB)
    // during the epilogue; the weight ring follows the larger of the two (a.ring_off floats)
    float* a_hi = smem;
    float* a_lo = a_hi + (size_t)(PAIR ? 7 : TcP<BF>::CCH) * a.npos * 4;
    float* wring = smem + a.ring_off;
    // 3xTF32: the epilogue's E_hi | E_lo | hidden kernel alias the operand buffer (e_off = 0), so the next tile's operand
    // load waits for the epilogue.  bf16 split: the operand tile is 80 KB and the epilogue operands - bf16 as well,
    // [8][256][16 B] hi and lo + the 16 KB kernel - have their own 80 KB behind it (e_off): the next tile's load starts
    // the moment the last MMA of the mainloop has retired, and the hidden kernel is fetched once.
    constexpr int E_F = BF ? (TcP<true>::CCH * 2 * TC_M * 4) : PF_E_F;
    float* E_hi = smem + a.e_off;
    float* E_lo = E_hi + E_F;
    float* Wh = E_lo + E_F;
    const long long qtot = (long long)a.p * a.Lin;
    const long long ntiles = (qtot + 2 * TC_M - 1) / (2 * TC_M);
    if (tid < 64) hb_sm[tid] = (tid < NMA_C) ? a.hidb[tid] : 0.f;
    if (tid < 128) hw_sm[tid] = (tid < 2 * NMA_C) ? a.headw[tid] : 0.f;
    if (tid == 0) mbar_init(&wh_bar, 1);
    if (BF) {
        // chunk 7 (channels 56..63) of E is never written: zero it once
        uint4* z = reinterpret_cast<uint4*>(E_hi);
        for (int t = tid; t < 2 * 2 * TC_M; t += blockDim.x)
            z[(t < 2 * TC_M ? 7 * 2 * TC_M + t : E_F / 4 + 7 * 2 * TC_M + (t - 2 * TC_M))] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }
    const uint32_t tmem = p_setup(bars, &tmem_slot, BF ? 1 : 1 + P_EPI_WARPS);

    if (warp == 8) {
        if (PAIR) pp_producer<2>(bars, a_hi, a_lo, wring, a.src, a.npos, ntiles);          // 2 x 28 KB: what fits next to E
        else p_producer<TC_STAGES, BF>(bars, a_hi, a_lo, wring, a.src, a.npos, ntiles);
    } else if (warp == 9) {
        if (PAIR) pp_mma<2>(bars, a_hi, a_lo, wring, a.src.K, a.npos, tmem, ntiles);
        else p_mma<TC_STAGES, BF>(bars, a_hi, a_lo, wring, a.src.K, a.npos, tmem, ntiles);
    } else {
        const int acc = warp >> 2, quarter = warp & 3;
        const int col = acc * TC_M + quarter * 32 + lane;
        const float hb0 = a.headb[0], hb1 = a.headb[1];
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t set = (uint32_t)(it & 1);
            mbar_wait_backoff(&bars.acc_full[set], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            // the mainloop is done with the operand buffer: fetch the packed hidden kernel into its tail
            // (bf16 split: own region, fetched once)
            if (warp == 0 && (!BF || it == 0)) {
                if (elect_one()) {
                    mbar_expect_tx(&wh_bar, (uint32_t)TcP<BF>::WSTAGE * 4u);
                    bulk_g2s(Wh, a.whid, (uint32_t)TcP<BF>::WSTAGE * 4u, &wh_bar);
                }
                __syncwarp();
            }
            const long long q = tile * (2 * TC_M) + col;
            const int r = (int)(q / a.Lin);
            const int m = (int)(q - (long long)r * a.Lin);
            const bool ok = r < a.p && m < a.N;
            const int rr = ok ? r : 0;
            const uint32_t ta = tmem + ((uint32_t)(quarter * 32) << 16) + set * (4u * TC_N) + (uint32_t)(acc * 2 * TC_N);
            const float* tbr = a.tb + ((size_t)rr * 3 + 2) * NMA_C;
            // ---- e_0 = elu(A + theta-bias + conv bias)  (AR.py:70-72) -> saved, and A operand of the hidden layer ----
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[32];
                p_load_sum32(ta + (uint32_t)(half * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int f = half * 32 + i;
                    v[i] = (ok && f < NMA_C) ? elu_f(v[i] + __ldg(tbr + (f < NMA_C ? f : 0))) : 0.f;
                }
                if (a.save && ok) {
                    char* pb = reinterpret_cast<char*>(a.h0 + ((size_t)r * NMA_C + half * 32) * a.NP + m);
                    const size_t step = (size_t)a.NP * sizeof(float);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (half * 32 + i < NMA_C) *reinterpret_cast<float*>(pb) = v[i];
                        pb += step;
                    }
                }
                if (BF) {
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        if (half == 0 || c8 < 3) {
                            float w8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) w8[e] = v[8 * c8 + e];
                            uint4 h4, l4;
                            bf_split8(w8, h4, l4);
                            const size_t o = (size_t)(half * 4 + c8) * (2 * TC_M) + col;
                            reinterpret_cast<uint4*>(E_hi)[o] = h4;
                            reinterpret_cast<uint4*>(E_lo)[o] = l4;
                        }
                    }
                } else
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    if (half == 0 || cc < 6) {
                        const size_t o = ((size_t)(half * 8 + cc) * (2 * TC_M) + col) * 4;
                        const float4 h4 = make_float4(tf32_hi(v[4 * cc]), tf32_hi(v[4 * cc + 1]), tf32_hi(v[4 * cc + 2]),
                                                      tf32_hi(v[4 * cc + 3]));
                        *reinterpret_cast<float4*>(E_hi + o) = h4;
                        *reinterpret_cast<float4*>(E_lo + o) = make_float4(v[4 * cc] - h4.x, v[4 * cc + 1] - h4.y,
                                                                           v[4 * cc + 2] - h4.z, v[4 * cc + 3] - h4.w);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- hidden 1x1 layer (AR.py:74-76) on the tensor cores, both 128-position accumulators ----
            if (warp == 0) {
                if (!BF || it == 0) mbar_wait_backoff(&wh_bar, BF ? 0u : (uint32_t)(it & 1));
                tc_fence_after();
                if (elect_one()) {
                    constexpr uint32_t idesc = umma_idesc<BF>(TC_M, TC_N, 0, 0);
                    constexpr uint32_t idesc_wide = umma_idesc<BF>(TC_M, 2 * TC_N, 0, 0);
                    constexpr uint32_t e_lbo = 2u * TC_M * 16u;                 // between channel chunks: 256 positions
                    const uint32_t eh0 = desc_lo(smem_u32(E_hi), e_lbo), el0 = desc_lo(smem_u32(E_lo), e_lbo);
                    const uint32_t w0 = desc_lo(smem_u32(Wh), TC_WROWS * 16u);
                    const uint32_t hi32 = desc_hi(128u);
#pragma unroll
                    for (int a2 = 0; a2 < 2; ++a2) {
                        const uint32_t d = tmem + set * (4u * TC_N) + (uint32_t)(a2 * 2 * TC_N);
                        const uint32_t row = (uint32_t)(a2 * TC_M);             // 16-byte units
#pragma unroll
                        for (int ks = 0; ks < TcP<BF>::CCH / 2; ++ks) {
                            const uint64_t eh = desc_pack(eh0 + row + (uint32_t)ks * (2u * e_lbo / 16u), hi32);
                            const uint64_t el = desc_pack(el0 + row + (uint32_t)ks * (2u * e_lbo / 16u), hi32);
                            const uint64_t bw = desc_pack(w0 + (uint32_t)ks * (2u * TC_WROWS), hi32);
                            umma<BF>(d, eh, bw, idesc_wide, ks ? 1u : 0u);
                            umma<BF>(d + TC_N, el, bw, idesc, 1u);
                        }
                    }
                    tc_commit(&bars.hid);
                }
                __syncwarp();
            }
            mbar_wait_backoff(&bars.hid, (uint32_t)(it & 1));
            tc_fence_after();
            // ---- e_1 = elu(. + b), head (AR.py:77-78), softplus, affine update (AR.py:83-85): thread-local ----
            float mu = hb0, sr = hb1;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[32];
                p_load_sum32(ta + (uint32_t)(half * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int f = half * 32 + i;
                    if (f < NMA_C) {
                        v[i] = elu_f(v[i] + hb_sm[f]);
                        mu = fmaf(v[i], hw_sm[2 * f], mu);
                        sr = fmaf(v[i], hw_sm[2 * f + 1], sr);
                    }
                }
                if (a.save && ok) {
                    char* pb = reinterpret_cast<char*>(a.h1 + ((size_t)r * NMA_C + half * 32) * a.NP + m);
                    const size_t step = (size_t)a.NP * sizeof(float);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (half * 32 + i < NMA_C) *reinterpret_cast<float*>(pb) = v[i];
                        pb += step;
                    }
                }
            }
            if (ok) {
                const float xin = a.x_in[(size_t)r * a.XP + m + a.K];
                const float sigma = softplus_f(sr) + 1e-10f;            // AR.py:83
                const float xo = fmaf(xin, sigma, mu);                  // AR.py:85
                a.s[(size_t)r * a.NP + m] = sr;
                a.x_out[(size_t)r * a.XPn + m] = xo;
                if (a.nx_hi && m < a.nx_Lin) {                          // channel 0 of the next flow's conv operand
                    if (BF) {
                        const size_t qn = ((size_t)r * a.nx_Lin + m) * 8;    // bf16 elements: unit q of chunk 0, slot 0
                        uint32_t hi, lo;
                        bf_split(xo, hi, lo);
                        reinterpret_cast<uint16_t*>(a.nx_hi)[qn] = (uint16_t)hi;
                        reinterpret_cast<uint16_t*>(a.nx_lo)[qn] = (uint16_t)lo;
                    } else {
                        const size_t qn = ((size_t)r * a.nx_Lin + m) * 4;
                        const float hi = tf32_hi(xo);
                        a.nx_hi[qn] = hi;
                        a.nx_lo[qn] = xo - hi;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive1(&bars.acc_free[set]);
                if (!BF) mbar_arrive1(&bars.a_free);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// forward-orientation pack of ONE 1x1 kernel [50 in][50 out]: element (cch, n, e) = W[4*cch + e][n]
__global__ void k_tc_pack_w1x1(const float* __restrict__ W, float* __restrict__ out) {
    for (int t = threadIdx.x; t < TC_WSTAGE; t += blockDim.x) {
        const int e = t & 3, row = (t >> 2) & 127, cch = t >> 9;
        const int n = row & 63, c = 4 * cch + e;
        const float v = (c < NMA_C && n < NMA_C) ? W[c * NMA_C + n] : 0.f;
        const float hi = tf32_hi(v);
        out[t] = (row < 64) ? hi : v - hi;
    }
}

// the same in the bf16 split: [8][64 hi | 64 lo rows][8 x bf16], element (cch, n, e) = W[8*cch + e][n]
__global__ void k_tc_pack_w1x1_bf(const float* __restrict__ W, uint16_t* __restrict__ out) {
    for (int t = threadIdx.x; t < 8 * TC_N * 8; t += blockDim.x) {
        const int e = t & 7, n = (t >> 3) & (TC_N - 1), cch = t >> 9;
        const int c = 8 * cch + e;
        const float v = (c < NMA_C && n < NMA_C) ? W[c * NMA_C + n] : 0.f;
        uint32_t hi, lo;
        bf_split(v, hi, lo);
        out[((size_t)cch * TC_WROWS + n) * 8 + e] = (uint16_t)hi;
        out[((size_t)cch * TC_WROWS + n + TC_N) * 8 + e] = (uint16_t)lo;
    }
}

static int fwd_p_npos(int K) {
    const int n = tc_conv_npos(2, K);
    return n < 320 ? 320 : n;       // the epilogue needs E_hi | E_lo | hidden kernel = 140 KB inside the operand buffer
}

int conv_fwd_tcp_supported(const nma_handle_s* h) {
    if (!(h->use_tc && h->use_tc_persist && h->tc_nacc == 2 && h->cfg.H == 1 && !h->cfg.bn && h->cfg.D == 1)) return 0;
    const size_t smem = (size_t)2 * TC_CCH * fwd_p_npos(h->cfg.K) * 16 + (size_t)TC_STAGES * TC_WSTAGE * 4;
    return smem + 1024 <= 227 * 1024;      // (the bf16 operand tile is smaller: same bound)
}

// hidden 1x1 kernel of flow i in the forward orientation -> slot 9 of the flow's pack buffer (part of the step's weight
// packing, launch_pack_weights_tc: off the forward pass's serial chain)
int launch_pack_w1x1_fwd(nma_handle_s* h, int i, const float* params, cudaStream_t st) {
    float* whid = h->ws[i].wtc_feat + (size_t)9 * TC_WSTAGE;
    if (h->use_bf16) k_tc_pack_w1x1_bf<<<1, 256, 0, st>>>(params + h->po[i].hidw[0], (uint16_t*)whid);
    else k_tc_pack_w1x1<<<1, 256, 0, st>>>(params + h->po[i].hidw[0], whid);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_conv_fwd_tcp(nma_handle_s* h, int i, const float* params, int p, bool save, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    float* whid = h->ws[i].wtc_feat + (size_t)9 * TC_WSTAGE;     // slot 9 of the flow's pack buffer (launch_pack_w1x1_fwd)
    ConvFwdP a;
    a.src.a_hi = h->ws[i].tin_hi; a.src.a_lo = h->ws[i].tin_lo; a.src.Qalloc = h->ws[i].tin_Q;
    a.src.wt = h->ws[i].wtc_f; a.src.K = h->cfg.K;
    a.src.diag = (nma_diag_bits() & 4) ? 1 : 0;
    a.tb = h->ws[i].tb; a.whid = whid;
    a.hidb = params + h->po[i].hidb[0]; a.headw = params + h->po[i].headw; a.headb = params + h->po[i].headb;
    a.x_in = h->ws[i].x; a.x_out = h->ws[i + 1].x; a.h0 = h->ws[i].h[0]; a.h1 = h->ws[i].h[1]; a.s = h->ws[i].s;
    const bool next_tc = (i + 1 < h->cfg.F);
    a.nx_hi = next_tc ? h->ws[i + 1].tin_hi : nullptr;
    a.nx_lo = next_tc ? h->ws[i + 1].tin_lo : nullptr;
    a.nx_Lin = next_tc ? h->fd[i + 1].Lin : 0;
    a.Lin = d.Lin; a.p = p; a.npos = fwd_p_npos(h->cfg.K); a.N = d.N; a.NP = d.NP;
    a.XP = (d.L + 3) & ~3; a.XPn = (h->fd[i + 1].L + 3) & ~3; a.K = h->cfg.K; a.save = save ? 1 : 0;
    const long long ntiles = ((long long)p * d.Lin + 2 * TC_M - 1) / (2 * TC_M);
    const int grid = (int)(ntiles < h->sm_count ? ntiles : h->sm_count);
    // operand region = max(conv operand tile, epilogue view E_hi | E_lo | hidden kernel); the ring follows it
    const int epi_f = 2 * PF_E_F + TC_WSTAGE;
    if (h->use_bf16 && h->tap_pairs) {
        // [operand tile hi | lo, 7 chunks][E_hi | E_lo | hidden kernel, all bf16][ring: 2 pair stages]
        a.npos = pair_npos(h->cfg.K);
        a.e_off = 2 * 7 * a.npos * 4;
        a.ring_off = a.e_off + 2 * (TcP<true>::CCH * 2 * TC_M * 4) + TcP<true>::WSTAGE;
        const int smem = (a.ring_off + 2 * PW_STAGE) * 4;
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tcp<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_fwd_tcp<true, true><<<grid, P_THREADS, smem, st>>>(a);
        nma_count_launch(1);
        NMA_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (h->use_bf16) {
        // [operand tile hi | lo][E_hi | E_lo | hidden kernel, all bf16][ring]
        a.e_off = 2 * TcP<true>::CCH * a.npos * 4;
        a.ring_off = a.e_off + 2 * (TcP<true>::CCH * 2 * TC_M * 4) + TcP<true>::WSTAGE;
        const int smem = (a.ring_off + TC_STAGES * TcP<true>::WSTAGE) * 4;
        static int configured = 0;
        if (configured < smem) {
            NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tcp<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        k_conv_fwd_tcp<true, false><<<grid, P_THREADS, smem, st>>>(a);
        nma_count_launch(1);
        NMA_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    a.ring_off = 2 * TC_CCH * a.npos * 4;
    a.e_off = 0;
    (void)epi_f;
    const int smem = 2 * TC_CCH * a.npos * 16 + TC_STAGES * TC_WSTAGE * 4;
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tcp<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    k_conv_fwd_tcp<false, false><<<grid, P_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}
