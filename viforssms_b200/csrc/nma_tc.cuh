// tcgen05 / TMEM / bulk-copy building blocks for the tensor-core form of the K-tap moving-average conv
// (A3 of SURVEY §8a; AR.py:61-62) on sm_100a.
//
// The conv  out[q][n] = sum_k sum_c in[q + k][c] * W[k][c][n]  is a GEMM whose A rows for tap k are the
// SAME input rows shifted by k positions.  The input tile is kept in shared memory ONCE, in the
// no-swizzle K-major "interleaved" UMMA layout  [channel/4][position][4 channels]  (every position
// is one 16-byte unit, core matrices = 8 consecutive positions), so tap k is nothing but a +16k-byte
// start address in the A descriptor: no im2col, no re-load.  Weights stream tap by tap through a
// shared-memory ring filled by the TMA engine (cp.async.bulk) and released by tcgen05.commit.
//
// Precision: the contraction is 2550 deep over un-normalised activations and must hold 1e-4 on
// gradients, so every product is the 3xTF32 split  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
// (hi = top 19 bits, lo = a - hi, both exact in fp32; fp32 accumulation in TMEM).
//
// Second operand format ("BF", TcP<true>): the 2-term bfloat16 split  a ~= a_hi + a_lo  (both round-to-nearest bf16,
// 16 significand bits together) with the same three products on kind::f16, which runs at twice the kind::tf32 rate
// and packs 8 channels into a 16-byte unit: 4 k-steps of 16 channels per tap instead of 7 k-steps of 8.  The layout
// is byte-for-byte analogous - [channel/8][position][8 x bf16], one 16-byte unit per position and chunk - so every
// descriptor below is shared; only the chunk count, the instruction descriptor and the MMA kind differ.
// Measured on the fp64 oracle (tools/split_precision.py): worst per-variable gradient error of the whole step
// 4e-6 .. 8e-6 with this split on all three conv GEMMs (bar 1e-4; 3xTF32: 4e-7).
#pragma once
#include <cuda_bf16.h>
#include <stdlib.h>
#include "nma_conv_core.cuh"

#define TC_CCH 14                 // reduction chunks of 4 channels: 56 >= 51 (fwd) / 50 (dgrad)
#define TC_N 64                   // UMMA N: 50 / 51 outputs padded to 64 (M=128 needs N % 16 == 0)
#define TC_M 128
#define TC_WHALF (TC_CCH * TC_N * 4)          // floats of one (hi or lo) weight part of one tap
#define TC_WSTAGE (2 * TC_WHALF)              // one tap = [14][128 rows: 64 hi | 64 lo][4]: 7168 floats = 28672 B
#define TC_WROWS (2 * TC_N)                   // B-operand rows per channel chunk (hi rows then lo rows)
#define TC_STAGES 3

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// operand format traits: BF = false 3xTF32 (4 channels per 16-byte unit), BF = true 2-term bf16 split (8 channels)
template <bool BF>
struct TcP {
    static constexpr int CCH = BF ? 8 : TC_CCH;        // 16-byte channel chunks per position (64 / 56 channel slots)
    static constexpr int CPU = BF ? 8 : 4;             // channels per 16-byte unit
    static constexpr int WHALF = CCH * TC_N * 4;       // floats (= 4-byte words) of the hi or lo part of one tap
    static constexpr int WSTAGE = 2 * WHALF;           // one tap: [CCH][64 hi rows | 64 lo rows][16 B]
};

// 2-term bf16 split, both parts round-to-nearest
__device__ __forceinline__ void bf_split(float v, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi = (uint32_t)__bfloat16_as_ushort(h);
    lo = (uint32_t)__bfloat16_as_ushort(l);
}
// 8 consecutive channels of one position -> the 16-byte hi unit and the 16-byte lo unit
__device__ __forceinline__ void bf_split8(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bf_split(v[e], h[e], l[e]);
    hi = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    lo = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
}
__device__ __forceinline__ float bf_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }

// NMA_DIAG timing experiments (tools/r02_diag.sh) skip loads or MMAs and produce WRONG results on purpose: they are
// honoured only when NMA_DIAG_I_KNOW_RESULTS_ARE_INVALID=1 is set as well, so a stray variable cannot corrupt a run.
static inline int nma_diag_bits() {
    const char* ok = getenv("NMA_DIAG_I_KNOW_RESULTS_ARE_INVALID");
    const char* ed = getenv("NMA_DIAG");
    return (ok && ok[0] == '1' && ed) ? atoi(ed) : 0;
}

// ---- mbarrier extras ----
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done) {
        if (clock64() - t0 > 8000000000LL) __trap();   // ~4 s: a lost arrival must fail loudly, not hang the device
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// warp-collective wait with a single polling lane
// (all lanes poll: measured 30 % faster on B200 than one polling lane + __syncwarp, which breaks the warp-uniform
// issue path of the following MMAs)
__device__ __forceinline__ void mbar_wait_lane0(uint64_t* bar, uint32_t parity) { mbar_wait_backoff(bar, parity); }

// ---- TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors ----
// shared-memory matrix descriptor, SWIZZLE_NONE: start address, leading-dimension byte offset (between the two
// 16-byte K chunks of one MMA for K-major operands / between 8-element K groups for MN-major operands), stride
// byte offset (between 8-row groups for K-major / between 4-element MN chunks for MN-major); all >> 4.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version: Blackwell
    return d;                   // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE
}
// instruction descriptor for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// instruction descriptor for kind::f16 with bf16 operands, fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool BF>
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return BF ? umma_idesc_bf16(M, N, a_mn_major, b_mn_major) : umma_idesc_tf32(M, N, a_mn_major, b_mn_major);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <bool BF>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (BF) umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
    else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}

// ---------------------------------------------------------------------------
// Shared-memory carve-up of the conv mainloop
// ---------------------------------------------------------------------------
struct TcConvSmem {
    float* a_hi;       // [14][npos][4]
    float* a_lo;       // [14][npos][4]
    float* wring;      // [TC_STAGES][TC_WSTAGE]
    uint64_t* full;    // [TC_STAGES]
    uint64_t* empty;   // [TC_STAGES]
    uint64_t* a_bar;   // input tile landed
    uint64_t* acc_bar; // all MMAs retired
    uint32_t* tmem_slot;
};
__host__ __device__ inline int tc_conv_npos(int nacc, int K) { return (nacc * TC_M + K - 1 + 7) & ~7; }
__host__ __device__ inline size_t tc_conv_smem_floats(int nacc, int K, int cch = TC_CCH) {
    return (size_t)2 * cch * tc_conv_npos(nacc, K) * 4 + (size_t)TC_STAGES * (2 * cch * TC_N * 4);
}
__device__ __forceinline__ TcConvSmem tc_conv_carve(float* smem, int npos, uint64_t* bars, uint32_t* tmem_slot,
                                                    int cch = TC_CCH) {
    TcConvSmem s;
    s.a_hi = smem;
    s.a_lo = s.a_hi + (size_t)cch * npos * 4;
    s.wring = s.a_lo + (size_t)cch * npos * 4;
    s.full = bars;
    s.empty = bars + TC_STAGES;
    s.a_bar = bars + 2 * TC_STAGES;
    s.acc_bar = bars + 2 * TC_STAGES + 1;
    s.tmem_slot = tmem_slot;
    return s;
}
#define TC_NBARS (2 * TC_STAGES + 2)

struct TcConvSrc {
    const float* a_hi;     // [CCH][Qalloc][16 B]   hi parts of the flattened input (position q = row*Lin + slot)
    const float* a_lo;
    long long Qalloc;
    const float* wt;       // [K][CCH][128][16 B] packed taps (rows 0-63 hi parts, rows 64-127 lo parts)
    int K;
    int diag;              // timing experiments only (NMA_DIAG), 0 in every product path
};

// one lane of a converged warp; the surrounding control flow stays warp-uniform so descriptors live in
// uniform registers and the MMA issues as a plain predicated UTCHMMA
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred));
    return pred;
}
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
// low word of a SWIZZLE_NONE descriptor: start address and leading byte offset, both >> 4
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3fffu) | ((lbo_bytes >> 4) << 16);
}
// high word: stride byte offset >> 4, descriptor version 1 (bit 46 of the descriptor)
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14); }

// Runs the whole contraction for NACC x 128 consecutive flattened positions starting at q0.
// Must be called by ALL threads of the CTA (>= 2 warps); returns after every MMA has retired and the
// accumulators are readable with tcgen05.ld: accumulator a holds the main sum in columns [128a, 128a+64) and
// the 3xTF32 correction terms in [128a+64, 128a+128) of tmem_base (lane = position); their sum is the result.
// Barriers and TMEM must have been set up by tc_conv_setup().
template <int NACC, bool BF = false>
__device__ __forceinline__ void tc_conv_mainloop(const TcConvSmem& s, const TcConvSrc& src, long long q0, int npos,
                                                 uint32_t tmem_base) {
    constexpr int CCH = TcP<BF>::CCH;
    constexpr int WSTAGE = TcP<BF>::WSTAGE;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
        const uint32_t slab_bytes = (uint32_t)npos * 16u;
        if (elect_one()) {
            mbar_expect_tx(s.a_bar, 2u * CCH * slab_bytes);
            for (int c = 0; c < CCH; ++c) {
                bulk_g2s(s.a_hi + (size_t)c * npos * 4, src.a_hi + ((size_t)c * src.Qalloc + q0) * 4, slab_bytes, s.a_bar);
                bulk_g2s(s.a_lo + (size_t)c * npos * 4, src.a_lo + ((size_t)c * src.Qalloc + q0) * 4, slab_bytes, s.a_bar);
            }
        }
        __syncwarp();
        for (int k = 0; k < src.K; ++k) {
            const int st = k % TC_STAGES;
            if (k >= TC_STAGES) mbar_wait_lane0(&s.empty[st], (uint32_t)(((k / TC_STAGES) - 1) & 1));
            if (elect_one()) {
                mbar_expect_tx(&s.full[st], WSTAGE * 4u);
                bulk_g2s(s.wring + (size_t)st * WSTAGE, src.wt + (size_t)k * WSTAGE, WSTAGE * 4u, &s.full[st]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc<BF>(TC_M, TC_N, 0, 0);
        const uint32_t a_lbo = (uint32_t)npos * 16u;          // between channel chunks
        const uint32_t ah_lo0 = desc_lo(smem_u32(s.a_hi), a_lbo), al_lo0 = desc_lo(smem_u32(s.a_lo), a_lbo);
        const uint32_t a_hi32 = desc_hi(128u), b_hi32 = desc_hi(128u);
        const uint32_t w_lo0 = desc_lo(smem_u32(s.wring), TC_WROWS * 16u);
        const uint32_t ks_step_a = 2u * (uint32_t)npos;       // (2 channel chunks) >> 4
        constexpr uint32_t ks_step_b = 2u * TC_WROWS;         // 2 chunks of 128 rows x 16 B, >> 4
        constexpr uint32_t idesc_wide = umma_idesc<BF>(TC_M, 2 * TC_N, 0, 0);
        mbar_wait_lane0(s.a_bar, 0);
        for (int k = 0; k < src.K; ++k) {
            const int st = k % TC_STAGES;
            mbar_wait_lane0(&s.full[st], (uint32_t)((k / TC_STAGES) & 1));
            tc_fence_after();
            if (elect_one()) {
                const uint32_t wb = w_lo0 + (uint32_t)st * (WSTAGE * 4u / 16u);
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const uint32_t row = (uint32_t)(a * TC_M + k);     // 16-byte units
                    const uint32_t d = tmem_base + (uint32_t)(a * 2 * TC_N);
#pragma unroll
                    for (int ks = 0; ks < CCH / 2; ++ks) {
                        const uint64_t ah = desc_pack(ah_lo0 + row + (uint32_t)ks * ks_step_a, a_hi32);
                        const uint64_t al = desc_pack(al_lo0 + row + (uint32_t)ks * ks_step_a, a_hi32);
                        const uint64_t bw = desc_pack(wb + (uint32_t)ks * ks_step_b, b_hi32);
                        // 3xTF32: the B tile holds 64 hi rows then 64 lo rows, so ONE N=128 MMA with a_hi gives
                        // a_hi*b_hi (columns 0-63, the main sum) and a_hi*b_lo (columns 64-127, the correction sum)
                        // for a single read of A; a_lo*b_hi (N=64) then lands on the correction columns.
                        // The tensor core accumulates with truncation: keeping the 2^-11 smaller correction terms in
                        // their own accumulator makes the main chain 3x shorter; the epilogue adds the two in fp32.
                        umma<BF>(d, ah, bw, idesc_wide, (k | ks) ? 1u : 0u);
                        umma<BF>(d + TC_N, al, bw, idesc, 1u);
                    }
                }
                tc_commit(&s.empty[st]);      // frees the weight stage once these MMAs have read it
                if (k == src.K - 1) tc_commit(s.acc_bar);
            }
            __syncwarp();
        }
    }
    // one lane polls for the last commit; everybody else parks on the hardware barrier (32-lane polling of a
    // shared-memory mbarrier for the whole mainloop would compete with the MMA operand reads)
    if (warp == 1) mbar_wait_backoff(s.acc_bar, 0);
    __syncthreads();
    tc_fence_after();
}

// one-time per-CTA setup: barrier init + TMEM allocation (warp 0 allocates and later frees)
__device__ __forceinline__ uint32_t tc_conv_setup(const TcConvSmem& s, uint32_t tmem_cols) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
        mbar_init(s.a_bar, 1);
        mbar_init(s.acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(s.tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *s.tmem_slot;
}
__device__ __forceinline__ void tc_conv_teardown(uint32_t tmem_base, uint32_t tmem_cols) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem_base, tmem_cols);
}
