// Causal K-tap moving-average convolution core (A3 of SURVEY §8a; AR.py:61-62),
// shared by the forward conv and the data-gradient conv.
//
// FP32 SIMT formulation (the correctness anchor; the conv is a 2550-deep x 50-wide contraction
// whose inputs are un-normalised features, so it is kept in full fp32):
//   * one lane owns TM=10 consecutive output positions of one row and 10 output channels
//     (one warp = one group of 10 output channels): 100 accumulators per thread;
//   * the K taps slide a 10-float register window over the row: per tap one new input value
//     (LDS.32) and 10 weights (3 x LDS.128, warp-uniform broadcast) feed 100 FMAs
//     (issued as 50 packed FFMA2 on sm_100a);
//   * the loop is input-channel outermost, so only ONE input channel of the tile and its
//     [groups][taps][12] weight slab are resident at a time: both stream through a 4-stage
//     shared-memory ring filled by bulk async copies (cp.async.bulk -> UBLKCP) that complete
//     on mbarriers.  Nothing of size [rows, L, 50] is ever re-read from HBM.
#pragma once
#include "nma_common.cuh"

#define CONV_TM 10          // positions per lane == tap unroll
#define CONV_WPAD 12        // 10 weights padded to 3 float4
#define CONV_STAGES 4

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct ConvSrc {
    // input channel c of row r lives at  chan0 (c==0 && chan0!=nullptr) or rest + (c - has0) * chan_stride
    const float* chan0;       // [rows][row_stride0]            (may be nullptr)
    long long row_stride0;
    const float* rest;        // [rows][nrest][chan_stride]
    long long row_stride;
    int chan_stride;
    int copy_floats;          // floats copied per (row, channel) — multiple of 4
    int dst_off;              // float offset inside the smem row where the copy lands (multiple of 4)
};

// Shared-memory ring geometry
struct ConvRing {
    int cin;          // input channels
    int ngroups;      // weight groups per channel (5 fwd, 6 dgrad)
    int KP;           // taps padded to a multiple of 10
    int rc;           // rows handled by this CTA
    int row_pitch;    // floats per smem input row
    int stage_floats; // floats per stage = ngroups*KP*12 + rc*row_pitch
};

// acc[j][q]: position m0+j, output channels 10*group + 2q, 2q+1.
// `my_x_off`: offset inside the smem row of (output position m0, tap 0).
// `wide`: warp-uniform; false = the group has a single live output channel (x-channel of the data gradient).
__device__ __forceinline__ void conv_main_loop(float2 (&acc)[CONV_TM][5], float* ring, uint64_t* full_bar,
                                               const ConvRing& rg, const ConvSrc& src, const float* wpk,
                                               int row0, int my_row, int my_x_off, int group, bool lane_active,
                                               bool wide, int c_begin = 0, int c_end = -1) {
    // [c_begin, c_end): the input channels this CTA reduces over (a channel split of the launch, see conv_split_reduce)
    if (c_end < 0) c_end = rg.cin;
    const int nch = c_end - c_begin;
    const int tid = threadIdx.x;
    const int wslab = rg.ngroups * rg.KP * CONV_WPAD;  // floats
    const uint32_t bytes_per_stage = (uint32_t)(wslab + rg.rc * src.copy_floats) * 4u;

    auto issue = [&](int cl) {
        const int c = c_begin + cl;
        int st = cl % CONV_STAGES;
        float* sbase = ring + (size_t)st * rg.stage_floats;
        mbar_expect_tx(&full_bar[st], bytes_per_stage);
        bulk_g2s(sbase, wpk + (size_t)c * wslab, (uint32_t)wslab * 4u, &full_bar[st]);
        for (int r = 0; r < rg.rc; ++r) {
            const float* g;
            if (c == 0 && src.chan0)
                g = src.chan0 + (size_t)(row0 + r) * src.row_stride0;
            else
                g = src.rest + (size_t)(row0 + r) * src.row_stride + (size_t)(c - (src.chan0 ? 1 : 0)) * src.chan_stride;
            bulk_g2s(sbase + wslab + r * rg.row_pitch + src.dst_off, g, (uint32_t)src.copy_floats * 4u, &full_bar[st]);
        }
    };

    if (tid == 0) {
        for (int c = 0; c < CONV_STAGES && c < nch; ++c) issue(c);
    }

#pragma unroll
    for (int j = 0; j < CONV_TM; ++j)
#pragma unroll
        for (int q = 0; q < 5; ++q) acc[j][q] = make_float2(0.f, 0.f);

    const int nkb = rg.KP / CONV_TM;
    for (int c = 0; c < nch; ++c) {
        const int st = c % CONV_STAGES;
        mbar_wait(&full_bar[st], (uint32_t)((c / CONV_STAGES) & 1));
        const float* sbase = ring + (size_t)st * rg.stage_floats;
        const float4* w4 = reinterpret_cast<const float4*>(sbase) + (size_t)group * rg.KP * 3;
        const float* xs = sbase + wslab + my_row * rg.row_pitch + my_x_off;
        if (lane_active) {
            float2 xx[CONV_TM];     // the sliding window, each value duplicated for the packed FMA
#pragma unroll
            for (int j = 0; j < CONV_TM; ++j) { const float v = xs[j]; xx[j] = make_float2(v, v); }
            if (wide) {
                for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                    for (int u = 0; u < CONV_TM; ++u) {
                        const int k = kb * CONV_TM + u;
                        const float4 wa = w4[k * 3 + 0];
                        const float4 wb = w4[k * 3 + 1];
                        const float4 wc = w4[k * 3 + 2];
                        const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w);
                        const float2 w2 = make_float2(wb.x, wb.y), w3 = make_float2(wb.z, wb.w);
                        const float2 w5 = make_float2(wc.x, wc.y);
#pragma unroll
                        for (int j = 0; j < CONV_TM; ++j) {
                            const float2 xv = xx[(j + u) % CONV_TM];
                            acc[j][0] = __ffma2_rn(xv, w0, acc[j][0]);
                            acc[j][1] = __ffma2_rn(xv, w1, acc[j][1]);
                            acc[j][2] = __ffma2_rn(xv, w2, acc[j][2]);
                            acc[j][3] = __ffma2_rn(xv, w3, acc[j][3]);
                            acc[j][4] = __ffma2_rn(xv, w5, acc[j][4]);
                        }
                        const float v = xs[k + CONV_TM];
                        xx[u] = make_float2(v, v);
                    }
                }
            } else {
                for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                    for (int u = 0; u < CONV_TM; ++u) {
                        const int k = kb * CONV_TM + u;
                        const float w = reinterpret_cast<const float*>(w4 + k * 3)[0];
#pragma unroll
                        for (int j = 0; j < CONV_TM; ++j) acc[j][0].x = fmaf(xx[(j + u) % CONV_TM].x, w, acc[j][0].x);
                        const float v = xs[k + CONV_TM];
                        xx[u] = make_float2(v, v);
                    }
                }
            }
        }
        __syncthreads();   // everyone is done reading this stage
        if (tid == 0 && c + CONV_STAGES < nch) issue(c + CONV_STAGES);
    }
}

// Channel split of a conv launch whose item count alone cannot fill the machine (the scripts' own shapes: p = 1..50 rows).
// gridDim.y CTAs share one block of 32 items, each reducing over its own range of input channels.  Every CTA stores its
// 100 accumulators per thread to `part` (coalesced float4), takes a ticket, and the CTA that draws the last one sums the
// partials in split order - so the result does not depend on which CTA finishes last - and carries on into the epilogue.
// Returns false for the CTAs that are done.  `ticket` is left at zero for the next launch.
__device__ __forceinline__ bool conv_split_reduce(float2 (&acc)[CONV_TM][5], float* part, unsigned* ticket) {
    const int nsplit = gridDim.y;
    if (nsplit == 1) return true;
    __shared__ unsigned s_ticket;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t cta_f4 = (size_t)nthr * 25;
    float4* mine = reinterpret_cast<float4*>(part) + ((size_t)blockIdx.x * nsplit + blockIdx.y) * cta_f4;
#pragma unroll
    for (int t = 0; t < 25; ++t) {
        const float2 lo = acc[(2 * t) / 5][(2 * t) % 5], hi = acc[(2 * t + 1) / 5][(2 * t + 1) % 5];
        __stcg(mine + (size_t)t * nthr + tid, make_float4(lo.x, lo.y, hi.x, hi.y));
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(ticket + blockIdx.x, 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(nsplit - 1)) return false;
    __threadfence();
    const float4* all = reinterpret_cast<const float4*>(part) + (size_t)blockIdx.x * nsplit * cta_f4;
    float4 sum[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) sum[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < nsplit; ++s) {
        const float4* src = all + (size_t)s * cta_f4 + tid;
#pragma unroll
        for (int t = 0; t < 25; ++t) {
            const float4 v = __ldcg(src + (size_t)t * nthr);
            sum[t].x += v.x; sum[t].y += v.y; sum[t].z += v.z; sum[t].w += v.w;
        }
    }
#pragma unroll
    for (int t = 0; t < 25; ++t) {
        acc[(2 * t) / 5][(2 * t) % 5] = make_float2(sum[t].x, sum[t].y);
        acc[(2 * t + 1) / 5][(2 * t + 1) % 5] = make_float2(sum[t].z, sum[t].w);
    }
    if (tid == 0) ticket[blockIdx.x] = 0u;
    return true;
}

// host side: number of channel splits for `item_ctas` CTAs of work over `cin` channels (>= one ring of channels per split)
#define CONV_SPLIT_MAX 32
#define CONV_SPLIT_PART_FLOATS (6 * 32 * 100)      // per CTA: up to 6 warps x 32 lanes x 100 accumulators
static inline int conv_split_count(int item_ctas, int cin, int sm_count) {
    int s = (2 * sm_count) / (item_ctas > 0 ? item_ctas : 1);
    if (s > CONV_SPLIT_MAX) s = CONV_SPLIT_MAX;
    if (s > cin / CONV_STAGES) s = cin / CONV_STAGES;
    return s < 1 ? 1 : s;
}
