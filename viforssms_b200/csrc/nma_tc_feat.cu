// tcgen05 (5th-gen tensor core) form of the feature MLP (A2 of SURVEY §8a; AR.py:53-56: 4 x dense(50, elu) applied
// pointwise over the window positions), fused with the window gather (A1, AR.py:267-283) and with the conv's
// operand build.
//
// The FP32 SIMT kernel (k_feat_fwd, nma_fwd.cu) is bound by shared-memory operand fetch at ~19 % of the FP32 FMA
// peak; as a GEMM the four layers are  Y[pos][g] = elu(b[g] + sum_f X[pos][f] W[f][g])  with M = positions,
// N = 50 (padded 64), K = Cf_in (14 -> 16) or 50 (-> 56).  Here a tile of 128 flattened positions
// (q = row*Lin + slot, the same flattening as the conv operand `tin`) goes through all four layers without leaving
// the SM: the A operand lives in shared memory in the no-swizzle K-major layout [channel/4][position][4] (hi and lo
// parts of the 3xTF32 split, see nma_tc.cuh), the packed weights of all layers stay resident in shared memory, the
// accumulator lives in TMEM, and the epilogue (bias, ELU, save for the backward pass, hi/lo split) writes the next
// layer's A operand straight back to shared memory.  The first layer sees raw magnitudes up to T (the time channel,
// AR.py:139-140); the 3xTF32 split keeps ~2^-21 relative error per product whatever the magnitude, so it runs on
// the tensor cores as well.
//
// A CTA hosts TWO independent 256-thread "slots", each walking its own tiles with its own operand buffers, TMEM
// columns, mbarrier and named barrier; they share only the read-only weights.  While one slot runs its SIMT
// epilogue the other's MMAs execute, which is all the overlap the kernel needs (MMA time per layer is ~1/3 of
// the epilogue time).  Per slot: 8 warps; warp w reads TMEM lanes 32*(w%4).. (positions) and columns 32*(w/4)..
// (output channels).
#include <string.h>
#include <stdlib.h>
#include "nma_tc.cuh"

#define FT_SLOTS 2
#define FT_SLOT_THREADS 256
#define FT_THREADS (FT_SLOTS * FT_SLOT_THREADS)
#define FT_M 128
#define FT_CHUNK_F (FT_M * 4)                  // floats of one 4-channel chunk of an operand tile (128 rows x 16 B)
#define FT_WLAYER_F (TC_CCH * FT_CHUNK_F)      // B tile of one dense layer: [14][64 hi rows | 64 lo rows][4]
#define FT_A_F (TC_CCH * FT_CHUNK_F)           // A tile (hi or lo part) of one slot
#define FT_TMEM_COLS 256                       // 2 slots x (64 main + 64 correction columns)

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------
// weight packing: dense kernel [nin][50] -> UMMA B tile [14][128 rows][4], rows 0-63 hi parts, 64-127 lo parts,
// element (cch, n, e) = W[4*cch + e][n].  Slot l of a flow's pack buffer holds layer l (forward); slots 4..7
// hold the transposed kernels of the data gradient, element (cch, n, e) = W[n][4*cch + e].
// ---------------------------------------------------------------------------
struct FeatPackArgs {
    const float* w[NMA_MAX_FLOWS][4];
    float* out[NMA_MAX_FLOWS];
    int Cf_in, with_bwd;
};

__global__ void k_tc_pack_feat(FeatPackArgs a) {
    const int per_flow = a.with_bwd ? 8 : 4;
    const int i = blockIdx.x / per_flow, s = blockIdx.x % per_flow;
    const int l = s & 3, transposed = s >> 2;
    const float* __restrict__ W = a.w[i][l];
    const int nin = (l == 0) ? a.Cf_in : NMA_C;
    float* __restrict__ out = a.out[i] + (size_t)s * FT_WLAYER_F;
    for (int t = threadIdx.x; t < FT_WLAYER_F; t += blockDim.x) {
        const int e = t & 3, row = (t >> 2) & 127, cch = t >> 9;
        const int n = row & 63, c = 4 * cch + e;
        float v = 0.f;
        if (!transposed) {
            if (c < nin && n < NMA_C) v = W[c * NMA_C + n];          // reduction over inputs c, output n
        } else {
            if (c < NMA_C && n < nin) v = W[n * NMA_C + c];          // reduction over outputs c, result = input n
        }
        const float hi = tf32_hi(v);
        out[t] = (row < 64) ? hi : v - hi;
    }
}

int launch_pack_feat_tc(nma_handle_s* h, const float* params, bool need_bwd, cudaStream_t st) {
    FeatPackArgs a;
    for (int i = 0; i < h->cfg.F; ++i) {
        for (int l = 0; l < 4; ++l) a.w[i][l] = params + h->po[i].featw[l];
        a.out[i] = h->ws[i].wtc_feat;
    }
    a.Cf_in = h->Cf_in;
    a.with_bwd = need_bwd ? 1 : 0;
    k_tc_pack_feat<<<h->cfg.F * (need_bwd ? 8 : 4), 256, 0, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
struct FeatTcArgs {
    const float* wpk[NMA_MAX_FLOWS];          // [8][FT_WLAYER_F] packed kernels (k_tc_pack_feat)
    const float* bias[NMA_MAX_FLOWS][4];
    float* a[NMA_MAX_FLOWS][5];               // a0 [p][Cf_in][LP], a1..a4 [p][50][LP]
    float* tin_hi[NMA_MAX_FLOWS];             // [14][tin_Q][4] conv operand (nma_tc.cuh)
    float* tin_lo[NMA_MAX_FLOWS];
    long long tin_Q[NMA_MAX_FLOWS];
    int Lin[NMA_MAX_FLOWS], LP[NMA_MAX_FLOWS];
    int cta_begin[NMA_MAX_FLOWS + 1];         // CTAs [cta_begin[i], cta_begin[i+1]) work on flow i
    float* x0;                                // ws[0].x: aligned copy of eps
    int XP0, p, L0, K, Cf_in, feat_off, save, F, ks0;
    int bf;                                   // conv operand in the bf16 split [8][tin_Q][8 x bf16] (TcP<true>)
    int diag;                                 // NMA_DIAG timing experiments (results invalid): 16 no activation saves, 32 no ELU,
                                              // 64 no next-layer operand stores, 128 no MMAs
};

__device__ __forceinline__ void ft_split_store(float* hi_dst, float* lo_dst, float a, float b, float c, float d) {
    const float4 h4 = make_float4(tf32_hi(a), tf32_hi(b), tf32_hi(c), tf32_hi(d));
    *reinterpret_cast<float4*>(hi_dst) = h4;
    *reinterpret_cast<float4*>(lo_dst) = make_float4(a - h4.x, b - h4.y, c - h4.z, d - h4.w);
}

// one dense layer on the tensor cores: D[128 x 64 | 64] = A_hi x [B_hi | B_lo]  (+)  A_lo x B_hi into the
// correction columns; issued by one elected lane of a converged warp, completion on `bar`.
__device__ __forceinline__ void ft_issue_layer(uint32_t a_hi_u, uint32_t a_lo_u, uint32_t w_u, int nks, uint32_t tmem_d,
                                               uint64_t* bar) {
    constexpr uint32_t idesc = umma_idesc_tf32(FT_M, TC_N, 0, 0);
    constexpr uint32_t idesc_wide = umma_idesc_tf32(FT_M, 2 * TC_N, 0, 0);
    constexpr uint32_t lbo = FT_M * 16u;                   // between 4-channel chunks (A: 128 positions, B: 128 rows)
    constexpr uint32_t ks_step = 2u * FT_M;                // 2 chunks, in 16-byte units
    tc_fence_after();
    if (elect_one()) {
        const uint32_t ah0 = desc_lo(a_hi_u, lbo), al0 = desc_lo(a_lo_u, lbo), b0 = desc_lo(w_u, lbo);
        const uint32_t hi32 = desc_hi(128u);
        for (int ks = 0; ks < nks; ++ks) {
            const uint64_t ah = desc_pack(ah0 + (uint32_t)ks * ks_step, hi32);
            const uint64_t al = desc_pack(al0 + (uint32_t)ks * ks_step, hi32);
            const uint64_t bw = desc_pack(b0 + (uint32_t)ks * ks_step, hi32);
            umma_tf32(tmem_d, ah, bw, idesc_wide, ks ? 1u : 0u);
            umma_tf32(tmem_d + TC_N, al, bw, idesc, 1u);
        }
        tc_commit(bar);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(FT_THREADS, 1) k_feat_fwd_tc(FeatTcArgs fa, SeriesView sv,
                                                               const int64_t* __restrict__ idx,
                                                               const float* __restrict__ eps) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t acc_bar[FT_SLOTS];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = tid / FT_SLOT_THREADS;
    const int wslot = (tid % FT_SLOT_THREADS) >> 5;
    const int quarter = wslot & 3, half = wslot >> 2;
    const int pos = quarter * 32 + lane;                   // position inside the tile == TMEM lane

    int i = 0;
    while (i + 1 < fa.F && (int)blockIdx.x >= fa.cta_begin[i + 1]) ++i;
    const int cta_local = (int)blockIdx.x - fa.cta_begin[i], ncta = fa.cta_begin[i + 1] - fa.cta_begin[i];

    const int w0_f = 2 * fa.ks0 * FT_CHUNK_F;
    float* Wl = smem;                                      // layers 1..3
    float* W0 = Wl + 3 * FT_WLAYER_F;                      // layer 0: the first 2*ks0 chunks only
    float* Abase = W0 + w0_f;
    float* A_hi = Abase + (size_t)slot * 2 * FT_A_F;
    float* A_lo = A_hi + FT_A_F;
    float* bias_sm = Abase + (size_t)FT_SLOTS * 2 * FT_A_F;   // [4][64]
    float* xchg = bias_sm + 256 + slot * FT_M;                // [128]

    // aligned copy of the base sample (x^(0) = eps, AR.py:31-32)
    {
        const long long n = (long long)fa.p * fa.XP0;
        for (long long t = (long long)blockIdx.x * blockDim.x + tid; t < n; t += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(t / fa.XP0), c = (int)(t - (long long)r * fa.XP0);
            fa.x0[t] = (c < fa.L0) ? eps[(size_t)r * fa.L0 + c] : 0.f;
        }
    }
    {
        const float4* src = reinterpret_cast<const float4*>(fa.wpk[i]);
        float4* d13 = reinterpret_cast<float4*>(Wl);
        for (int t = tid; t < 3 * FT_WLAYER_F / 4; t += blockDim.x) d13[t] = __ldg(src + FT_WLAYER_F / 4 + t);
        float4* d0 = reinterpret_cast<float4*>(W0);
        for (int t = tid; t < w0_f / 4; t += blockDim.x) d0[t] = __ldg(src + t);
        for (int t = tid; t < 256; t += blockDim.x) {
            const int l = t >> 6, f = t & 63;
            bias_sm[t] = (f < NMA_C) ? fa.bias[i][l][f] : 0.f;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < FT_SLOTS; ++s) mbar_init(&acc_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, FT_TMEM_COLS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot + (uint32_t)(slot * 2 * TC_N);
    const uint32_t ta = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 32);
    const uint32_t a_hi_u = smem_u32(A_hi), a_lo_u = smem_u32(A_lo);
    const uint32_t w0_u = smem_u32(W0), wl_u = smem_u32(Wl);

    const int Lin = fa.Lin[i], LP = fa.LP[i];
    const long long qtot = (long long)fa.p * Lin;
    const long long ntiles = (qtot + FT_M - 1) / FT_M;
    const int bar_id = 1 + slot;
    uint32_t phase = 0;

    for (long long tile = (long long)cta_local * FT_SLOTS + slot; tile < ntiles; tile += (long long)ncta * FT_SLOTS) {
        const long long q = tile * FT_M + pos;
        const bool valid = q < qtot;
        const int r = valid ? (int)(q / Lin) : 0;
        const int j = valid ? (int)(q - (long long)r * Lin) : 0;

        // ---- A1: gather the raw features of this position (slot = i*K + j + feat_off; AR.py:192-193 then :53) ----
        {
            const long long apos = (long long)sv.D * idx[r] + (long long)i * fa.K + j + fa.feat_off;
            float* ga0 = fa.a[i][0] + (size_t)r * fa.Cf_in * LP + j;
            for (int cc = half * fa.ks0; cc < (half + 1) * fa.ks0; ++cc) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * cc + e;
                    float x = 0.f;
                    if (valid && c < fa.Cf_in) {
                        x = (c < sv.Cf) ? series_val(sv, c, apos)
                                        : series_val(sv, c - sv.Cf, apos) - series_val(sv, c - sv.Cf, apos - 1);
                        if (fa.save) ga0[(size_t)c * LP] = x;
                    }
                    v[e] = x;
                }
                const size_t o = ((size_t)cc * FT_M + pos) * 4;
                ft_split_store(A_hi + o, A_lo + o, v[0], v[1], v[2], v[3]);
            }
        }
        tc_fence_before();
        fence_proxy_async();
        bar_sync_named(bar_id, FT_SLOT_THREADS);

        for (int l = 0; l < 4; ++l) {
            if (wslot == 0 && (fa.diag & 128)) {
                if (elect_one()) tc_commit(&acc_bar[slot]);
                __syncwarp();
            } else if (wslot == 0)
                ft_issue_layer(a_hi_u, a_lo_u, l == 0 ? w0_u : wl_u + (uint32_t)((l - 1) * FT_WLAYER_F * 4),
                               l == 0 ? fa.ks0 : TC_CCH / 2, tmem, &acc_bar[slot]);
            mbar_wait_backoff(&acc_bar[slot], phase);
            phase ^= 1u;
            tc_fence_after();

            // ---- epilogue: a_{l+1} = elu(acc + correction + bias) for 32 output channels of this position ----
            float v[32];
            {
                float c2[32];
                tmem_ld32(ta, v);
                tmem_ld32(ta + TC_N, c2);
                const float* b = bias_sm + l * 64 + half * 32;
#pragma unroll
                if (fa.diag & 32) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = v[k] + c2[k] + b[k];
                } else
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = elu_f(v[k] + c2[k] + b[k]);   // rows past the last position: finite, never stored
            }
            if (fa.save && valid && !(fa.diag & 16)) {
                float* dst = fa.a[i][l + 1] + ((size_t)r * NMA_C + half * 32) * LP + j;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (half * 32 + k < NMA_C) dst[(size_t)k * LP] = v[k];
            }
            if (l < 3 && (fa.diag & 64)) {
                bar_sync_named(bar_id, FT_SLOT_THREADS);
            } else if (l < 3) {
                // next layer's A operand: channels 32*half + 4cc .. +3 -> chunk 8*half + cc (channels >= 50 are zero)
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    if (half == 0 || cc < 6) {
                        const size_t o = ((size_t)(half * 8 + cc) * FT_M + pos) * 4;
                        ft_split_store(A_hi + o, A_lo + o, v[4 * cc], v[4 * cc + 1], v[4 * cc + 2], v[4 * cc + 3]);
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                bar_sync_named(bar_id, FT_SLOT_THREADS);
            } else {
                // conv operand: channel 0 = the flow's input sample (eps for flow 0; flows > 0 receive it from the
                // previous flow's epilogue), channel 1 + f = feature f.  The one-channel shift makes chunk 8 start
                // with feature 31, which lives in the other half's registers: pass it through shared memory.
                if (half == 0) xchg[pos] = v[31];
                tc_fence_before();
                bar_sync_named(bar_id, FT_SLOT_THREADS);
                if (fa.tin_hi[i] && valid) {
                    const float first = half ? xchg[pos] : ((i == 0) ? eps[(size_t)r * fa.L0 + j] : 0.f);
                    float* th = fa.tin_hi[i];
                    float* tl = fa.tin_lo[i];
                    const long long Q = fa.tin_Q[i];
                    if (fa.bf) {
                        // channel 32*half + 8*c8 + e: e0 = `first` for the half's first channel, else feature (channel - 1)
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            if (half == 0 || c8 < 3) {
                                float w8[8];
                                w8[0] = (c8 == 0) ? first : v[8 * c8 - 1 + (c8 == 0)];
#pragma unroll
                                for (int e = 1; e < 8; ++e) w8[e] = v[8 * c8 + e - 1];
                                uint4 h4, l4;
                                bf_split8(w8, h4, l4);
                                const size_t o = (size_t)(half * 4 + c8) * Q + q;
                                reinterpret_cast<uint4*>(th)[o] = h4;
                                reinterpret_cast<uint4*>(tl)[o] = l4;
                            }
                        }
                    } else
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        if (half == 0 || cc < 6) {
                            const size_t o = ((size_t)(half * 8 + cc) * Q + q) * 4;
                            const float e0 = (cc == 0) ? first : v[(4 * cc - 1) & 31];
                            ft_split_store(th + o, tl + o, e0, v[4 * cc], v[4 * cc + 1], v[4 * cc + 2]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, FT_TMEM_COLS);
}

// CTAs per flow in proportion to the flow's tile count (every flow gets at least one)
static void feat_tc_partition(const nma_handle_s* h, int p, int G, int* cta_begin) {
    const int F = h->cfg.F;
    long long tiles[NMA_MAX_FLOWS], total = 0;
    for (int i = 0; i < F; ++i) {
        tiles[i] = ((long long)p * h->fd[i].Lin + FT_M - 1) / FT_M;
        total += tiles[i];
    }
    int used = 0;
    cta_begin[0] = 0;
    for (int i = 0; i < F; ++i) {
        int n = (int)((double)G * (double)tiles[i] / (double)total + 0.5);
        const int left_flows = F - 1 - i;
        if (n < 1) n = 1;
        if (used + n + left_flows > G) n = G - used - left_flows;
        if (i == F - 1) n = G - used;
        if (n < 1) n = 1;
        used += n;
        cta_begin[i + 1] = used;
    }
}

int launch_feat_fwd_tc(nma_handle_s* h, const float* params, const int64_t* idx, const float* eps, int p, bool save,
                       cudaStream_t st) {
    FeatTcArgs fa;
    memset(&fa, 0, sizeof(fa));
    const int F = h->cfg.F;
    for (int i = 0; i < F; ++i) {
        fa.wpk[i] = h->ws[i].wtc_feat;
        for (int l = 0; l < 4; ++l) fa.bias[i][l] = params + h->po[i].featb[l];
        for (int l = 0; l < 5; ++l) fa.a[i][l] = h->ws[i].a[l];
        fa.tin_hi[i] = h->ws[i].tin_hi;
        fa.tin_lo[i] = h->ws[i].tin_lo;
        fa.tin_Q[i] = h->ws[i].tin_Q;
        fa.Lin[i] = h->fd[i].Lin;
        fa.LP[i] = h->fd[i].LP;
    }
    int G = h->sm_count;
    if (G < F) G = F;
    feat_tc_partition(h, p, G, fa.cta_begin);
    fa.x0 = h->ws[0].x;
    fa.XP0 = (h->fd[0].L + 3) & ~3;
    fa.p = p; fa.L0 = h->L0; fa.K = h->cfg.K; fa.Cf_in = h->Cf_in; fa.feat_off = h->feat_off;
    fa.save = save ? 1 : 0; fa.F = F; fa.ks0 = (h->Cf_in + 7) / 8;
    fa.bf = h->use_bf16;
    fa.diag = nma_diag_bits();
    const int smem = (3 * FT_WLAYER_F + 2 * fa.ks0 * FT_CHUNK_F + FT_SLOTS * 2 * FT_A_F + 256 + FT_SLOTS * FT_M) * 4;
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_feat_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    SeriesView sv = nma_series_view(h);
    k_feat_fwd_tc<<<G, FT_THREADS, smem, st>>>(fa, sv, idx, eps);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// backward of the feature MLP on the tensor cores (AR.py:53-56 differentiated; replaces k_feat_bwd, nma_bwd.cu)
//
// Per tile of 128 flattened positions, layers l = 3..0 (G_l = gradient w.r.t. the layer's pre-activation):
//     G_3 = df (.) elu'(a_4);    G_{l-1} = dA_l (.) elu'(a_l)
//     data gradient    dA_l[pos][f] = sum_g G_l[pos][g] W_l[f][g]      M = positions, N = 64, K = 56: the forward
//                      layer's MMA sequence with the transposed packed kernel (streamed through one 28 KB buffer by
//                      the TMA engine while the previous layer is being processed)
//     weight gradient  gW_l[f][g]  += sum_pos a_l[pos][f] G_l[pos][g]  reduction over positions: both operands K-major
//                      in 16-byte units of 4 consecutive positions, [position/4][row][4].  Rows are STACKED hi over lo:
//                      A = [a_hi (64 rows); a_lo (64 rows)], B = [G_hi | G_lo], so ONE 128x128x8 MMA yields all four
//                      products of the split; the drain adds the quadrants.  Row 50 of A is all ones: its
//                      accumulator row is the bias gradient sum_pos G_l, for free.
// The tensor core accumulates with truncation, so the weight-gradient chain is kept to the 16 MMAs of one tile:
// after every tile-layer the accumulator is drained TMEM -> registers (16 per thread and layer) and summed there
// in round-to-nearest fp32; one atomicAdd per weight and CTA at the end.
// 512 threads: thread = (position, group of 16 channels); warp w owns TMEM lanes 32*(w%4).. and columns 16*(w/4)...
// ---------------------------------------------------------------------------
#define FB_LBO 2064                              // bytes between position chunks: 128 rows x 16 B + 16 (bank spread)
#define FB_OPW_F (32 * FB_LBO / 4)               // floats of one weight-gradient operand tile
#define FB_THREADS 512

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// sum of the main and the correction accumulator (64 columns to the right) for 16 columns, 8 at a time (register budget)
template <typename F>
__device__ __forceinline__ void tmem_sum16(uint32_t taddr, F&& consume) {
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
        float v[8], c2[8];
        tmem_ld8(taddr + 8 * h8, v);
        tmem_ld8(taddr + TC_N + 8 * h8, c2);
#pragma unroll
        for (int k = 0; k < 8; ++k) consume(8 * h8 + k, v[k] + c2[k]);
    }
}

struct FeatBwdTcArgs {
    const float* wpk;        // [8][FT_WLAYER_F]; slots 4..7 = transposed kernels of layers 0..3
    const float* act[5];     // a0 [p][Cf_in][LP], a1..a4 [p][50][LP]
    const float* df;         // [p][50][LP]  d objective / d a4
    float* gw[4];
    float* gb[4];
    int Lin, LP, Cf_in, p;
};

__global__ void __launch_bounds__(FB_THREADS, 1) k_feat_bwd_tc(FeatBwdTcArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t dbar, wbar, wt_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quarter = warp & 3, cg = warp >> 2;
    const int pos = quarter * 32 + lane;
    float* Ad_hi = smem;                         // data-gradient A operand [14][128][4]
    float* Ad_lo = Ad_hi + FT_A_F;
    float* Aw = Ad_lo + FT_A_F;                  // weight-gradient A operand: rows f (hi) / 64 + f (lo), K = positions
    float* Bw = Aw + FB_OPW_F;                   // weight-gradient B operand: rows g (hi) / 64 + g (lo)
    float* Wt = Bw + FB_OPW_F;                   // transposed packed kernel of the current layer

    if (tid == 0) {
        mbar_init(&dbar, 1);
        mbar_init(&wbar, 1);
        mbar_init(&wt_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, FT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot, tmem_w = tmem_slot + 2 * TC_N;
    const uint32_t td = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(16 * cg);
    const uint32_t tw = tmem_w + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(16 * cg);
    const uint32_t ad_hi_u = smem_u32(Ad_hi), ad_lo_u = smem_u32(Ad_lo), aw_u = smem_u32(Aw), bw_u = smem_u32(Bw),
                   wt_u = smem_u32(Wt);

    const int Lin = a.Lin, LP = a.LP;
    const long long qtot = (long long)a.p * Lin;
    const long long ntiles = (qtot + FT_M - 1) / FT_M;
    uint32_t dph = 0, wph = 0, tph = 0;
    float acc[4][16];
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[l][k] = 0.f;

    auto load_wt = [&](int l) {     // one thread: stream the transposed kernel of layer l into Wt
        mbar_expect_tx(&wt_bar, FT_WLAYER_F * 4u);
        bulk_g2s(Wt, a.wpk + (size_t)(4 + l) * FT_WLAYER_F, FT_WLAYER_F * 4u, &wt_bar);
    };
    if (tid == 0 && (long long)blockIdx.x < ntiles) load_wt(3);

    // this thread's 16 rows inside a weight-gradient operand tile (float offsets): hi row f, lo row 64 + f
    const int wofs = (pos >> 2) * (FB_LBO / 4) + (pos & 3);
    bool first_tile = true;

    // 16 channels [16*cg, 16*cg + 16) of activation a_l at one position (zero outside the layer's width / the tile)
    auto load_act = [&](float (&dst)[16], int l, bool ok, int r, int j) {
        const int nin = (l == 0) ? a.Cf_in : NMA_C;
        const float* src = a.act[l] + ((size_t)r * nin + 16 * cg) * LP + j;
        // one branch on the lane's validity around 16 loads under warp-uniform predicates (a select per element compiles
        // into a branch per element, which keeps the loads from being issued back to back)
        if (ok) {
            // running pointer: two integer instructions per load instead of a 64-bit multiply-add chain each
            const char* pb = reinterpret_cast<const char*>(src);
            const size_t step = (size_t)LP * sizeof(float);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                dst[k] = (16 * cg + k < nin) ? __ldg(reinterpret_cast<const float*>(pb)) : 0.f;
                pb += step;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) dst[k] = 0.f;
        }
    };
    auto load_raw = [&](float (&dst)[16], const float* base, bool ok, int r, int j) {
        const float* src = base + ((size_t)r * NMA_C + 16 * cg) * LP + j;
        if (ok) {
            const char* pb = reinterpret_cast<const char*>(src);
            const size_t step = (size_t)LP * sizeof(float);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                dst[k] = (16 * cg + k < NMA_C) ? *reinterpret_cast<const float*>(pb) : 0.f;
                pb += step;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) dst[k] = 0.f;
        }
    };
    float g[16], al[16];
    {   // first tile: raw df and a_4
        const long long q = (long long)blockIdx.x * FT_M + pos;
        const bool ok = q < qtot;
        const int r = ok ? (int)(q / Lin) : 0;
        const int j = ok ? (int)(q - (long long)r * Lin) : 0;
        load_raw(g, a.df, ok, r, j);
        load_act(al, 4, ok, r, j);
    }

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long q = tile * FT_M + pos;
        const bool valid = q < qtot;
        const int r = valid ? (int)(q / Lin) : 0;
        const int j = valid ? (int)(q - (long long)r * Lin) : 0;
        const bool more = tile + gridDim.x < ntiles;

        // G_3 = df (.) elu'(a_4): the raw values were prefetched into g / al during layer 0 of the previous tile
#pragma unroll
        for (int k = 0; k < 16; ++k) g[k] = (valid && (16 * cg + k < NMA_C)) ? g[k] * elu_grad_from_out(al[k]) : 0.f;
        load_act(al, 3, valid, r, j);
#pragma unroll
        for (int l = 3; l >= 0; --l) {
            // Order inside a layer: the data-gradient operand first, so that its MMAs - the ones the next layer waits for -
            // are queued behind the previous layer's weight-gradient MMAs while the threads are still busy with that
            // weight gradient's accumulator and with the next weight-gradient operands.
            if (l > 0) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int ch = 4 * cg + cc;
                    if (ch < TC_CCH) {
                        const size_t o = ((size_t)ch * FT_M + pos) * 4;
                        ft_split_store(Ad_hi + o, Ad_lo + o, g[4 * cc], g[4 * cc + 1], g[4 * cc + 2], g[4 * cc + 3]);
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                __syncthreads();
                if (warp == 0) {
                    mbar_wait_backoff(&wt_bar, tph);
                    tph ^= 1u;
                    ft_issue_layer(ad_hi_u, ad_lo_u, wt_u, TC_CCH / 2, tmem_d, &dbar);
                }
            }
            // the previous weight-gradient MMAs still read Aw / Bw: wait for them, then drain their accumulator
            if (l < 3 || !first_tile) {
                mbar_wait_backoff(&wbar, wph);
                wph ^= 1u;
                tc_fence_after();
                tmem_sum16(tw, [&](int k, float x) { acc[(l + 1) & 3][k] += x; });
            }
            // ---- weight-gradient operands of layer l (al holds a_l, prefetched one layer ahead) ----
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int row = 16 * cg + k;
                const int o_hi = wofs + (row >> 3) * 32 + (row & 7) * 4;
                const int o_lo = o_hi + 8 * 32;                    // row + 64
                const float gh = tf32_hi(g[k]);
                Bw[o_hi] = gh;
                Bw[o_lo] = g[k] - gh;
                const float av = (row == NMA_C) ? 1.f : al[k];       // row 50: ones -> bias gradient
                const float ah = tf32_hi(av);
                Aw[o_hi] = ah;
                Aw[o_lo] = av - ah;
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                    constexpr uint32_t idesc_wide = umma_idesc_tf32(FT_M, 2 * TC_N, 0, 0);
                    const uint32_t a0 = desc_lo(aw_u, FB_LBO), b0 = desc_lo(bw_u, FB_LBO);
                    const uint32_t hi32 = desc_hi(128u);
#pragma unroll 4
                    for (int ks = 0; ks < FT_M / 8; ++ks) {
                        const uint32_t step = (uint32_t)ks * (2u * FB_LBO / 16u);
                        umma_tf32(tmem_w, desc_pack(a0 + step, hi32), desc_pack(b0 + step, hi32), idesc_wide, ks ? 1u : 0u);
                    }
                    tc_commit(&wbar);
                }
                __syncwarp();
            }
            // ---- prefetch while the MMAs run: a_{l-1}, or (after layer 0) the raw df / a_4 of the next tile ----
            if (l > 0) {
                load_act(al, l - 1, valid, r, j);
            } else {
                const long long qn = (tile + gridDim.x) * FT_M + pos;
                const bool vn = qn < qtot;
                const int rn = vn ? (int)(qn / Lin) : 0;
                const int jn = vn ? (int)(qn - (long long)rn * Lin) : 0;
                load_raw(g, a.df, vn, rn, jn);
                load_act(al, 4, vn, rn, jn);
            }
            if (l > 0) {
                mbar_wait_backoff(&dbar, dph);
                dph ^= 1u;
                tc_fence_after();
                // the data-gradient MMAs are done with Wt: fetch the next layer's kernel (layer 3 of the next tile after layer 1)
                if (tid == 0) {
                    if (l > 1) load_wt(l - 1);
                    else if (more) load_wt(3);
                }
                // G_{l-1} = dA_l (.) elu'(a_l); a_l is read back from the operand tile (hi + lo is exact), the
                // registers that held it already carry the prefetch
                tmem_sum16(td, [&](int k, float x) {
                    const int row = 16 * cg + k;
                    const int o_hi = wofs + (row >> 3) * 32 + (row & 7) * 4;
                    const float e = Aw[o_hi] + Aw[o_hi + 8 * 32];
                    g[k] = x * elu_grad_from_out(e);
                });
            }
        }
        first_tile = false;
    }
    if (!first_tile) {   // drain the last tile's layer 0
        mbar_wait_backoff(&wbar, wph);
        tc_fence_after();
        tmem_sum16(tw, [&](int k, float x) { acc[0][k] += x; });
    }
    // TMEM lane = operand row: f (hi rows 0..63) or 64 + f (lo rows); column 16*cg + k = output channel g
    {
        const int f = pos & 63;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int nin = (l == 0) ? a.Cf_in : NMA_C;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int gc = 16 * cg + k;
                if (gc < NMA_C) {
                    if (f < nin) atomicAdd(a.gw[l] + f * NMA_C + gc, acc[l][k]);
                    else if (pos == NMA_C) atomicAdd(a.gb[l] + gc, acc[l][k]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, FT_TMEM_COLS);
}

int launch_feat_bwd_tc(nma_handle_s* h, int i, const float* params, int p, float* gp, cudaStream_t st) {
    (void)params;
    const FlowDims& d = h->fd[i];
    FeatBwdTcArgs a;
    a.wpk = h->ws[i].wtc_feat;
    for (int l = 0; l < 5; ++l) a.act[l] = h->ws[i].a[l];
    a.df = h->ws[i].df;
    for (int l = 0; l < 4; ++l) {
        a.gw[l] = gp + h->po[i].featw[l];
        a.gb[l] = gp + h->po[i].featb[l];
    }
    a.Lin = d.Lin; a.LP = d.LP; a.Cf_in = h->Cf_in; a.p = p;
    const long long ntiles = ((long long)p * d.Lin + FT_M - 1) / FT_M;
    const int grid = (int)(ntiles < h->sm_count ? ntiles : h->sm_count);
    const int smem = (2 * FT_A_F + 2 * FB_OPW_F + FT_WLAYER_F) * 4;
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_feat_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    k_feat_bwd_tc<<<grid, FB_THREADS, smem, st>>>(a);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// backward of the flow layer's head and hidden 1x1 layer on the tensor cores (replaces k_epi_bwd, nma_bwd.cu, for the
// configuration of the AR scripts: one hidden layer, no batch-norm, flow_dims = 1; AR.py:74-89 differentiated).
//
// Per position (r, m) of the conv output: the affine flow layer and the softplus head give d objective / d (mu, s)
// (SIMT, a handful of flops); G = gradient w.r.t. the hidden layer's pre-activation = (dmu hw[.,0] + ds hw[.,1]) elu'(e_1);
// then exactly one layer of k_feat_bwd_tc: data gradient G W^T on tcgen05 (-> dA = . elu'(e_0), written straight
// into the conv gradient operand `dat`, hi/lo split), weight gradient e_0^T G as one stacked 128x128x8 MMA per 8
// positions, bias gradient through the ones row.  Head weight gradients are register partial sums per thread,
// flushed once.  The per-row sums of dA the theta-bias backward needs are taken from `dat` by k_dtb_from_dat.
// ---------------------------------------------------------------------------
struct EpiBwdTcArgs {
    const float* wpk;        // transposed packed hidden kernel [14][128][4] (k_tc_pack_w1x1)
    const float* headw;      // [50][2]
    const float* e0;         // [p][50][NP]  elu(conv + theta-bias)
    const float* e1;         // [p][50][NP]  output of the hidden layer
    const float* s;          // [p][NP]
    const float* x_in;       // [p][XP]
    const float* dx_next;    // [p][XPn]
    float* dx;               // [p][XP]
    float* dat_hi;           // [14][dat_Q][4], dA(r, m) at q = K-1 + r*Lin + m
    float* dat_lo;
    long long dat_Q;
    float* g_hidw; float* g_hidb; float* g_headw; float* g_headb;
    int XP, XPn, N, NP, K, S, Lin, p;
    float cq;
    int bf;                  // `dat` in the bf16 split [8][dat_Q][8 x bf16]
    float* dtb;              // [p][50] per-row sums of dA, accumulated here when non-null (zeroed by the launcher)
};

__global__ void __launch_bounds__(FB_THREADS, 1) k_epi_bwd_tc(EpiBwdTcArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t dbar, wbar;
    __shared__ uint32_t tmem_slot;
    __shared__ float hw[2 * 64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quarter = warp & 3, cg = warp >> 2;
    const int pos = quarter * 32 + lane;
    float* Ad_hi = smem;
    float* Ad_lo = Ad_hi + FT_A_F;
    float* Aw = Ad_lo + FT_A_F;
    float* Bw = Aw + FB_OPW_F;
    float* Wt = Bw + FB_OPW_F;

    // dx[r][0..K) = 0 (the affine layer only touches slots >= K); positions of `dat` past the last row read as zero
    for (long long t = (long long)blockIdx.x * blockDim.x + tid; t < (long long)a.p * a.K; t += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(t / a.K), j = (int)(t - (long long)r * a.K);
        a.dx[(size_t)r * a.XP + j] = 0.f;
    }
    if (blockIdx.x == 0) {
        const long long q_lo = (long long)a.p * a.Lin + (a.K - 1);
        for (int t = tid; t < (a.bf ? 8 : TC_CCH) * 128; t += blockDim.x) {      // same 16-byte units in both formats
            const int fch = t / 128, q = t - fch * 128;
            if (q_lo + q < a.dat_Q) {
                const size_t o = ((size_t)fch * a.dat_Q + q_lo + q) * 4;
                *reinterpret_cast<float4*>(a.dat_hi + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(a.dat_lo + o) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    {
        const float4* src = reinterpret_cast<const float4*>(a.wpk);
        float4* dst = reinterpret_cast<float4*>(Wt);
        for (int t = tid; t < FT_WLAYER_F / 4; t += blockDim.x) dst[t] = __ldg(src + t);
        if (tid < 128) hw[tid] = (tid < 2 * NMA_C) ? a.headw[tid] : 0.f;
    }
    if (tid == 0) {
        mbar_init(&dbar, 1);
        mbar_init(&wbar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, FT_TMEM_COLS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot, tmem_w = tmem_slot + 2 * TC_N;
    const uint32_t td = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(16 * cg);
    const uint32_t tw = tmem_w + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(16 * cg);
    const uint32_t ad_hi_u = smem_u32(Ad_hi), ad_lo_u = smem_u32(Ad_lo), aw_u = smem_u32(Aw), bw_u = smem_u32(Bw),
                   wt_u = smem_u32(Wt);

    const int N = a.N, NP = a.NP;
    const long long qtot = (long long)a.p * N;
    const long long ntiles = (qtot + FT_M - 1) / FT_M;
    uint32_t dph = 0, wph = 0;
    float acc[16], hacc0[16], hacc1[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { acc[k] = 0.f; hacc0[k] = 0.f; hacc1[k] = 0.f; }
    float hb0 = 0.f, hb1 = 0.f;
    const int wofs = (pos >> 2) * (FB_LBO / 4) + (pos & 3);
    bool first_tile = true;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long q = tile * FT_M + pos;
        const bool valid = q < qtot;
        const int r = valid ? (int)(q / N) : 0;
        const int m = valid ? (int)(q - (long long)r * N) : 0;
        // ---- affine flow layer + softplus head (AR.py:83-88) ----
        float dmu = 0.f, dsr = 0.f;
        if (valid) {
            const float dxo = a.dx_next[(size_t)r * a.XPn + m];
            const float sr = a.s[(size_t)r * NP + m];
            const float sigma = softplus_f(sr) + 1e-10f;
            const float xin = a.x_in[(size_t)r * a.XP + m + a.K];
            float dsig = dxo * xin;
            if (m >= N - a.S) dsig -= a.cq / sigma;        // logq -= log sigma over the last S slots
            dmu = dxo;
            dsr = dsig * sigmoid_f(sr);
            if (cg == 0) a.dx[(size_t)r * a.XP + m + a.K] = dxo * sigma;
        }
        float g[16], e0v[16];
        {
            const float* p1 = a.e1 + ((size_t)r * NMA_C + 16 * cg) * NP + m;
            const float* p0 = a.e0 + ((size_t)r * NMA_C + 16 * cg) * NP + m;
            // one branch on the lane's validity around the 32 loads (warp-uniform channel predicates inside): a select
            // per element compiles into a branch per element and the loads are no longer issued back to back
            float ehv[16];
            if (valid) {
                const char* pb1 = reinterpret_cast<const char*>(p1);
                const char* pb0 = reinterpret_cast<const char*>(p0);
                const size_t step = (size_t)NP * sizeof(float);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const bool ok = 16 * cg + k < NMA_C;
                    ehv[k] = ok ? __ldg(reinterpret_cast<const float*>(pb1)) : 0.f;
                    e0v[k] = ok ? __ldg(reinterpret_cast<const float*>(pb0)) : 0.f;
                    pb1 += step; pb0 += step;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) { ehv[k] = 0.f; e0v[k] = 0.f; }
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int ch = 16 * cg + k;          // hw is zero past channel 49, dmu = dsr = 0 on invalid lanes
                hacc0[k] = fmaf(ehv[k], dmu, hacc0[k]);
                hacc1[k] = fmaf(ehv[k], dsr, hacc1[k]);
                g[k] = fmaf(dmu, hw[2 * ch], dsr * hw[2 * ch + 1]) * elu_grad_from_out(ehv[k]);
            }
        }
        if (cg == 0) { hb0 += dmu; hb1 += dsr; }
        // ---- data gradient first: A = G [g/4][pos][4]; its MMAs queue behind the previous tile's weight-gradient MMAs
        // while the threads drain that weight gradient and stage the next one ----
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int ch = 4 * cg + cc;
            if (ch < TC_CCH) {
                const size_t o = ((size_t)ch * FT_M + pos) * 4;
                ft_split_store(Ad_hi + o, Ad_lo + o, g[4 * cc], g[4 * cc + 1], g[4 * cc + 2], g[4 * cc + 3]);
            }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) ft_issue_layer(ad_hi_u, ad_lo_u, wt_u, TC_CCH / 2, tmem_d, &dbar);
        if (!first_tile) {
            mbar_wait_backoff(&wbar, wph);
            wph ^= 1u;
            tc_fence_after();
            tmem_sum16(tw, [&](int k, float x) { acc[k] += x; });
        }
        // ---- weight gradient: A = [e_0; ones], B = G, K = positions ----
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int row = 16 * cg + k;
            const int o_hi = wofs + (row >> 3) * 32 + (row & 7) * 4;
            const int o_lo = o_hi + 8 * 32;
            const float gh = tf32_hi(g[k]);
            Bw[o_hi] = gh;
            Bw[o_lo] = g[k] - gh;
            const float av = (row == NMA_C) ? 1.f : e0v[k];
            const float ah = tf32_hi(av);
            Aw[o_hi] = ah;
            Aw[o_lo] = av - ah;
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                constexpr uint32_t idesc_wide = umma_idesc_tf32(FT_M, 2 * TC_N, 0, 0);
                const uint32_t a0 = desc_lo(aw_u, FB_LBO), b0 = desc_lo(bw_u, FB_LBO);
                const uint32_t hi32 = desc_hi(128u);
#pragma unroll 4
                for (int ks = 0; ks < FT_M / 8; ++ks) {
                    const uint32_t step = (uint32_t)ks * (2u * FB_LBO / 16u);
                    umma_tf32(tmem_w, desc_pack(a0 + step, hi32), desc_pack(b0 + step, hi32), idesc_wide, ks ? 1u : 0u);
                }
                tc_commit(&wbar);
            }
            __syncwarp();
        }
        mbar_wait_backoff(&dbar, dph);
        dph ^= 1u;
        tc_fence_after();
        // dA = (G W^T) (.) elu'(e_0) -> the conv gradient operand, channels 16*cg .. +15 of this position
        tmem_sum16(td, [&](int k, float x) { g[k] = x * elu_grad_from_out(e0v[k]); });
        if (a.dtb) {
            // per-row sums of dA for the theta-bias backward (AR.py:63-72): a warp holds 32 consecutive flattened
            // positions, i.e. one row or the end of one and the start of the next; anything else (rows shorter than a
            // warp) falls back to one atomic per lane
            const int rk = valid ? r : -1;
            const int r0 = __shfl_sync(0xffffffffu, rk, 0), r1 = __shfl_sync(0xffffffffu, rk, 31);
            const bool two = __all_sync(0xffffffffu, rk == r0 || rk == r1);
            if (two) {
                // 16 channel sums per row segment in 16 shuffles (warp_sum16_transposed) - one warp_sum per channel and
                // segment was 160 shuffles and a quarter of this kernel's instruction stream
                float part[16];
                int j;
#pragma unroll
                for (int k = 0; k < 16; ++k) part[k] = (rk == r0) ? g[k] : 0.f;
                const float s0 = warp_sum16_transposed(part, lane, &j);
                if (!(lane & 1) && r0 >= 0 && 16 * cg + j < NMA_C) atomicAdd(a.dtb + (size_t)r0 * NMA_C + 16 * cg + j, s0);
                if (r1 != r0) {                                   // warp-uniform: the warp straddles two rows
#pragma unroll
                    for (int k = 0; k < 16; ++k) part[k] = (rk == r1) ? g[k] : 0.f;
                    const float s1 = warp_sum16_transposed(part, lane, &j);
                    if (!(lane & 1) && r1 >= 0 && 16 * cg + j < NMA_C) atomicAdd(a.dtb + (size_t)r1 * NMA_C + 16 * cg + j, s1);
                }
            } else if (valid) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (16 * cg + k < NMA_C) atomicAdd(a.dtb + (size_t)r * NMA_C + 16 * cg + k, g[k]);
            }
        }
        if (valid) {
            const long long qd = (long long)(a.K - 1) + (long long)r * a.Lin + m;
            if (a.bf) {
                // channels 16*cg .. +15 = chunks 2*cg, 2*cg + 1 of 8 channels (channels >= 50 are exact zeros)
#pragma unroll
                for (int c8 = 0; c8 < 2; ++c8) {
                    float w8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) w8[e] = g[8 * c8 + e];
                    uint4 h4, l4;
                    bf_split8(w8, h4, l4);
                    const size_t o = (size_t)(2 * cg + c8) * a.dat_Q + qd;
                    reinterpret_cast<uint4*>(a.dat_hi)[o] = h4;
                    reinterpret_cast<uint4*>(a.dat_lo)[o] = l4;
                }
            } else
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int ch = 4 * cg + cc;
                if (ch < TC_CCH) {
                    const size_t o = ((size_t)ch * a.dat_Q + qd) * 4;
                    ft_split_store(a.dat_hi + o, a.dat_lo + o, g[4 * cc], g[4 * cc + 1], g[4 * cc + 2], g[4 * cc + 3]);
                }
            }
        }
        first_tile = false;
    }
    if (!first_tile) {
        mbar_wait_backoff(&wbar, wph);
        tc_fence_after();
        tmem_sum16(tw, [&](int k, float x) { acc[k] += x; });
    }
    {
        const int f = pos & 63;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int gc = 16 * cg + k;
            if (gc < NMA_C) {
                if (f < NMA_C) atomicAdd(a.g_hidw + f * NMA_C + gc, acc[k]);
                else if (pos == NMA_C) atomicAdd(a.g_hidb + gc, acc[k]);
                // head kernel gradient [50][2]: partial sums of this thread's positions, reduced over the warp first
                const float s0 = warp_sum(hacc0[k]), s1 = warp_sum(hacc1[k]);
                if (lane == 0) { atomicAdd(a.g_headw + 2 * gc, s0); atomicAdd(a.g_headw + 2 * gc + 1, s1); }
            }
        }
        if (cg == 0) {
            hb0 = warp_sum(hb0);
            hb1 = warp_sum(hb1);
            if (lane == 0) { atomicAdd(a.g_headb, hb0); atomicAdd(a.g_headb + 1, hb1); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, FT_TMEM_COLS);
}

// dtb[r][f] = sum_m dA[r][f][m] (upstream of the theta-bias MLP, AR.py:63-72) and the conv bias gradient, from the
// hi + lo parts of the gradient operand: one CTA per row, warp w owns channel chunk w, lanes along the positions
__global__ void __launch_bounds__(TC_CCH * 32) k_dtb_from_dat(const float* __restrict__ dat_hi,
                                                             const float* __restrict__ dat_lo, long long Q, int K,
                                                             int Lin, int N, float* __restrict__ dtb,
                                                             float* __restrict__ g_convb) {
    const int r = blockIdx.x, c4 = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t base = ((size_t)c4 * Q + (size_t)(K - 1) + (size_t)r * Lin) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = lane; m < N; m += 32) {
        const float4 h4 = __ldg(reinterpret_cast<const float4*>(dat_hi + base + (size_t)m * 4));
        const float4 l4 = __ldg(reinterpret_cast<const float4*>(dat_lo + base + (size_t)m * 4));
        acc.x += h4.x + l4.x; acc.y += h4.y + l4.y; acc.z += h4.z + l4.z; acc.w += h4.w + l4.w;
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
    if (lane == 0) {
        const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int f = 4 * c4 + e;
            if (f < NMA_C) {
                dtb[(size_t)r * NMA_C + f] = v[e];
                atomicAdd(g_convb + f, v[e]);
            }
        }
    }
}

// the same from the bf16-split operand [8][Q][8 x bf16]: warp w owns the 8 channels of chunk w
__global__ void __launch_bounds__(8 * 32) k_dtb_from_dat_bf(const uint4* __restrict__ dat_hi, const uint4* __restrict__ dat_lo,
                                                            long long Q, int K, int Lin, int N, float* __restrict__ dtb,
                                                            float* __restrict__ g_convb) {
    const int r = blockIdx.x, c8 = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t base = (size_t)c8 * Q + (size_t)(K - 1) + (size_t)r * Lin;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int m = lane; m < N; m += 32) {
        const uint4 h4 = __ldg(dat_hi + base + m), l4 = __ldg(dat_lo + base + m);
        const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[2 * e] += bf_to_float(hw[e] & 0xffffu) + bf_to_float(lw[e] & 0xffffu);
            acc[2 * e + 1] += bf_to_float(hw[e] >> 16) + bf_to_float(lw[e] >> 16);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = warp_sum(acc[e]);
    if (lane == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int f = 8 * c8 + e;
            if (f < NMA_C) {
                dtb[(size_t)r * NMA_C + f] = acc[e];
                atomicAdd(g_convb + f, acc[e]);
            }
        }
    }
}

// conv bias gradient from the per-row sums: g_convb[f] += sum_r dtb[r][f]
__global__ void __launch_bounds__(256) k_convb_from_dtb(const float* __restrict__ dtb, int p, float* __restrict__ g_convb) {
    __shared__ float part[4][64];
    const int f = threadIdx.x & 63, rl = threadIdx.x >> 6;
    float acc = 0.f;
    if (f < NMA_C)
        for (int r = blockIdx.x * 4 + rl; r < p; r += gridDim.x * 4) acc += __ldg(dtb + (size_t)r * NMA_C + f);
    part[rl][f] = acc;
    __syncthreads();
    if (rl == 0 && f < NMA_C) atomicAdd(g_convb + f, part[0][f] + part[1][f] + part[2][f] + part[3][f]);
}

// transposed pack of ONE 1x1 kernel [50][50] (same element map as the feature kernels' transposed slots)
__global__ void k_tc_pack_w1x1_t(const float* __restrict__ W, float* __restrict__ out) {
    for (int t = threadIdx.x; t < FT_WLAYER_F; t += blockDim.x) {
        const int e = t & 3, row = (t >> 2) & 127, cch = t >> 9;
        const int n = row & 63, c = 4 * cch + e;
        const float v = (c < NMA_C && n < NMA_C) ? W[n * NMA_C + c] : 0.f;
        const float hi = tf32_hi(v);
        out[t] = (row < 64) ? hi : v - hi;
    }
}

int epi_bwd_tc_supported(const nma_handle_s* h) {
    return h->use_tc && h->use_tc_feat && h->cfg.H == 1 && !h->cfg.bn && h->cfg.D == 1;
}

// transposed hidden 1x1 kernel of flow i -> slot 8 of the flow's pack buffer (part of the step's weight packing)
int launch_pack_w1x1_t(nma_handle_s* h, int i, const float* params, cudaStream_t st) {
    float* wt = h->ws[i].wtc_feat + (size_t)8 * FT_WLAYER_F;
    k_tc_pack_w1x1_t<<<1, 256, 0, st>>>(params + h->po[i].hidw[0], wt);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_epi_bwd_tc(nma_handle_s* h, int i, const float* params, int p, int objective, float* gp, cudaStream_t st) {
    const FlowDims& d = h->fd[i];
    // slot 8 of the flow's pack buffer: the transposed hidden kernel (launch_pack_w1x1_t; weights are constant during a step)
    float* wt = h->ws[i].wtc_feat + (size_t)8 * FT_WLAYER_F;
    EpiBwdTcArgs a;
    a.wpk = wt;
    a.headw = params + h->po[i].headw;
    a.e0 = h->ws[i].h[0]; a.e1 = h->ws[i].h[1];
    a.s = h->ws[i].s; a.x_in = h->ws[i].x; a.dx_next = h->ws[i + 1].dx; a.dx = h->ws[i].dx;
    a.dat_hi = h->ws[i].dat_hi; a.dat_lo = h->ws[i].dat_lo; a.dat_Q = h->ws[i].dat_Q;
    a.g_hidw = gp + h->po[i].hidw[0]; a.g_hidb = gp + h->po[i].hidb[0];
    a.g_headw = gp + h->po[i].headw; a.g_headb = gp + h->po[i].headb;
    a.XP = (d.L + 3) & ~3; a.XPn = (h->fd[i + 1].L + 3) & ~3; a.N = d.N; a.NP = d.NP; a.K = h->cfg.K; a.S = h->S;
    a.Lin = d.Lin; a.p = p;
    a.cq = (objective == NMA_OBJ_ELBO) ? (float)h->cfg.scale : 0.f;
    a.bf = h->use_bf16;
    a.dtb = h->ws[i].dtb;                                    // per-row sums of dA taken from registers in this kernel
    if (a.dtb) NMA_CHECK_CUDA(cudaMemsetAsync(a.dtb, 0, (size_t)p * NMA_C * 4, st));
    const long long ntiles = ((long long)p * d.N + FT_M - 1) / FT_M;
    const int grid = (int)(ntiles < h->sm_count ? ntiles : h->sm_count);
    const int smem = (2 * FT_A_F + 2 * FB_OPW_F + FT_WLAYER_F) * 4;
    static int configured = 0;
    if (configured < smem) {
        NMA_CHECK_CUDA(cudaFuncSetAttribute(k_epi_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    k_epi_bwd_tc<<<grid, FB_THREADS, smem, st>>>(a);
    if (a.dtb) {
        int g = (p + 3) / 4;
        if (g > h->sm_count) g = h->sm_count;
        k_convb_from_dtb<<<g, 256, 0, st>>>(a.dtb, p, gp + h->po[i].convb);
    } else if (h->use_bf16)
        k_dtb_from_dat_bf<<<p, 8 * 32, 0, st>>>((const uint4*)h->ws[i].dat_hi, (const uint4*)h->ws[i].dat_lo, h->ws[i].dat_Q,
                                               h->cfg.K, d.Lin, d.N, h->ws[i].dtb, gp + h->po[i].convb);
    else
        k_dtb_from_dat<<<p, TC_CCH * 32, 0, st>>>(h->ws[i].dat_hi, h->ws[i].dat_lo, h->ws[i].dat_Q, h->cfg.K, d.Lin, d.N,
                                                 h->ws[i].dtb, gp + h->po[i].convb);
    nma_count_launch(2);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}
