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
            if (wslot == 0)
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
                for (int k = 0; k < 32; ++k) v[k] = valid ? elu_f(v[k] + c2[k] + b[k]) : 0.f;
            }
            if (fa.save && valid) {
                float* dst = fa.a[i][l + 1] + ((size_t)r * NMA_C + half * 32) * LP + j;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (half * 32 + k < NMA_C) dst[(size_t)k * LP] = v[k];
            }
            if (l < 3) {
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
