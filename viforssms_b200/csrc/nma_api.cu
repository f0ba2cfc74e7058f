// C-ABI of the NMA ELBO step (include/nma_b200.h): handle, workspace arena, step orchestration.
#include <stdarg.h>
#include <string.h>
#include <new>
#include <stdlib.h>
#include "nma_tc.cuh"

int launch_feat_fwd_eps(nma_handle_s* h, const float* params, const int64_t* idx, const float* eps, int p, bool save,
                        cudaStream_t st);

static thread_local char g_err[512] = "";

void nma_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* nma_last_error(void) { return g_err; }
extern "C" int nma_version(void) { return 100; }

SeriesView nma_series_view(const nma_handle_s* h) {
    SeriesView sv;
    memset(&sv, 0, sizeof(sv));
    for (int i = 0; i < NMA_MAX_ARRAYS; ++i) {
        sv.base[i] = h->base[i];
        sv.len[i] = h->base_len[i];
    }
    for (int c = 0; c < NMA_MAX_CHAN; ++c) {
        sv.chan_array[c] = h->cfg.chan_array[c];
        sv.chan_offset[c] = h->cfg.chan_offset[c];
    }
    sv.Cf = h->cfg.Cf;
    sv.D = h->cfg.D;
    sv.feat_aug = h->cfg.feat_aug;
    return sv;
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static void compute_layout(nma_handle_s* h) {
    const nma_config& c = h->cfg;
    int64_t off = 0;
    for (int i = 0; i < c.F; ++i) {
        FlowParamOff& po = h->po[i];
        for (int l = 0; l < 4; ++l) {
            const int nout = (l == 3) ? h->feat_out[i] : NMA_C;
            po.featw[l] = off; off += (int64_t)(l == 0 ? h->Cf_in : NMA_C) * nout;
            po.featb[l] = off; off += nout;
        }
        po.convw = off; off += (int64_t)c.K * h->conv_cin * NMA_C;
        po.convb = off; off += NMA_C;
        for (int l = 0; l < 3; ++l) {
            po.thw[l] = off; off += (int64_t)(l == 0 ? c.dtheta : NMA_C) * NMA_C;
            po.thb[l] = off; off += NMA_C;
        }
        for (int l = 0; l < NMA_MAXH; ++l) { po.hidw[l] = po.hidb[l] = po.gam[l] = po.bet[l] = -1; }
        for (int l = 0; l < c.H; ++l) {
            po.hidw[l] = off; off += NMA_C * NMA_C;
            po.hidb[l] = off; off += NMA_C;
            if (c.bn) {
                po.gam[l] = off; off += NMA_C;
                po.bet[l] = off; off += NMA_C;
            }
        }
        po.headw = off; off += NMA_C * 2;
        po.headb = off; off += 2;
    }
    h->n_params = off;
}

static int bf16_path_ok(const nma_handle_s* h);

extern "C" int nma_create(const nma_config* cfg, nma_handle* out) {
    if (!cfg || !out) { nma_set_error("nma_create: null argument"); return -1; }
    if (cfg->C != NMA_C) { nma_set_error("nma_create: network_dims[0] must be %d (got %d)", NMA_C, cfg->C); return -1; }
    if (cfg->D != 1 && cfg->D != 2) { nma_set_error("nma_create: flow_dims must be 1 or 2"); return -1; }
    if (cfg->F < 1 || cfg->F > NMA_MAX_FLOWS) { nma_set_error("nma_create: no_flows out of range"); return -1; }
    if (cfg->H < 0 || cfg->H > NMA_MAXH) { nma_set_error("nma_create: at most %d hidden layers", NMA_MAXH); return -1; }
    if (cfg->p < 1 || cfg->K < 1 || cfg->B < 1) { nma_set_error("nma_create: p, K, B must be positive"); return -1; }
    if (cfg->dtheta < 1 || cfg->dtheta > 6) { nma_set_error("nma_create: dtheta must be in 1..6"); return -1; }
    const int cf_in = cfg->Cf + (cfg->feat_aug ? cfg->Cf - 2 : 0);
    if (cfg->Cf < 1 || cf_in > NMA_MAX_CHAN) { nma_set_error("nma_create: bad feature channel count"); return -1; }
    if (cfg->n_arrays < 1 || cfg->n_arrays > NMA_MAX_ARRAYS) { nma_set_error("nma_create: bad n_arrays"); return -1; }
    if (cfg->D == 2 && ((cfg->K & 1) != 0)) { nma_set_error("nma_create: flow_dims=2 needs an even kernel_len"); return -1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        nma_set_error("nma_create: no CUDA device (this library has no CPU path)");
        return -2;
    }
    nma_handle_s* h = new (std::nothrow) nma_handle_s;
    if (!h) { nma_set_error("nma_create: out of host memory"); return -1; }
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->L0 = cfg->F * cfg->K + cfg->D * cfg->B + cfg->D;      // AR.py:132; fitz_nag_NVP.py:182-183
    h->S = cfg->D * cfg->B;
    h->Cf_in = cf_in;
    h->feat_off = cfg->feat_aug ? 1 : 0;
    h->KP = (cfg->K + 9) / 10 * 10;
    for (int i = 0; i <= cfg->F; ++i) {
        FlowDims& d = h->fd[i];
        d.L = h->L0 - i * cfg->K;
        d.Lin = d.L - 1;
        d.N = d.L - cfg->K;
        d.LP = (d.Lin + 3) & ~3;
        d.NP = (d.N + 3) & ~3;
        if (i < cfg->F && d.N < 1) { delete h; nma_set_error("nma_create: window too short"); return -1; }
    }
    h->is_lv = (cfg->model == NMA_MODEL_LV || cfg->model == NMA_MODEL_LVR || cfg->model == NMA_MODEL_LVB) ? 1 : 0;
    h->LW = h->L0 - 1;
    h->LWP = (h->LW + 3) & ~3;
    h->conv_cin = h->is_lv ? 1 + h->LW : NMA_C1;
    for (int i = 0; i < cfg->F; ++i) h->feat_out[i] = h->is_lv ? h->fd[i].Lin : NMA_C;
    if (h->is_lv && cfg->D != 2) { delete h; nma_set_error("nma_create: the Lotka-Volterra model has flow_dims = 2"); return -1; }
    compute_layout(h);
    // tensor-core conv (nma_tc_conv.cu): 1-D flows whose operand tile fits in shared memory
    h->tc_ok = (cfg->D == 1 && cfg->K <= 190) ? 1 : 0;
    h->tc_nacc = (tc_conv_smem_floats(2, cfg->K) * 4 + 4096 <= 227 * 1024) ? 2 : 1;
    {
        const char* env = getenv("NMA_TC");
        h->use_tc = h->tc_ok && !(env && env[0] == '0');
        h->use_tc_persist = 1;
        const char* envf = getenv("NMA_TC_FEAT");
        h->use_tc_feat = h->use_tc && !(envf && envf[0] == '0');
        h->use_bf16 = 0;            // decided below, once the kernels that understand the format are known to run
        h->dgrad_wide = 1;
        // taps in pairs for even kernel_len (every reference script); odd kernel_len keeps the tap-by-tap kernels
        h->tap_pairs = (cfg->K % 2 == 0) ? 1 : 0;
    }
    NMA_CHECK_CUDA(cudaGetDevice(&h->dev));
    NMA_CHECK_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->dev));
    h->aux = nullptr; h->aux2 = nullptr; h->ev_fork = nullptr; h->ev_join = nullptr; h->ev_join2 = nullptr; h->aux_pending = 0;
    NMA_CHECK_CUDA(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
    NMA_CHECK_CUDA(cudaStreamCreateWithFlags(&h->aux2, cudaStreamNonBlocking));
    NMA_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    NMA_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    NMA_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming));

    // ---- carve the workspace arena ----
    const int64_t p = cfg->p;
    int64_t total = 0;
    auto reserve = [&](int64_t floats) { int64_t o = total; total += align_up(floats * 4, 256); return o; };
    struct Off { int64_t x, dx, a[5], h[NMA_MAXH + 1], s, dA, df, df3, tb, dtb, wpk, wdpk, tin_hi, tin_lo, dat_hi, dat_lo, wtc_f, wtc_d, wtc_feat; } off[NMA_MAX_FLOWS + 1];
    for (int i = 0; i <= cfg->F; ++i) {
        const FlowDims& d = h->fd[i];
        const int64_t XP = (d.L + 3) & ~3;
        off[i].x = reserve(p * XP);
        off[i].dx = reserve(p * XP);
        if (i == cfg->F) break;
        if (h->is_lv) {
            off[i].a[0] = reserve(p * cf_in * h->LWP);
            for (int l = 1; l < 4; ++l) off[i].a[l] = reserve(p * NMA_C * h->LWP);
            off[i].a[4] = reserve(p * h->LW * d.LP);            // [window position w][unit m]: conv channel 1 + w
            off[i].df3 = reserve(p * NMA_C * h->LWP);
        } else {
            off[i].a[0] = reserve(p * cf_in * d.LP);
            for (int l = 1; l < 5; ++l) off[i].a[l] = reserve(p * NMA_C * d.LP);
            off[i].df3 = 0;
        }
        for (int l = 0; l <= cfg->H; ++l) off[i].h[l] = reserve(p * NMA_C * d.NP);
        off[i].s = reserve(p * d.NP);
        off[i].dA = reserve(p * NMA_C * d.NP);
        off[i].df = reserve(p * (h->is_lv ? h->LW : NMA_C) * d.LP);
        off[i].tb = reserve(p * 3 * NMA_C);
        off[i].dtb = reserve(p * NMA_C);
        off[i].wpk = reserve((int64_t)h->conv_cin * 5 * h->KP * 12);
        off[i].wdpk = reserve((int64_t)NMA_C * 6 * h->KP * 12);
        if (h->tc_ok) {
            const int64_t Q = (p * d.Lin + 255) / 256 * 256 + 640 + cfg->K;
            h->ws[i].tin_Q = h->ws[i].dat_Q = Q;
            off[i].tin_hi = reserve((int64_t)TC_CCH * Q * 4);
            off[i].tin_lo = reserve((int64_t)TC_CCH * Q * 4);
            off[i].dat_hi = reserve((int64_t)TC_CCH * Q * 4);
            off[i].dat_lo = reserve((int64_t)TC_CCH * Q * 4);
            off[i].wtc_f = reserve((int64_t)cfg->K * TC_WSTAGE);
            off[i].wtc_d = reserve((int64_t)cfg->K * TC_WSTAGE);
            off[i].wtc_feat = reserve((int64_t)10 * TC_CCH * 128 * 4);
        }
    }
    // whole-iteration entry point (nma_step.cu)
    struct { int64_t eps, z0, theta, u, logq, gth, terms, relbo, flags, norm, counter; } so;
    so.eps = reserve(p * h->L0); so.z0 = reserve(p * 8); so.theta = reserve(p * 8); so.u = reserve(p * 8); so.logq = reserve(p);
    so.gth = reserve(p * 8); so.terms = reserve(p * 4); so.relbo = reserve(p); so.flags = reserve(p);
    so.norm = reserve(1024 + 64); so.counter = reserve(4);
    // channel split of the SIMT conv kernels: at most 2 x SMs + one item block's worth of CTAs are in a split launch
    const int64_t split_ctas = 2 * h->sm_count + CONV_SPLIT_MAX;
    const int64_t off_part = reserve(split_ctas * CONV_SPLIT_PART_FLOATS), off_ticket = reserve(split_ctas);
    h->arena_bytes = total;
    cudaError_t e = cudaMalloc(&h->arena, (size_t)total);
    if (e != cudaSuccess) {
        nma_set_error("nma_create: cudaMalloc(%lld bytes) failed: %s", (long long)total, cudaGetErrorString(e));
        delete h;
        return -2;
    }
    e = cudaMemset(h->arena, 0, (size_t)total);   // pad columns must stay finite (zero) forever
    if (e != cudaSuccess) { nma_set_error("nma_create: memset failed: %s", cudaGetErrorString(e)); cudaFree(h->arena); delete h; return -2; }
    char* base = (char*)h->arena;
    for (int i = 0; i <= cfg->F; ++i) {
        FlowWs& w = h->ws[i];
        w.x = (float*)(base + off[i].x);
        w.dx = (float*)(base + off[i].dx);
        if (i == cfg->F) break;
        for (int l = 0; l < 5; ++l) w.a[l] = (float*)(base + off[i].a[l]);
        for (int l = 0; l <= cfg->H; ++l) w.h[l] = (float*)(base + off[i].h[l]);
        w.s = (float*)(base + off[i].s);
        w.dA = (float*)(base + off[i].dA);
        w.df = (float*)(base + off[i].df);
        w.df3 = h->is_lv ? (float*)(base + off[i].df3) : nullptr;
        w.tb = (float*)(base + off[i].tb);
        w.dtb = (float*)(base + off[i].dtb);
        w.wpk = (float*)(base + off[i].wpk);
        w.wdpk = (float*)(base + off[i].wdpk);
        if (h->tc_ok) {
            w.tin_hi = (float*)(base + off[i].tin_hi); w.tin_lo = (float*)(base + off[i].tin_lo);
            w.dat_hi = (float*)(base + off[i].dat_hi); w.dat_lo = (float*)(base + off[i].dat_lo);
            w.wtc_f = (float*)(base + off[i].wtc_f); w.wtc_d = (float*)(base + off[i].wtc_d);
            w.wtc_feat = (float*)(base + off[i].wtc_feat);
        }
    }
    h->step.eps = (float*)(base + so.eps); h->step.z0 = (float*)(base + so.z0); h->step.theta = (float*)(base + so.theta);
    h->step.u = (float*)(base + so.u);
    h->step.logq_theta = (float*)(base + so.logq); h->step.g_theta = (float*)(base + so.gth);
    h->step.terms = (float*)(base + so.terms); h->step.row_elbo = (float*)(base + so.relbo);
    h->step.flags = (uint32_t*)(base + so.flags); h->step.norm = (float*)(base + so.norm);
    h->step.counter = (unsigned long long*)(base + so.counter);
    h->step.seed = 1;
    h->split_part = (float*)(base + off_part); h->split_ticket = (unsigned*)(base + off_ticket);
    {
        const char* envb = getenv("NMA_TC_BF16");
        h->bf16_ok = bf16_path_ok(h);
        // default where the kernels cover the model (AR-type): the 2-term bf16 split; NMA_TC_BF16=0 selects 3xTF32
        h->use_bf16 = (h->bf16_ok && !(envb && envb[0] == '0')) ? 1 : 0;
    }
    *out = h;
    return 0;
}

extern "C" int nma_destroy(nma_handle h) {
    if (!h) return 0;
    comm_release(h);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_join2) cudaEventDestroy(h->ev_join2);
    if (h->aux) cudaStreamDestroy(h->aux);
    if (h->aux2) cudaStreamDestroy(h->aux2);
    if (h->arena) cudaFree(h->arena);
    delete h;
    return 0;
}

// The bf16-split conv operands are produced by k_feat_fwd_tc / k_conv_fwd_tcp / k_epi_bwd_tc only: the format is
// available exactly when those kernels are the ones that run (AR-type model, one hidden layer, no batch-norm).
static int bf16_path_ok(const nma_handle_s* h) {
    return h->use_tc && h->use_tc_feat && conv_fwd_tcp_supported(h) && epi_bwd_tc_supported(h);
}

// bit 0: conv on tcgen05; bit 1: feature MLP and head backward on tcgen05 as well; bit 2: the conv GEMMs in the
// 2-term bf16 split (kind::f16) instead of 3xTF32 - needs bits 0 and 1 and a configuration the persistent kernels cover.
extern "C" int nma_set_tensor_cores(nma_handle h, int32_t on) {
    if (!h) { nma_set_error("null handle"); return -1; }
    if (on && !h->tc_ok) { nma_set_error("the tensor-core conv does not support this configuration (flow_dims=%d, kernel_len=%d)", h->cfg.D, h->cfg.K); return -1; }
    const int was_bf = h->use_bf16;
    h->use_tc = (on & 1) ? 1 : 0;
    h->use_tc_feat = ((on & 3) == 3) ? 1 : 0;
    h->bf16_ok = bf16_path_ok(h);
    if ((on & 4) && !h->bf16_ok) {
        h->use_bf16 = 0;
        nma_set_error("the bf16-split conv needs the tensor-core feature/head kernels and the persistent conv kernels "
                      "(AR-type model: flow_dims=1, one hidden layer, no batch-norm)");
        return -1;
    }
    h->use_bf16 = (on & 4) ? 1 : 0;
    if (h->use_bf16 != was_bf && h->tc_ok) {
        // the operand buffers change format: pad slots and the gaps between rows must read as zero in the new one
        for (int i = 0; i < h->cfg.F; ++i) {
            const size_t bytes = (size_t)TC_CCH * h->ws[i].tin_Q * 16;
            NMA_CHECK_CUDA(cudaMemset(h->ws[i].tin_hi, 0, bytes));
            NMA_CHECK_CUDA(cudaMemset(h->ws[i].tin_lo, 0, bytes));
            NMA_CHECK_CUDA(cudaMemset(h->ws[i].dat_hi, 0, bytes));
            NMA_CHECK_CUDA(cudaMemset(h->ws[i].dat_lo, 0, bytes));
        }
    }
    return 0;
}
extern "C" int nma_get_tensor_cores(nma_handle h) { return h ? (h->use_tc | (h->use_tc_feat << 1) | (h->use_bf16 << 2)) : -1; }

extern "C" int64_t nma_param_count(nma_handle h) { return h ? h->n_params : -1; }
extern "C" int64_t nma_workspace_bytes(nma_handle h) { return h ? h->arena_bytes : -1; }

// offsets_out[i*32 + s]: s = 0..3 featw, 4..7 featb, 8 convw, 9 convb, 10..12 thw, 13..15 thb,
// 16..19 hidw, 20..23 hidb, 24..27 gamma (or -1), 28 headw, 29 headb, 30 = -1, 31 = -1 (beta = gamma + 50)
extern "C" int nma_param_layout(nma_handle h, int64_t* o, int32_t n) {
    if (!h || !o || n < h->cfg.F * 32) { nma_set_error("nma_param_layout: buffer too small"); return -1; }
    for (int i = 0; i < h->cfg.F; ++i) {
        const FlowParamOff& po = h->po[i];
        int64_t* q = o + i * 32;
        for (int l = 0; l < 4; ++l) { q[l] = po.featw[l]; q[4 + l] = po.featb[l]; }
        q[8] = po.convw; q[9] = po.convb;
        for (int l = 0; l < 3; ++l) { q[10 + l] = po.thw[l]; q[13 + l] = po.thb[l]; }
        for (int l = 0; l < NMA_MAXH; ++l) { q[16 + l] = po.hidw[l]; q[20 + l] = po.hidb[l]; q[24 + l] = po.gam[l]; }
        q[28] = po.headw; q[29] = po.headb; q[30] = -1; q[31] = -1;
    }
    return 0;
}

extern "C" int nma_set_series(nma_handle h, const float* const* d_arrays, const int64_t* lengths, int32_t n) {
    if (!h || !d_arrays || !lengths) { nma_set_error("nma_set_series: null argument"); return -1; }
    if (n != h->cfg.n_arrays) { nma_set_error("nma_set_series: expected %d arrays, got %d", h->cfg.n_arrays, n); return -1; }
    for (int c = 0; c < h->cfg.Cf; ++c)
        if (h->cfg.chan_array[c] < 0 || h->cfg.chan_array[c] >= n) { nma_set_error("nma_set_series: channel table refers to a missing array"); return -1; }
    for (int i = 0; i < NMA_MAX_ARRAYS; ++i) { h->base[i] = nullptr; h->base_len[i] = 0; }
    for (int i = 0; i < n; ++i) {
        if (!d_arrays[i] || lengths[i] <= 0) { nma_set_error("nma_set_series: array %d is empty", i); return -1; }
        h->base[i] = d_arrays[i];
        h->base_len[i] = lengths[i];
    }
    return 0;
}

static int check_step_args(nma_handle h, int p, const void* a, const void* b, const void* c, const void* d) {
    if (!h) { nma_set_error("null handle"); return -1; }
    if (p < 1 || p > h->cfg.p) { nma_set_error("p=%d outside 1..%d (nma_create sized the workspace)", p, h->cfg.p); return -1; }
    if (!a || !b || !c || !d) { nma_set_error("null device pointer"); return -1; }
    if (!h->base[0]) { nma_set_error("nma_set_series has not been called"); return -1; }
    return 0;
}
static int check_model_built(nma_handle h) {
    if (h->cfg.model < NMA_MODEL_AR || h->cfg.model > NMA_MODEL_LVB) {
        nma_set_error("unknown model %d", h->cfg.model);
        return -3;
    }
    return 0;
}

extern "C" int nma_gather(nma_handle h, const int64_t* d_idx, int32_t p, float* d_tf, float* d_mask, float* d_shift,
                          void* stream) {
    if (check_step_args(h, p, d_idx, d_tf, d_tf, d_tf)) return -1;
    return launch_gather(h, d_idx, p, d_tf, d_mask, d_shift, (cudaStream_t)stream);
}

// Launches that do not fill the machine (the scripts' own row counts: under four 256-position tiles per SM) run their
// independent kernels side by side on the handle's second stream; at thousands of rows every kernel fills the machine
// by itself and the order of one stream is kept.
static bool step_is_small(const nma_handle_s* h, int p) {
    return (long long)p * h->fd[0].Lin < (long long)h->sm_count * 1024 && !getenv("NMA_NO_AUX_STREAM");
}
static int aux_fork(nma_handle_s* h, cudaStream_t st, cudaStream_t to = nullptr) {
    NMA_CHECK_CUDA(cudaEventRecord(h->ev_fork, st));
    NMA_CHECK_CUDA(cudaStreamWaitEvent(to ? to : h->aux, h->ev_fork, 0));
    return 0;
}
static int aux_join(nma_handle_s* h, cudaStream_t st) {
    NMA_CHECK_CUDA(cudaEventRecord(h->ev_join, h->aux));
    NMA_CHECK_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    return 0;
}

int step_aux_join(nma_handle_s* h, cudaStream_t st) {
    const int pending = h->aux_pending;      // bit 0: aux, bit 1: aux2 (a stream that was never forked must not be joined:
    h->aux_pending = 0;                      // inside a capture that would pull an uncaptured event into the graph)
    if (pending & 2) {
        NMA_CHECK_CUDA(cudaEventRecord(h->ev_join2, h->aux2));
        NMA_CHECK_CUDA(cudaStreamWaitEvent(st, h->ev_join2, 0));
    }
    if (pending & 1) return aux_join(h, st);
    return 0;
}

static int forward_all(nma_handle_s* h, const float* params, const float* eps, const float* theta, const int64_t* idx,
                       int p, bool save, cudaStream_t st) {
    int rc;
    const bool side = step_is_small(h, p);
    if (side) {
        // the conv tap kernels are packed next to theta-bias MLP + feature forward; the conv itself waits for both
        if ((rc = aux_fork(h, st))) return rc;
        if ((rc = launch_pack_weights(h, params, save, h->aux, 1))) return rc;
        if ((rc = launch_pack_weights(h, params, save, st, 2))) return rc;
    } else if ((rc = launch_pack_weights(h, params, save, st))) return rc;
    if ((rc = launch_theta_fwd(h, params, theta, p, st))) return rc;
    if ((rc = (h->is_lv ? launch_lv_feat_fwd(h, params, idx, eps, p, save, st)
                        : launch_feat_fwd_eps(h, params, idx, eps, p, save, st))))
        return rc;
    if (side && (rc = aux_join(h, st))) return rc;
    for (int i = 0; i < h->cfg.F; ++i)
        if ((rc = launch_conv_fwd(h, i, params, p, save, st))) return rc;
    return 0;
}

// section of the flat gradient that belongs to flow i (TF creation order: everything of flow i is contiguous)
static void flow_section(const nma_handle_s* h, int i, int64_t* off, int64_t* count) {
    *off = h->po[i].featw[0];
    *count = (i + 1 < h->cfg.F ? h->po[i + 1].featw[0] : h->n_params) - *off;
}

// forward + ELBO + backward on the handle's workspace.  With a communicator (nma_comm.cu) and per_flow_collective, the
// all-reduce of flow i's gradient section is issued on the side stream as soon as that flow's backward kernels are
// queued, so it runs under the backward pass of the earlier flows.
int step_forward_backward(nma_handle_s* h, const float* d_params, const float* d_eps, const float* d_theta,
                          const int64_t* d_idx, int p, int objective, float path_target, float* d_terms, float* d_lf,
                          float* d_grad_params, float* d_grad_theta, uint32_t* d_flags, bool per_flow_collective,
                          cudaStream_t st, bool defer_last_join) {
    int rc;
    NMA_CHECK_CUDA(cudaMemsetAsync(d_grad_params, 0, (size_t)h->n_params * 4, st));
    if ((rc = forward_all(h, d_params, d_eps, d_theta, d_idx, p, true, st))) return rc;
    if ((rc = launch_elbo(h, d_theta, d_eps, d_idx, p, objective, path_target, d_terms, d_lf, d_grad_theta, d_flags,
                          true, st)))
        return rc;
    const bool side = step_is_small(h, p);
    const bool collectives = per_flow_collective && h->comm.comm;
    for (int i = h->cfg.F - 1; i >= 0; --i) {
        if ((rc = launch_epi_bwd(h, i, d_params, p, objective, d_grad_params, st))) return rc;
        if (side && !collectives) {
            // Small launches (the scripts' own row counts), no collective waiting for the flow's section: only the chain
            // head backward -> data gradient -> head backward of the next flow is serial.  The conv weight gradient (needs
            // dA) and the feature backward (needs the data gradient's df) leave it for the handle's second stream; the
            // theta-bias MLP backward is short and stays, which keeps the d/dtheta accumulation in one stream.  For the
            // last flow the data gradient leaves as well: st goes on to the theta chain of the caller.  Everything on the
            // second stream is joined once - by nma_train_step right before the optimiser, else at the end of this call.
            if ((rc = aux_fork(h, st))) return rc;
            if (h->is_lv) {
                if ((rc = launch_lv_conv_wgrad(h, i, p, d_grad_params, h->aux))) return rc;
            } else if ((rc = (h->use_bf16 ? launch_conv_wgrad_bf(h, i, p, d_grad_params, h->aux)
                              : h->use_tc ? launch_conv_wgrad_tc(h, i, p, d_grad_params, h->aux)
                                          : launch_conv_wgrad(h, i, p, d_grad_params, h->aux))))
                return rc;
            // (FP32 SIMT models: a third stream for the feature branch - weight gradient and feature backward do not
            // depend on each other; measured: LV batch 1.77 -> 1.56 ms.  The tensor-core kernels each claim whole SMs, 210 KB of
            // shared memory per CTA, and three of them side by side were 3 % slower than two on the AR default shape)
            cudaStream_t fs = h->use_tc ? h->aux : h->aux2;
            cudaStream_t ds = (i == 0) ? fs : st;
            if (i == 0 && fs != h->aux && (rc = aux_fork(h, st, fs))) return rc;
            if (h->is_lv) {
                if ((rc = launch_lv_conv_dgrad(h, i, d_params, p, ds))) return rc;
            } else if ((rc = (h->use_tc ? launch_conv_dgrad_tc(h, i, p, ds) : launch_conv_dgrad(h, i, p, ds)))) return rc;
            if (i > 0 && (rc = aux_fork(h, st, fs))) return rc;       // the feature backward waits for this flow's df
            if (h->is_lv && (rc = launch_lv_feat4_bwd(h, i, d_params, p, d_grad_params, fs))) return rc;
            if ((rc = launch_feat_bwd(h, i, d_params, p, d_grad_params, fs))) return rc;
            if ((rc = launch_theta_bwd(h, d_params, d_theta, p, d_grad_params, d_grad_theta, i, st))) return rc;
            h->aux_pending |= (fs == h->aux2) ? 3 : 1;
            continue;
        }
        // the conv weight gradient needs dA only: at small row counts it runs on the second stream next to the data
        // gradient and the feature backward (which need each other), and is joined before the flow's section is used
        cudaStream_t ws = side ? h->aux : st;
        if (side && (rc = aux_fork(h, st))) return rc;
        if (h->is_lv) {
            if ((rc = launch_lv_conv_wgrad(h, i, p, d_grad_params, ws))) return rc;
            if ((rc = launch_lv_conv_dgrad(h, i, d_params, p, st))) return rc;
            if ((rc = launch_lv_feat4_bwd(h, i, d_params, p, d_grad_params, st))) return rc;
            if ((rc = launch_feat_bwd(h, i, d_params, p, d_grad_params, st))) return rc;
        } else {
            if (side) {
                if ((rc = (h->use_bf16 ? launch_conv_wgrad_bf(h, i, p, d_grad_params, ws)
                           : h->use_tc ? launch_conv_wgrad_tc(h, i, p, d_grad_params, ws)
                                       : launch_conv_wgrad(h, i, p, d_grad_params, ws))))
                    return rc;
            }
            if ((rc = (h->use_tc ? launch_conv_dgrad_tc(h, i, p, st) : launch_conv_dgrad(h, i, p, st)))) return rc;
            if (!side) {
                if ((rc = (h->use_bf16 ? launch_conv_wgrad_bf(h, i, p, d_grad_params, st)
                           : h->use_tc ? launch_conv_wgrad_tc(h, i, p, d_grad_params, st)
                                       : launch_conv_wgrad(h, i, p, d_grad_params, st))))
                    return rc;
            }
            if ((rc = launch_feat_bwd(h, i, d_params, p, d_grad_params, st))) return rc;
        }
        // theta-bias MLP of this flow (AR.py:63-68): its gradient completes the flow's section of the blob
        if ((rc = launch_theta_bwd(h, d_params, d_theta, p, d_grad_params, d_grad_theta, i, st))) return rc;
        if (side && (rc = aux_join(h, st))) return rc;
        if (per_flow_collective && h->comm.comm) {
            int64_t off, count;
            flow_section(h, i, &off, &count);
            if ((rc = comm_allreduce_after(h, d_grad_params + off, count, i, st))) return rc;
        }
    }
    if (!defer_last_join && (rc = step_aux_join(h, st))) return rc;
    return 0;
}

extern "C" int nma_elbo_fwd_bwd(nma_handle h, const float* d_params, const float* d_eps, const float* d_theta,
                                const int64_t* d_idx, int32_t p, int32_t objective, float path_target, float* d_terms,
                                float* d_lf, float* d_grad_params, float* d_grad_theta, uint32_t* d_flags,
                                void* stream) {
    if (check_step_args(h, p, d_params, d_params, d_theta, d_idx)) return -1;
    if (check_model_built(h)) return -3;
    if (!d_terms || !d_grad_params || !d_grad_theta) { nma_set_error("null output pointer"); return -1; }
    if (objective < 0 || objective > 2) { nma_set_error("unknown objective %d", objective); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    const bool draw = (d_eps == nullptr);      // eps == NULL: the library draws the base noise (Philox4x32-10, nma_step.cu)
    if (draw) {
        if ((rc = launch_philox_normal(h->step.eps, (int64_t)p * h->L0, h->step.seed, h->step.counter, 0, 0u, 0.f, 1.f, st)))
            return rc;
        d_eps = h->step.eps;
    }
    if ((rc = step_forward_backward(h, d_params, d_eps, d_theta, d_idx, p, objective, path_target, d_terms, d_lf,
                                    d_grad_params, d_grad_theta, d_flags, true, st)))
        return rc;
    if (draw && (rc = launch_counter_bump(h, st))) return rc;
    return 0;
}

// profiling hook: re-launch ONE stage of the last step on the workspace it left behind (bench.py times the
// dominant kernel in isolation with CUDA events; results of the re-launch are discarded by the caller).
// stage: 0 conv_fwd, 1 conv_dgrad, 2 conv_wgrad, 3 epi_bwd, 4 feat_fwd, 5 feat_bwd
extern "C" int nma_launch_stage(nma_handle h, int32_t stage, int32_t flow, const float* d_params, const float* d_eps,
                                const int64_t* d_idx, int32_t p, float* d_grad_params, void* stream) {
    if (!h || flow < 0 || flow >= h->cfg.F || p < 1 || p > h->cfg.p) { nma_set_error("nma_launch_stage: bad argument"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    switch (stage) {
        case 0: return launch_conv_fwd(h, flow, d_params, p, true, st);
        case 1: return h->use_tc ? launch_conv_dgrad_tc(h, flow, p, st) : launch_conv_dgrad(h, flow, p, st);
        case 2: return h->use_bf16 ? launch_conv_wgrad_bf(h, flow, p, d_grad_params, st)
                     : h->use_tc ? launch_conv_wgrad_tc(h, flow, p, d_grad_params, st)
                                 : launch_conv_wgrad(h, flow, p, d_grad_params, st);
        case 3: return launch_epi_bwd(h, flow, d_params, p, NMA_OBJ_ELBO, d_grad_params, st);
        case 4: return launch_feat_fwd_eps(h, d_params, d_idx, d_eps, p, true, st);
        case 5: return launch_feat_bwd(h, flow, d_params, p, d_grad_params, st);
        default: nma_set_error("nma_launch_stage: unknown stage %d", stage); return -1;
    }
}

static long long g_launches = 0;
void nma_count_launch(int n) { g_launches += n; }
extern "C" int64_t nma_launch_count(void) { return g_launches; }

extern "C" int nma_forward_paths(nma_handle h, const float* d_params, const float* d_eps, const float* d_theta,
                                 const int64_t* d_idx, int32_t p, float* d_terms, float* d_lf, void* stream) {
    if (check_step_args(h, p, d_params, d_params, d_theta, d_idx)) return -1;
    if (check_model_built(h)) return -3;
    if (!d_terms) { nma_set_error("null output pointer"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    const bool draw = (d_eps == nullptr);
    if (draw) {
        if ((rc = launch_philox_normal(h->step.eps, (int64_t)p * h->L0, h->step.seed, h->step.counter, 0, 0u, 0.f, 1.f, st)))
            return rc;
        d_eps = h->step.eps;
    }
    if ((rc = forward_all(h, d_params, d_eps, d_theta, d_idx, p, false, st))) return rc;
    if ((rc = launch_elbo(h, d_theta, d_eps, d_idx, p, NMA_OBJ_ELBO, 0.f, d_terms, d_lf, nullptr, nullptr, false, st)))
        return rc;
    if (draw && (rc = launch_counter_bump(h, st))) return rc;
    return 0;
}
