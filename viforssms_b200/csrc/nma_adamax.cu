// A9: tf.global_norm + tf.clip_by_global_norm + AdamaxOptimizer._apply_dense as two streaming kernels
// (AR.py:230-234; optimisers/adamax.py:42-58).  HBM-bound: reads w,g,m,v (+g once more for the norm),
// writes w,m,v = 32 B per parameter.
#include "nma_common.cuh"

#define AM_BLOCKS 592      // 4 x 148 SMs
#define AM_THREADS 256

// `head` leading elements (0..3) are handled one by one so that the float4 body starts on a 16-byte boundary: the four
// buffers may be views into the tail of a larger blob (the theta posterior's variables behind the NMA variables).
__global__ void __launch_bounds__(AM_THREADS) k_sumsq(const float* __restrict__ g, int64_t n, int head,
                                                      float* __restrict__ part) {
    float acc = 0.f;
    const int64_t nb = n - head, n4 = nb >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g + head);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < head) acc = fmaf(g[threadIdx.x], g[threadIdx.x], acc);
        for (int64_t i = head + (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) acc = fmaf(g[i], g[i], acc);
    }
    __shared__ float red[AM_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < AM_THREADS / 32 ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) part[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(AM_THREADS) k_adamax(float* __restrict__ w, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, int64_t n, int head,
                                                       float lr,
                                                       float b1, float b2, float eps, float clip,
                                                       const float* __restrict__ part, int nparts,
                                                       float* __restrict__ norm_out) {
    // every block re-reduces the (<= 592) partial sums in a fixed order: deterministic and cheaper than a launch
    __shared__ float red[AM_THREADS / 32];
    __shared__ float s_scale;
    float acc = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += part[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < AM_THREADS / 32; ++i) t += red[i];
        const float norm = sqrtf(t);
        // tf.clip_by_global_norm: g * clip_norm / max(global_norm, clip_norm); clip <= 0 disables clipping
        s_scale = (clip > 0.f) ? clip / fmaxf(norm, clip) : 1.f;
        if (blockIdx.x == 0 && norm_out) norm_out[0] = norm;
    }
    __syncthreads();
    const float sc = s_scale, omb1 = 1.f - b1;
    const int64_t n4 = (n - head) >> 2;
    float4* w4 = reinterpret_cast<float4*>(w + head);
    const float4* g4 = reinterpret_cast<const float4*>(g + head);
    float4* m4 = reinterpret_cast<float4*>(m + head);
    float4* v4 = reinterpret_cast<float4*>(v + head);
    auto upd = [&](float& wi, float gi, float& mi, float& vi) {
        gi *= sc;
        vi = b1 * vi + omb1 * gi;                    // adamax.py:52
        mi = fmaxf(b2 * mi + eps, fabsf(gi));        // adamax.py:54
        wi -= lr * (vi / mi);                        // adamax.py:55,57
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 ww = w4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldg(g4 + i);
        upd(ww.x, gg.x, mm.x, vv.x); upd(ww.y, gg.y, mm.y, vv.y); upd(ww.z, gg.z, mm.z, vv.z); upd(ww.w, gg.w, mm.w, vv.w);
        w4[i] = ww; m4[i] = mm; v4[i] = vv;
    }
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < head) upd(w[threadIdx.x], g[threadIdx.x], m[threadIdx.x], v[threadIdx.x]);
        for (int64_t i = head + (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) upd(w[i], g[i], m[i], v[i]);
    }
}

extern "C" int nma_adamax_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n, float lr,
                               float beta1, float beta2, float eps, float clip, float* d_norm_out, float* d_scratch,
                               void* stream) {
    if (!d_params || !d_grads || !d_m || !d_v || !d_scratch || n < 1) { nma_set_error("nma_adamax_step: bad argument"); return -1; }
    // float4 body: the four buffers must share their offset within a 16-byte line (views at the same element offset of
    // same-shaped tensors always do); the 0..3 elements in front of the first boundary are handled one by one
    const uintptr_t mis = (uintptr_t)d_params & 15;
    if ((((uintptr_t)d_grads & 15) != mis) || (((uintptr_t)d_m & 15) != mis) || (((uintptr_t)d_v & 15) != mis) || (mis & 3)) {
        nma_set_error("nma_adamax_step: params, grads, m and v must share their alignment within 16 bytes (fp32 elements)");
        return -1;
    }
    int head = (int)(((16 - mis) & 15) >> 2);
    if (head > n) head = (int)n;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t want = (n / 4 + AM_THREADS - 1) / AM_THREADS;
    int blocks = (int)(want < 1 ? 1 : (want > AM_BLOCKS ? AM_BLOCKS : want));
    k_sumsq<<<blocks, AM_THREADS, 0, st>>>(d_grads, n, head, d_scratch);
    nma_count_launch(1);
    k_adamax<<<blocks, AM_THREADS, 0, st>>>(d_params, d_grads, d_m, d_v, n, head, lr, beta1, beta2, eps, clip, d_scratch, blocks,
                                            d_norm_out);
    nma_count_launch(1);
    NMA_CHECK_CUDA(cudaGetLastError());
    return 0;
}
