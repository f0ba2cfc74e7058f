// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) for the instruction shapes and
// operand sources the conv kernels use.  One CTA per SM, one issuing thread, operands are whatever shared / tensor
// memory holds (timing only).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../viforssms_b200/csrc
#include <cstdio>
#include <cstdlib>
#include "nma_tc.cuh"

void nma_set_error(const char*, ...) {}
void nma_count_launch(int) {}

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode bits: 0 = A from TMEM; 1 = MN-major smem operands; nacc = accumulator tiles cycled through; N = instruction N
template <int N, int NACC, bool TS, bool MN>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* out) {
    extern __shared__ __align__(128) uint4 sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < 8192; t += blockDim.x) sm[t] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, (!TS && MN) ? 1 : 0, MN ? 1 : 0);
        const uint32_t base = smem_u32(sm);
        // K-major: [k chunk][row][16 B]; MN-major: [mn chunk][k position][16 B]; strides chosen like the conv kernels'
        const uint64_t ad = desc_pack(desc_lo(base, MN ? 128u : 4096u), desc_hi(MN ? 1152u : 128u));
        const uint64_t bd = desc_pack(desc_lo(base + 65536u, MN ? 128u : 4096u), desc_hi(MN ? 1152u : 128u));
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t d = tmem + (uint32_t)(((i * 8 + u) % NACC) * N);
                    if (TS) mma_ts(d, tmem + 480u, bd, idesc, 1u);
                    else umma_bf16(d, ad, bd, idesc, 1u);
                }
            }
            tc_commit(&bar);
        }
        __syncwarp();
        mbar_wait_backoff(&bar, 0);
        t1 = clock64();
        if (elect_one() && blockIdx.x == 0) out[0] = (t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// round trip of tcgen05.commit -> mbarrier -> try_wait by the issuing warp, with `nmma` N=64 TS instructions before each
// commit; and of a plain mbarrier arrive by another warp (ping-pong between two warps)
__global__ void __launch_bounds__(128, 1) k_commit_latency(int iters, int nmma, int mode, long long* out) {
    extern __shared__ __align__(128) uint4 sm[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < 8192; t += blockDim.x) sm[t] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
    const uint64_t bd = desc_pack(desc_lo(smem_u32(sm) + 65536u, 128u), desc_hi(1152u));
    if (mode == 0 && warp == 1) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (elect_one()) {
                for (int u = 0; u < nmma; ++u) mma_ts(tmem + (uint32_t)((u & 3) * 64), tmem + 480u, bd, idesc, 1u);
                tc_commit(&bar);
            }
            __syncwarp();
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(i & 1)) : "memory");
            tc_fence_after();
        }
        if (elect_one() && blockIdx.x == 0) out[0] = clock64() - t0;
    }
    if (mode == 1 && (warp == 1 || warp == 2)) {          // ping-pong: warp 1 arrives on bar, warp 2 answers on bar2
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint64_t* mine = warp == 1 ? &bar2 : &bar;
            uint64_t* other = warp == 1 ? &bar : &bar2;
            if (warp == 1) { __syncwarp(); if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(other)) : "memory"); }
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(smem_u32(mine)), "r"((uint32_t)(i & 1)) : "memory");
            if (warp == 2) { __syncwarp(); if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(other)) : "memory"); }
        }
        if (warp == 1 && elect_one() && blockIdx.x == 0) out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// throughput of the tensor pipe with `ncommit` tcgen05.commit per group of `nmma` instructions, nobody waiting in between
__global__ void __launch_bounds__(128, 1) k_commit_tput(int iters, int nmma, int ncommit, long long* out) {
    extern __shared__ __align__(128) uint4 sm[];
    __shared__ uint64_t bars[8], fin;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < 8192; t += blockDim.x) sm[t] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1 << 20); mbar_init(&fin, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
    const uint64_t bd = desc_pack(desc_lo(smem_u32(sm) + 65536u, 128u), desc_hi(1152u));
    if (warp == 1) {
        const long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
                for (int u = 0; u < nmma; ++u) mma_ts(tmem + (uint32_t)((u & 3) * 64), tmem + 480u, bd, idesc, 1u);
                for (int c = 0; c < ncommit; ++c) tc_commit(&bars[c & 7]);
            }
            tc_commit(&fin);
        }
        __syncwarp();
        mbar_wait_backoff(&fin, 0);
        if (elect_one() && blockIdx.x == 0) out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// nw warps issue concurrently, each into its own accumulator tiles (N = 64, A from TMEM): cycles per instruction overall
__global__ void __launch_bounds__(192, 1) k_multi_issue(int iters, int nw, long long* out) {
    extern __shared__ __align__(128) uint4 sm[];
    __shared__ uint64_t fin;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < 8192; t += blockDim.x) sm[t] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&fin, nw); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
    const uint64_t bd = desc_pack(desc_lo(smem_u32(sm) + 65536u, 128u), desc_hi(1152u));
    const long long t0 = clock64();
    if (warp >= 1 && warp <= nw) {
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int u = 0; u < 8; ++u) mma_ts(tmem + (uint32_t)((warp - 1) * 64), tmem + 480u, bd, idesc, 1u);
            }
            tc_commit(&fin);
        }
        __syncwarp();
    }
    if (warp == 1) {
        mbar_wait_backoff(&fin, 0);
        if (elect_one() && blockIdx.x == 0) out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

void run_multi(int nw, long long* d_out) {
    const int iters = 1000;
    cudaFuncSetAttribute(k_multi_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_multi_issue<<<148, 192, 200 * 1024>>>(iters, nw, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%d warps issuing concurrently (N=64 TS, own tiles): %7.1f cycles per instruction overall %s\n", nw, (double)c / (iters * 8.0 * nw),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// the issue stream of the conv kernels: per instruction a fresh descriptor pair whose low words live in VECTOR registers
// (derived from a value the compiler cannot prove warp-uniform), as in p_mma / pp_mma of nma_tc_conv2.cu.
// mode 0: SS N=128 + SS N=64 alternating (forward kernel's pair of instructions); mode 1: the same with uniform descriptors
__global__ void __launch_bounds__(128, 1) k_issue_stream(int iters, int mode, const int* __restrict__ zero, long long* out) {
    extern __shared__ __align__(128) uint4 sm[];
    __shared__ uint64_t fin;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < 8192; t += blockDim.x) sm[t] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&fin, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    constexpr uint32_t id128 = umma_idesc_bf16(128, 128, 0, 0), id64 = umma_idesc_bf16(128, 64, 0, 0);
    if (warp == 1) {
        // per-thread (vector) base: zero[] is 0 at run time, unknown at compile time
        const uint32_t vbase = smem_u32(sm) + (uint32_t)zero[threadIdx.x & 31];
        const uint32_t ubase = smem_u32(sm);
        const uint32_t hi32 = desc_hi(128u);
        const long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
                const uint32_t row = (uint32_t)(2 * (i % 25));
#pragma unroll
                for (int a = 0; a < 2; ++a) {
#pragma unroll
                    for (int ks = 0; ks < 7; ++ks) {
                        const uint32_t base = mode == 0 ? vbase : ubase;
                        const uint32_t alo = desc_lo(base, 4096u) + row + (uint32_t)(a * 128) + (uint32_t)(ks % 3) * 512u;
                        const uint32_t blo = desc_lo(base + 65536u, 2048u) + (uint32_t)ks * 256u;
                        umma_bf16(tmem + (uint32_t)(a * 128), desc_pack(alo, hi32), desc_pack(blo, hi32), id128, 1u);
                        umma_bf16(tmem + (uint32_t)(a * 128 + 64), desc_pack(alo + 8192u, hi32), desc_pack(blo, hi32), id64, 1u);
                    }
                }
            }
            tc_commit(&fin);
        }
        __syncwarp();
        mbar_wait_backoff(&fin, 0);
        if (elect_one() && blockIdx.x == 0) out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

void run_stream(int mode, long long* d_out) {
    const int iters = 500;
    int* zero;
    cudaMalloc(&zero, 128);
    cudaMemset(zero, 0, 128);
    cudaFuncSetAttribute(k_issue_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_issue_stream<<<148, 128, 200 * 1024>>>(iters, mode, zero, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("forward-kernel issue stream (14 x [SS N=128, SS N=64] per stage, math 112 cycles per pair), descriptors from %s registers: %7.1f cycles per pair %s\n",
           mode == 0 ? "VECTOR" : "uniform", (double)c / (iters * 14.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(zero);
}

void run_tput(int nmma, int ncommit, long long* d_out) {
    const int iters = 1000;
    cudaFuncSetAttribute(k_commit_tput, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_commit_tput<<<148, 128, 200 * 1024>>>(iters, nmma, ncommit, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%2d MMAs (%4d cycles of math) + %d commits per group, nobody waiting: %7.1f cycles per group %s\n", nmma, nmma * 32, ncommit,
           (double)c / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

void run_commit(int nmma, int mode, long long* d_out) {
    const int iters = 2000;
    cudaFuncSetAttribute(k_commit_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_commit_latency<<<148, 128, 200 * 1024>>>(iters, nmma, mode, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    if (mode == 0) printf("commit round trip with %2d MMAs (N=64 TS, %4d cycles of math) : %7.1f cycles %s\n", nmma, nmma * 32, (double)c / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    else printf("mbarrier arrive -> try_wait ping-pong between two warps (2 hops)  : %7.1f cycles %s\n", (double)c / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N, int NACC, bool TS, bool MN>
void run(const char* name, long long* d_out) {
    const int iters = 2048;
    cudaFuncSetAttribute(k_rate<N, NACC, TS, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_rate<N, NACC, TS, MN><<<148, 128, 200 * 1024>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-34s N=%3d acc tiles=%d : %7.1f cycles / MMA   (math %d)%s\n", name, N, NACC, (double)c / (iters * 8.0), N / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    run<64, 1, false, false>("SS K-major", d_out);
    run<128, 1, false, false>("SS K-major", d_out);
    run<256, 1, false, false>("SS K-major", d_out);
    run<64, 4, false, false>("SS K-major", d_out);
    run<128, 2, false, false>("SS K-major", d_out);
    run<64, 1, false, true>("SS MN-major", d_out);
    run<128, 1, false, true>("SS MN-major", d_out);
    run<64, 4, false, true>("SS MN-major", d_out);
    run<64, 1, true, false>("TS, B K-major", d_out);
    run<128, 1, true, false>("TS, B K-major", d_out);
    run<256, 1, true, false>("TS, B K-major", d_out);
    run<64, 4, true, false>("TS, B K-major", d_out);
    run<64, 1, true, true>("TS, B MN-major", d_out);
    run<128, 1, true, true>("TS, B MN-major", d_out);
    run<64, 4, true, true>("TS, B MN-major", d_out);
    run<128, 2, true, true>("TS, B MN-major", d_out);
    run<256, 1, true, true>("TS, B MN-major", d_out);
    run_commit(0, 0, d_out);
    run_commit(1, 0, d_out);
    run_commit(8, 0, d_out);
    run_commit(32, 0, d_out);
    run_commit(0, 1, d_out);
    run_tput(32, 0, d_out);
    run_tput(32, 1, d_out);
    run_tput(32, 2, d_out);
    run_tput(32, 4, d_out);
    run_tput(0, 1, d_out);
    run_tput(8, 1, d_out);
    run_multi(1, d_out);
    run_multi(2, d_out);
    run_multi(4, d_out);
    run_stream(0, d_out);
    run_stream(1, d_out);
    return 0;
}
