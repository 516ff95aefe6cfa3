// Microbench 4: N-way all-to-all exchange, effect of slot layout (stride between producers' slots), number of
// replicas, poll start delay.  One warp per CTA does both publish and poll (loads do not wait on stores, see MB3).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ u64 ld_s(const u64* p) { u64 w; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory"); return w; }
__device__ __forceinline__ void st_s(u64* p, u64 w) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); }

// slot of (producer c, word w) inside a replica/parity block: c*stride + w   (stride >= W, in 8-byte words)
template <int MAXLD>
__global__ void xchg_kernel(u64* buf, int N, int W, int stride, int R, int iters, int delay, long long* out, long long* polls_out) {
    const int lane = threadIdx.x, cta = blockIdx.x;
    const size_t blk = (size_t)N * stride;             // words per (replica, parity)
    const u64* mine = buf + (size_t)(cta % R) * 2 * blk;
    const int total = N * W;
    const int nld = (total + 31) / 32;
    long long polls = 0;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        const long long ts = clock64();
        for (int idx = lane; idx < W * R; idx += 32) {
            const int w = idx % W, r = idx / W;
            st_s(buf + ((size_t)r * 2 + par) * blk + (size_t)cta * stride + w, (u64)it);
        }
        if (delay) while (clock64() - ts < delay) {}
        bool ok;
        do {
            ok = true;
            u64 v[MAXLD];
#pragma unroll
            for (int k = 0; k < MAXLD; ++k)
                if (k < nld) {
                    const int idx = k * 32 + lane;
                    const int c = idx / W, w = idx % W;
                    v[k] = (idx < total) ? ld_s(mine + par * blk + (size_t)c * stride + w) : (u64)it;
                }
#pragma unroll
            for (int k = 0; k < MAXLD; ++k)
                if (k < nld) ok = ok && (v[k] == (u64)it);
            ++polls;
            if (clock64() - ts > 200000000LL) { out[cta] = -1; return; }
        } while (!__all_sync(0xffffffffu, ok));
    }
    if (lane == 0) { out[cta] = clock64() - t0; polls_out[cta] = polls; }
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    CK(cudaSetDevice(0));
    u64* buf; CK(cudaMalloc(&buf, 64 << 20));
    long long* out; CK(cudaMallocManaged(&out, 8192));
    long long* polls; CK(cudaMallocManaged(&polls, 8192));
    const int iters = 2000;
    printf("N-way exchange: cycles/exchange (mean over CTAs) and poll rounds/exchange\n");
    for (int N : {2, 128})
        for (int W : {1, 2, 7})
            for (int stride : {0, 4, 16})
                for (int R : {1, 4})
                    for (int delay : {0, 300, 600}) {
                        int st = stride == 0 ? W : stride;
                        if (st < W) continue;
                        if (R > N) continue;
                        if (N == 2 && (R > 1)) continue;
                        CK(cudaMemset(buf, 0, (size_t)R * 2 * N * st * 8 + 4096));
                        int n = N, w = W, r = R, it = iters, d = delay;
                        void* a[] = {&buf, &n, &w, &st, &r, &it, &d, &out, &polls};
                        CK(cudaLaunchCooperativeKernel((void*)xchg_kernel<28>, dim3(N), dim3(32), a, 0, 0));
                        CK(cudaDeviceSynchronize());
                        double mean = 0, pm = 0;
                        for (int c = 0; c < N; ++c) { mean += (double)out[c]; pm += (double)polls[c]; }
                        printf("  N=%3d W=%d stride=%2d words R=%d delay=%3d : %6.0f cycles, %.2f polls\n", N, W, st, R, delay,
                               mean / N / iters, pm / N / iters);
                    }
    return 0;
}
