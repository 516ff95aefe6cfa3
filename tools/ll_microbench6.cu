// Microbench 6: tight N-way all-to-all with padded slots: producer c owns S words (S*8 bytes) per parity.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void ld2(const u64* p, u64& a, u64& b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ void st_s(u64* p, u64 w) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); }
__device__ __forceinline__ void st_s2(u64* p, u64 a, u64 b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }

// N producers, W words each (W even), slot stride S words, R replicas (CTA c polls replica c % R), V2: 16-byte stores
template <int N, int W, int S, int R, bool V2>
__global__ void __launch_bounds__(32, 1) pad_kernel(u64* buf, int iters, long long* out, long long* rounds) {
    constexpr int PAIRS = W / 2;                    // 16-byte loads per producer
    constexpr int NLD = N * PAIRS / 32;             // loads per lane
    constexpr size_t BLK = (size_t)N * S;           // words per (replica, parity)
    const int lane = threadIdx.x, cta = blockIdx.x;
    if (cta >= N) return;
    const u64* mine = buf + (size_t)(cta % R) * 2 * BLK;
    long long nr = 0;
    const long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        if (V2) {
            if (lane < PAIRS * R) {
                const int r = lane / PAIRS, pr = lane % PAIRS;
                st_s2(buf + ((size_t)r * 2 + par) * BLK + (size_t)cta * S + 2 * pr, (u64)it, (u64)it);
            }
        } else {
            if (lane < W * R) {
                const int r = lane / W, w = lane % W;
                st_s(buf + ((size_t)r * 2 + par) * BLK + (size_t)cta * S + w, (u64)it);
            }
        }
        const long long ts = clock64();
        bool ok;
        do {
            u64 a[NLD], b[NLD];
#pragma unroll
            for (int k = 0; k < NLD; ++k) {
                const int idx = k * 32 + lane;        // (producer, pair) index: consecutive lanes -> consecutive pairs
                const int c = idx / PAIRS, pr = idx % PAIRS;
                ld2(mine + par * BLK + (size_t)c * S + 2 * pr, a[k], b[k]);
            }
            ok = true;
#pragma unroll
            for (int k = 0; k < NLD; ++k) ok = ok && (a[k] == (u64)it) && (b[k] == (u64)it);
            ++nr;
            if (clock64() - ts > 200000000LL) { out[cta] = -1; return; }
        } while (!__all_sync(0xffffffffu, ok));
    }
    if (lane == 0) { out[cta] = clock64() - t0; rounds[cta] = nr; }
}

template <int N, int W, int S, int R, bool V2> void run(u64* buf, long long* out, long long* rounds, int nsm) {
    const int iters = 3000;
    CK(cudaMemset(buf, 0, 4 << 20));
    int it = iters;
    void* a[] = {&buf, &it, &out, &rounds};
    CK(cudaLaunchCooperativeKernel((void*)pad_kernel<N, W, S, R, V2>, dim3(nsm), dim3(32), a, 0, 0));
    CK(cudaDeviceSynchronize());
    double m = 0, r = 0;
    for (int c = 0; c < N; ++c) { m += (double)out[c]; r += (double)rounds[c]; }
    printf("  N=%3d W=%d slot=%3d B R=%d %s : %6.0f cycles/exchange, %.2f poll rounds\n", N, W, S * 8, R, V2 ? "st.v2" : "st.u64",
           m / N / iters, r / N / iters);
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    u64* buf; CK(cudaMalloc(&buf, 16 << 20));
    long long* out; CK(cudaMallocManaged(&out, 8192));
    long long* rounds; CK(cudaMallocManaged(&rounds, 8192));
    printf("[W=2: r / o exchange]\n");
    run<128, 2, 2, 1, false>(buf, out, rounds, nsm);
    run<128, 2, 2, 1, true>(buf, out, rounds, nsm);
    run<128, 2, 4, 1, true>(buf, out, rounds, nsm);
    run<128, 2, 16, 1, true>(buf, out, rounds, nsm);
    run<128, 2, 16, 4, true>(buf, out, rounds, nsm);
    run<128, 2, 16, 16, true>(buf, out, rounds, nsm);
    run<128, 2, 4, 16, true>(buf, out, rounds, nsm);
    run<128, 2, 2, 16, true>(buf, out, rounds, nsm);
    printf("[W=8 (7 used): h exchange]\n");
    run<128, 8, 8, 1, false>(buf, out, rounds, nsm);
    run<128, 8, 8, 1, true>(buf, out, rounds, nsm);
    run<128, 8, 16, 1, true>(buf, out, rounds, nsm);
    run<128, 8, 16, 4, true>(buf, out, rounds, nsm);
    run<128, 8, 16, 8, true>(buf, out, rounds, nsm);
    run<128, 8, 8, 8, true>(buf, out, rounds, nsm);
    printf("[scaling with N, W=2, 128 B slots, R=1]\n");
    run<32, 2, 16, 1, true>(buf, out, rounds, nsm);
    run<64, 2, 16, 1, true>(buf, out, rounds, nsm);
    return 0;
}
