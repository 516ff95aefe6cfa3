// Microbench 3: does a strong load wait behind an earlier strong store (same thread / same SM)?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ u64 ld_s(const u64* p) { u64 w; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory"); return w; }
__device__ __forceinline__ void st_s(u64* p, u64 w) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); }
__device__ __forceinline__ void st_w(u64* p, u64 w) { asm volatile("st.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); }

// mode 0: loads only; 1: strong store to X then load; 2: weak store then load; 3: store by warp 1 of the same CTA;
// 4: strong store to X then load, X polled by another CTA (blockIdx 1) concurrently
__global__ void t_kernel(u64* ring, u64* X, int mode, int iters, long long* out, volatile int* stop) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 1) {                      // reader of X (mode 4 only)
        if (mode == 4 && threadIdx.x == 0) { u64 acc = 0; while (!*stop) acc += ld_s(X); out[8] = (long long)acc; }
        return;
    }
    if (blockIdx.x != 0 || lane != 0) return;
    if (warp == 1) {
        if (mode == 3) { u64 i = 0; while (!*stop) st_s(X, ++i); }
        return;
    }
    u64 idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (mode == 1 || mode == 4) st_s(X, (u64)i);
        if (mode == 2) st_w(X, (u64)i);
        idx = ld_s(ring + idx);
    }
    out[0] = clock64() - t0;
    out[1] = (long long)idx;
    *stop = 1;
    __threadfence();
}

// exchange N=2 with timestamps: CTA0 records, per iteration, cycles from store issue to (a) first poll return, (b) success
__global__ void x_kernel(u64* buf, int iters, long long* out, int delay) {
    if (threadIdx.x != 0 || blockIdx.x > 1) return;
    const int me = blockIdx.x;
    long long sum_first = 0, sum_succ = 0, polls = 0;
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        long long t0 = clock64();
        st_s(buf + par * 64 + me * 16, (u64)it);
        if (delay) { while (clock64() - t0 < delay) {} }
        u64 v = ld_s(buf + par * 64 + (1 - me) * 16);
        long long t1 = clock64();
        ++polls;
        while (v != (u64)it) { v = ld_s(buf + par * 64 + (1 - me) * 16); ++polls; if (clock64() - t0 > 100000000LL) return; }
        long long t2 = clock64();
        sum_first += t1 - t0; sum_succ += t2 - t0;
    }
    if (me == 0) { out[0] = sum_first; out[1] = sum_succ; out[2] = polls; }
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    CK(cudaSetDevice(0));
    u64* buf; CK(cudaMalloc(&buf, 16 << 20));
    long long* out; CK(cudaMallocManaged(&out, 4096));
    int* stop; CK(cudaMallocManaged(&stop, 4));
    const int n = 1 << 17;
    u64* h = (u64*)calloc(n, 8);
    u64 cur = 0;
    for (int i = 0; i < 200; ++i) { u64 nxt = (cur + 528) % n; h[cur] = nxt; cur = nxt; }
    h[cur] = 0;
    CK(cudaMemcpy(buf, h, n * 8, cudaMemcpyHostToDevice));
    u64* X = buf + (4 << 20) / 8;
    const char* names[] = {"loads only", "strong store + load (same thread)", "weak store + load (same thread)",
                           "load; other warp of the SM stores strong continuously", "strong store to a line another SM polls + load"};
    const int iters = 4000;
    printf("[T] dependent strong-load chain, cycles per iteration\n");
    for (int mode = 0; mode < 5; ++mode) {
        *stop = 0;
        void* a[] = {&buf, &X, &mode, (void*)&iters, &out, &stop};
        int it = iters; a[3] = &it;
        CK(cudaLaunchCooperativeKernel((void*)t_kernel, dim3(2), dim3(64), a, 0, 0));
        CK(cudaDeviceSynchronize());
        printf("  %-60s : %6.0f\n", names[mode], (double)out[0] / iters);
    }
    printf("[X] 2-CTA exchange (store own, poll peer): cycles from store issue to first poll return / to success; polls per exchange\n");
    for (int delay : {0, 200, 400, 600, 800}) {
        CK(cudaMemset(buf + (8 << 20) / 8, 0, 4096));
        u64* xb = buf + (8 << 20) / 8;
        int it = 2000, d = delay;
        void* a[] = {&xb, &it, &out, &d};
        CK(cudaLaunchCooperativeKernel((void*)x_kernel, dim3(2), dim3(32), a, 0, 0));
        CK(cudaDeviceSynchronize());
        printf("  delay %4d before first poll: first poll returns at %6.0f, success at %6.0f, %.2f polls\n", delay,
               (double)out[0] / it, (double)out[1] / it, (double)out[2] / it);
    }
    return 0;
}
