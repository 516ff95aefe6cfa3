// Microbench 7: ping-pong one-way latency when the poller READS with an L2 atomic (atom.add 0 / atom.or 0)
// instead of a strong load, and 128-way exchange with atomic polling.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ u64 ld_s(const u64* p) { u64 w; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory"); return w; }
__device__ __forceinline__ u64 ld_atom(u64* p) { u64 w; asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], 0;" : "=l"(w) : "l"(p) : "memory"); return w; }
__device__ __forceinline__ void st_s(u64* p, u64 w) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); }
__device__ __forceinline__ void st_atom(u64* p, u64 w) { u64 o; asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(w) : "memory"); }

template <int LD, int ST> __global__ void pp(u64* buf, int peer, int iters, long long* out) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me != 0 && me != peer) return;
    u64* mine = buf + (me == 0 ? 0 : 64);
    u64* theirs = buf + (me == 0 ? 64 : 0);
    const long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            if (ST) st_atom(theirs, (u64)i); else st_s(theirs, (u64)i);
            while ((LD ? ld_atom(mine) : ld_s(mine)) != (u64)i) { if (clock64() - t0 > 400000000LL) { out[0] = -1; return; } }
        } else {
            while ((LD ? ld_atom(mine) : ld_s(mine)) != (u64)i) { if (clock64() - t0 > 400000000LL) return; }
            if (ST) st_atom(theirs, (u64)i); else st_s(theirs, (u64)i);
        }
    }
    if (me == 0) out[0] = clock64() - t0;
}

// 128-way exchange, one word per CTA in its own 128 B slot; lane polls 4 slots with atomics or loads
template <int LD> __global__ void __launch_bounds__(32, 1) xchg(u64* buf, int iters, long long* out, long long* rounds) {
    const int lane = threadIdx.x, cta = blockIdx.x;
    if (cta >= 128) return;
    long long nr = 0;
    const long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        if (lane == 0) st_s(buf + (par * 128 + cta) * 16, (u64)it);
        const long long ts = clock64();
        bool ok;
        do {
            u64 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = LD ? ld_atom(buf + (par * 128 + 32 * k + lane) * 16) : ld_s(buf + (par * 128 + 32 * k + lane) * 16);
            ok = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) ok = ok && (v[k] == (u64)it);
            ++nr;
            if (clock64() - ts > 200000000LL) { out[cta] = -1; return; }
        } while (!__all_sync(0xffffffffu, ok));
    }
    if (lane == 0) { out[cta] = clock64() - t0; rounds[cta] = nr; }
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    u64* buf; CK(cudaMalloc(&buf, 16 << 20));
    long long* out; CK(cudaMallocManaged(&out, 8192));
    long long* rounds; CK(cudaMallocManaged(&rounds, 8192));
    const int iters = 2000;
    const char* names[2][2] = {{"st + ld.relaxed.gpu", "atom.exch + ld.relaxed.gpu"}, {"st + atom.add(0) read", "atom.exch + atom.add(0) read"}};
#define RUN_PP(LD, ST) { double s = 0; int n = 0; for (int peer : {1, 37, 74, 111, 147}) { if (peer >= nsm) continue; CK(cudaMemset(buf, 0, 4096)); int p = peer, it = iters; void* a[] = {&buf, &p, &it, &out}; CK(cudaLaunchCooperativeKernel((void*)pp<LD, ST>, dim3(nsm), dim3(32), a, 0, 0)); CK(cudaDeviceSynchronize()); s += (double)out[0] / iters / 2; ++n; } printf("  ping-pong %-30s: one-way %6.0f cycles\n", names[LD][ST], s / n); }
    RUN_PP(0, 0) RUN_PP(0, 1) RUN_PP(1, 0) RUN_PP(1, 1)
#define RUN_X(LD) { CK(cudaMemset(buf, 0, 1 << 20)); int it = 3000; void* a[] = {&buf, &it, &out, &rounds}; CK(cudaLaunchCooperativeKernel((void*)xchg<LD>, dim3(nsm), dim3(32), a, 0, 0)); CK(cudaDeviceSynchronize()); double m = 0, r = 0; for (int c = 0; c < 128; ++c) { m += (double)out[c]; r += (double)rounds[c]; } printf("  128-way exchange, poll with %-12s: %6.0f cycles/exchange, %.2f rounds\n", LD ? "atom.add(0)" : "ld.relaxed", m / 128 / 3000, r / 128 / 3000); }
    RUN_X(0) RUN_X(1)
    return 0;
}
