// Microbench 2: latency of load/store flavours for cross-SM signalling on B200 (diagnostic).
//   C  dependent-chain load latency (pointer chase in L2-resident buffer) per load flavour
//   D  ping-pong one-way latency for each (store flavour, load flavour) pair
//   F  publish cost: time for a warp to issue K strong stores to K distinct lines followed by a poll that
//      another CTA answers (measures whether outstanding strong stores throttle)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned long long u64;
#define SPIN_LIMIT 400000000LL
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int F> __device__ __forceinline__ u64 ld(const u64* p) {
    u64 w;
    if (F == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (F == 1) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (F == 2) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (F == 3) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (F == 4) asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if (F == 5) asm volatile("ld.relaxed.cta.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
template <int F> __device__ __forceinline__ void st(u64* p, u64 w) {
    if (F == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (F == 1) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (F == 2) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (F == 3) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    if (F == 4) { u64 old; asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(w) : "memory"); }
    if (F == 5) asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
const char* LDN[] = {"ld.relaxed.gpu", "ld.cg(weak)", "ld.volatile", "ld.acquire.gpu", "ld.ca(weak)", "ld.relaxed.cta"};
const char* STN[] = {"st.relaxed.gpu", "st.cg(weak)", "st.volatile", "st.release.gpu", "atom.exch", "red.max"};

template <int F> __global__ void chase_kernel(u64* buf, int iters, long long* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u64 idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = ld<F>(buf + idx);
    out[0] = clock64() - t0;
    out[1] = (long long)idx;
}

template <int SF, int LF> __global__ void pingpong_kernel(u64* buf, int peer, int iters, long long* out) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me != 0 && me != peer) return;
    u64* mine = buf + (me == 0 ? 0 : 64);
    u64* theirs = buf + (me == 0 ? 64 : 0);
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            st<SF>(theirs, (u64)i);
            while (ld<LF>(mine) != (u64)i) { if (clock64() - t0 > SPIN_LIMIT) { out[0] = -1; return; } }
        } else {
            while (ld<LF>(mine) != (u64)i) { if (clock64() - t0 > SPIN_LIMIT) return; }
            st<SF>(theirs, (u64)i);
        }
    }
    if (me == 0) out[0] = clock64() - t0;
}

// F: CTA 0 lane l stores to K lines (k*16 words apart) then waits for CTA peer's reply; CTA peer waits for ALL K
// lines then replies with one store.  Round trip minus the K=1 round trip = cost of K outstanding stores.
template <int SF, int LF> __global__ void fanout_kernel(u64* buf, int peer, int K, int iters, long long* out) {
    const int me = blockIdx.x, lane = threadIdx.x;
    if (me != 0 && me != peer) return;
    u64* fan = buf + 4096;          // K lines, 16 words (128 B) apart
    u64* reply = buf;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            for (int k = lane; k < K; k += 32) st<SF>(fan + k * 16, (u64)i);
            bool dead = false;
            if (lane == 0) while (ld<LF>(reply) != (u64)i) { if (clock64() - t0 > SPIN_LIMIT) { dead = true; break; } }
            if (__any_sync(0xffffffffu, dead)) { if (lane == 0) out[0] = -1; return; }
        } else {
            bool ok;
            do {
                ok = true;
                for (int k = lane; k < K; k += 32) ok = ok && (ld<LF>(fan + k * 16) == (u64)i);
                if (clock64() - t0 > SPIN_LIMIT) return;
            } while (!__all_sync(0xffffffffu, ok));
            if (lane == 0) st<SF>(reply, (u64)i);
        }
    }
    if (me == 0 && lane == 0) out[0] = clock64() - t0;
}

template <int SF, int LF> void run_pp(u64* buf, long long* out, int nsm) {
    const int iters = 2000;
    double s = 0; int n = 0;
    for (int peer : {1, 37, 74, 111, 147}) {
        if (peer >= nsm) continue;
        CK(cudaMemset(buf, 0, 1 << 20));
        int p = peer, it = iters;
        void* a[] = {&buf, &p, &it, &out};
        CK(cudaLaunchCooperativeKernel((void*)pingpong_kernel<SF, LF>, dim3(nsm), dim3(32), a, 0, 0));
        CK(cudaDeviceSynchronize());
        s += (double)out[0] / iters / 2.0; ++n;
    }
    printf("  %-16s + %-16s : one-way %6.0f cycles\n", STN[SF], LDN[LF], s / n);
}
template <int SF, int LF> void run_fan(u64* buf, long long* out, int nsm) {
    const int iters = 1000;
    printf("  %-16s + %-16s : round trip with K stores:", STN[SF], LDN[LF]);
    for (int K : {1, 4, 8, 32, 64, 224}) {
        CK(cudaMemset(buf, 0, 1 << 20));
        int p = 74, it = iters, k = K;
        void* a[] = {&buf, &p, &k, &it, &out};
        CK(cudaLaunchCooperativeKernel((void*)fanout_kernel<SF, LF>, dim3(nsm), dim3(32), a, 0, 0));
        CK(cudaDeviceSynchronize());
        printf("  K=%d: %.0f", K, (double)out[0] / iters);
    }
    printf("\n");
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    u64* buf; CK(cudaMalloc(&buf, 64 << 20));
    long long* out; CK(cudaMallocManaged(&out, 4096 * 8));
    printf("[C] dependent-load latency (cycles/load), 1 thread, L2-resident 1 MiB ring, stride 4 KiB+128\n");
    {
        const int n = 1 << 17;  // 1 MiB of u64
        u64* h = (u64*)malloc(n * 8);
        for (int i = 0; i < n; ++i) h[i] = 0;
        u64 cur = 0;
        for (int i = 0; i < 200; ++i) { u64 nxt = (cur + 528) % n; h[cur] = nxt; cur = nxt; }
        h[cur] = 0;
        CK(cudaMemcpy(buf, h, n * 8, cudaMemcpyHostToDevice));
        const int iters = 4000;
#define CH(F) { chase_kernel<F><<<1, 32>>>(buf, iters, out); CK(cudaDeviceSynchronize()); chase_kernel<F><<<1, 32>>>(buf, iters, out); CK(cudaDeviceSynchronize()); printf("  %-16s : %6.0f\n", LDN[F], (double)out[0] / iters); }
        CH(0) CH(1) CH(2) CH(3) CH(4) CH(5)
        free(h);
    }
    printf("[D] ping-pong one-way latency (mean over 5 peers)\n");
    run_pp<0, 0>(buf, out, nsm);
    run_pp<0, 1>(buf, out, nsm);
    run_pp<1, 1>(buf, out, nsm);
    run_pp<1, 0>(buf, out, nsm);
    run_pp<2, 2>(buf, out, nsm);
    run_pp<3, 3>(buf, out, nsm);
    run_pp<4, 0>(buf, out, nsm);
    run_pp<4, 1>(buf, out, nsm);
    run_pp<5, 1>(buf, out, nsm);
    printf("[F] fan-out: CTA0 stores K lines, CTA74 waits for all K then replies (cycles per round trip)\n");
    run_fan<0, 0>(buf, out, nsm);
    run_fan<1, 1>(buf, out, nsm);
    run_fan<0, 1>(buf, out, nsm);
    run_fan<4, 1>(buf, out, nsm);
    return 0;
}
