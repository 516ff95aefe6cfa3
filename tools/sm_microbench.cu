// SM-level instruction microbenchmarks behind the sample-loop design (B200, sm_100a): latency (dependent chain, one warp)
// and throughput (independent chains, 1..8 warps per CTA on one SM) of FFMA, FFMA2 (fma.rn.f32x2), SHFL, REDUX/CREDUX,
// MUFU.EX2/RCP, LDS, named barriers.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sm_microbench tools/sm_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
                 "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

constexpr int N = 256;   // instructions per chain set

// mode 0: FFMA dependent chain; 1: FFMA 8 independent chains; 2: FFMA2 dependent; 3: FFMA2 8 independent;
// 4: SHFL dependent; 5: SHFL 8 independent; 6: redux.max.f32 dependent; 7: redux.min.u32 dependent; 8: MUFU.EX2 dependent;
// 9: LDS dependent (pointer chase); 10: bar.sync of all warps; 11: FADD dependent; 12: FMNMX dependent; 13: ex2 8 independent
__global__ void bench(int mode, float* out, long long* cycles, float seed) {
    __shared__ int chase[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) chase[threadIdx.x] = (threadIdx.x + 1) & 31;
    __syncthreads();
    float a = seed + lane, b = 1.0001f, c[8];
    float2 a2 = make_float2(a, a + 1), b2 = make_float2(b, b), c2[8];
    for (int i = 0; i < 8; ++i) { c[i] = i; c2[i] = make_float2(i, i + 1); }
    unsigned u = lane;
    int idx = lane;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        t0 = clock64();
        switch (mode) {
        case 0: for (int i = 0; i < N; ++i) c[0] = ffma1(a, b, c[0]); break;
        case 1: for (int i = 0; i < N / 8; ++i) { _Pragma("unroll") for (int k = 0; k < 8; ++k) c[k] = ffma1(a, b, c[k]); } break;
        case 2: for (int i = 0; i < N; ++i) c2[0] = ffma2(a2, b2, c2[0]); break;
        case 3: for (int i = 0; i < N / 8; ++i) { _Pragma("unroll") for (int k = 0; k < 8; ++k) c2[k] = ffma2(a2, b2, c2[k]); } break;
        case 4: for (int i = 0; i < N; ++i) c[0] = __shfl_xor_sync(0xffffffffu, c[0], 1); break;
        case 5: for (int i = 0; i < N / 8; ++i) { _Pragma("unroll") for (int k = 0; k < 8; ++k) c[k] = __shfl_xor_sync(0xffffffffu, c[k], 1 + k); } break;
        case 6: for (int i = 0; i < N; ++i) { float m; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(c[0])); c[0] = m + lane; } break;
        case 7: for (int i = 0; i < N; ++i) { unsigned m; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(u)); u = m + lane; } break;
        case 8: for (int i = 0; i < N; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(c[0])); c[0] = y; } break;
        case 9: for (int i = 0; i < N; ++i) idx = *reinterpret_cast<volatile int*>(&chase[idx]); break;
        case 10: for (int i = 0; i < N; ++i) asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory"); break;
        case 11: for (int i = 0; i < N; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(c[0]) : "f"(b)); break;
        case 12: for (int i = 0; i < N; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(c[0]) : "f"(b)); break;
        case 13: for (int i = 0; i < N / 8; ++i) { _Pragma("unroll") for (int k = 0; k < 8; ++k) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(c[k])); c[k] = y; } } break;
        }
        float sink = 0.f;
        for (int k = 0; k < 8; ++k) sink += c[k] + c2[k].x + c2[k].y;
        sink += u + idx;
        asm volatile("" ::"f"(sink));
        t1 = clock64();
        if (sink == 12345.678f) out[threadIdx.x] = sink;
    }
    if (lane == 0) cycles[warp] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64 * 8);
    const char* names[] = {"FFMA dependent", "FFMA 8 chains", "FFMA2 dependent", "FFMA2 8 chains", "SHFL dependent", "SHFL 8 chains",
                           "REDUX.max.f32 dependent (+FADD)", "REDUX.min.u32 dependent (+IADD)", "MUFU.EX2 dependent", "LDS dependent",
                           "BAR.SYNC all warps", "FADD dependent", "FMNMX dependent", "MUFU.EX2 8 chains"};
    printf("cycles per instruction as seen by a warp (N = %d instructions per warp), warps per CTA = 1 / 2 / 4 / 8 (one CTA, one SM)\n", N);
    for (int mode = 0; mode < 14; ++mode) {
        printf("%-34s", names[mode]);
        for (int nw : {1, 2, 4, 8}) {
            bench<<<1, 32 * nw>>>(mode, out, cyc, 1.0f);
            long long h[8];
            cudaMemcpy(h, cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            printf("  %7.2f", double(mx) / N);
        }
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
