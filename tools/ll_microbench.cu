// Microbenchmarks for cross-SM signalling on B200 (diagnostic; results in profiles/ll_microbench_*.txt).
//   A  ping-pong between CTA 0 and CTA k through 8-byte LL words in L2 (one-way latency)
//   B  N-way all-to-all exchange (every CTA publishes W words into R replicated inboxes, polls all N*W words)
//   E  cluster exchange through distributed shared memory (st.shared::cluster + local polling)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ll_microbench tools/ll_microbench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;
typedef unsigned long long u64;

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e = (x);                                                              \
        if (e != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

__device__ __forceinline__ void st_ll(u64* p, uint32_t v, uint32_t tag) {
    u64 w = ((u64)tag << 32) | v;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ u64 ld_ll(const u64* p) {
    u64 w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ void ld_ll2(const u64* p, u64& a, u64& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

// ---------------------------------------------------------------- A: ping-pong
__global__ void pingpong_kernel(u64* buf, int peer, int iters, long long* out, unsigned* smids) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me == 0) smids[0] = smid();
    if (me == peer) smids[1] = smid();
    if (me != 0 && me != peer) return;
    u64* mine = buf + (me == 0 ? 0 : 16);      // separate 128B lines
    u64* theirs = buf + (me == 0 ? 16 : 0);
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            st_ll(theirs, 0, i);
            while ((uint32_t)(ld_ll(mine) >> 32) != (uint32_t)i) {}
        } else {
            while ((uint32_t)(ld_ll(mine) >> 32) != (uint32_t)i) {}
            st_ll(theirs, 0, i);
        }
    }
    if (me == 0) out[0] = clock64() - t0;
}

// ---------------------------------------------------------------- B: N-way all-to-all
// Each CTA: warp 0 only.  Per iteration: publish W words (lanes < W) into R replicas, poll N*W words of
// its replica (lane l takes words l, l+32, ...), record per-iteration cycles for CTA 0.
template <int MAXLD>
__global__ void alltoall_kernel(u64* buf, int N, int W, int R, int iters, int work, long long* out) {
    const int lane = threadIdx.x, cta = blockIdx.x;
    if (threadIdx.x >= 32) return;
    const int total = N * W;                      // words per replica per parity
    const u64* mine = buf + (size_t)(cta % R) * 2 * total;
    const int nld = (total + 31) / 32;
    long long t0 = clock64();
    float dummy = (float)cta;
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        // optional fake compute between exchanges
        for (int k = 0; k < work; ++k) dummy = __fmaf_rn(dummy, 1.0001f, 0.5f);
        for (int idx = lane; idx < W * R; idx += 32) {
            const int w = idx % W, r = idx / W;
            st_ll(buf + ((size_t)r * 2 + par) * total + cta * W + w, __float_as_uint(dummy), it);
        }
        bool ok;
        do {
            ok = true;
            u64 v[MAXLD];
#pragma unroll
            for (int k = 0; k < MAXLD; ++k)
                if (k < nld) {
                    const int idx = k * 32 + lane;
                    v[k] = (idx < total) ? ld_ll(mine + par * total + idx) : ((u64)it << 32);
                }
#pragma unroll
            for (int k = 0; k < MAXLD; ++k)
                if (k < nld) ok = ok && ((uint32_t)(v[k] >> 32) == (uint32_t)it);
        } while (!__all_sync(0xffffffffu, ok));
    }
    if (lane == 0) out[cta] = clock64() - t0;
    if (dummy == 12345.f) out[0] = 0;
}

// ---------------------------------------------------------------- E: cluster DSMEM exchange
// cluster of CS CTAs; each CTA's warp 0 lane l<W writes its W words into every peer's shared inbox, then
// polls its own inbox (local shared memory) for CS*W tags.
__global__ void cluster_kernel(int W, int iters, long long* out) {
    __shared__ u64 inbox[2][16 * 32];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = cluster.num_blocks(), rank = cluster.block_rank(), lane = threadIdx.x;
    for (int i = lane; i < 2 * 16 * 32; i += 32) (&inbox[0][0])[i] = 0;
    cluster.sync();
    const int total = cs * W;
    long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        for (int idx = lane; idx < W * cs; idx += 32) {
            const int w = idx % W, peer = idx / W;
            u64* dst = cluster.map_shared_rank(&inbox[par][rank * W + w], peer);
            u64 val = ((u64)it << 32) | (unsigned)rank;
            asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"l"(dst), "l"(val) : "memory");
        }
        bool ok;
        do {
            ok = true;
            for (int idx = lane; idx < total; idx += 32) {
                u64 v = *((volatile u64*)&inbox[par][idx]);
                ok = ok && ((uint32_t)(v >> 32) == (uint32_t)it);
            }
        } while (!__all_sync(0xffffffffu, ok));
    }
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
    cluster.sync();
}

// ---------------------------------------------------------------- B2: all-to-all, publisher warp != poller warp
// warp 0 polls (and "computes"), warp 1 publishes when released through named barrier 1.
template <int MAXLD>
__global__ void alltoall_split_kernel(u64* buf, int N, int W, int R, int iters, long long* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cta = blockIdx.x;
    const int total = N * W;
    const u64* mine = buf + (size_t)(cta % R) * 2 * total;
    const int nld = (total + 31) / 32;
    long long t0 = clock64();
    if (warp == 1) {
        for (int it = 1; it <= iters; ++it) {
            const int par = it & 1;
            asm volatile("bar.sync 1, 64;" ::: "memory");
            for (int idx = lane; idx < W * R; idx += 32) {
                const int w = idx % W, r = idx / W;
                st_ll(buf + ((size_t)r * 2 + par) * total + cta * W + w, 1u, it);
            }
        }
        return;
    }
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        asm volatile("bar.arrive 1, 64;" ::: "memory");
        bool ok;
        do {
            ok = true;
            u64 v[MAXLD];
#pragma unroll
            for (int k = 0; k < MAXLD; ++k)
                if (k < nld) {
                    const int idx = k * 32 + lane;
                    v[k] = (idx < total) ? ld_ll(mine + par * total + idx) : ((u64)it << 32);
                }
#pragma unroll
            for (int k = 0; k < MAXLD; ++k)
                if (k < nld) ok = ok && ((uint32_t)(v[k] >> 32) == (uint32_t)it);
        } while (!__all_sync(0xffffffffu, ok));
    }
    if (lane == 0) out[cta] = clock64() - t0;
}

// ---------------------------------------------------------------- E2: cluster exchange, st.async + mbarrier tx
// every CTA: lane l < W*cs sends one 8-byte word to peer (l / W) with st.async ... mbarrier::complete_tx; the
// receiver waits on its local mbarrier (expect_tx = cs*W*8 bytes per phase).
__global__ void cluster_async_kernel(int W, int iters, long long* out) {
    __shared__ __align__(8) u64 inbox[2][16 * 32];
    __shared__ __align__(8) u64 mbar[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = cluster.num_blocks(), rank = cluster.block_rank(), lane = threadIdx.x;
    const unsigned mb0 = (unsigned)__cvta_generic_to_shared(&mbar[0]);
    const unsigned mb1 = (unsigned)__cvta_generic_to_shared(&mbar[1]);
    const unsigned bytes = cs * W * 8;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // arm both phases
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb0), "r"(bytes) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb1), "r"(bytes) : "memory");
    }
    __syncwarp();
    cluster.sync();
    long long t0 = clock64();
    unsigned phase[2] = {0, 0};
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        const unsigned mb = par ? mb1 : mb0;
        for (int idx = lane; idx < W * cs; idx += 32) {
            const int w = idx % W, peer = idx / W;
            unsigned local_dst = (unsigned)__cvta_generic_to_shared(&inbox[par][rank * W + w]);
            unsigned remote_dst, remote_mb;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_dst) : "r"(local_dst), "r"(peer));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_mb) : "r"(mb), "r"(peer));
            u64 val = ((u64)it << 32) | (unsigned)rank;
            asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];" ::"r"(remote_dst),
                         "l"(val), "r"(remote_mb)
                         : "memory");
        }
        // wait for all cs*W words of this phase
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(mb), "r"(phase[par])
                : "memory");
        }
        phase[par] ^= 1;
        // consume (read one word) then re-arm this phase's barrier for its next use (two iterations later)
        volatile u64 sink = inbox[par][lane % (cs * W)];
        (void)sink;
        __syncwarp();
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    }
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
    cluster.sync();
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    printf("device %s, %d SMs, clock %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    u64* buf;
    const size_t BUF_WORDS = 8u << 20;
    CK(cudaMalloc(&buf, BUF_WORDS * 8));
    long long* out;
    CK(cudaMallocManaged(&out, 4096 * sizeof(long long)));
    unsigned* smids;
    CK(cudaMallocManaged(&smids, 16));
    const int nsm = prop.multiProcessorCount;

    // ---- A
    printf("\n[A] ping-pong CTA0 <-> CTA k, one-way latency in cycles (round trip / 2)\n");
    {
        const int iters = 2000;
        std::vector<double> lat;
        for (int peer = 1; peer < nsm; peer += 7) {
            CK(cudaMemset(buf, 0, 4096));
            void* args[] = {&buf, (void*)&peer, (void*)&iters, &out, &smids};
            int p = peer, it = iters;
            void* a2[] = {&buf, &p, &it, &out, &smids};
            (void)args;
            CK(cudaLaunchCooperativeKernel((void*)pingpong_kernel, dim3(nsm), dim3(32), a2, 0, 0));
            CK(cudaDeviceSynchronize());
            double one_way = (double)out[0] / iters / 2.0;
            lat.push_back(one_way);
            printf("  peer cta %3d (sm %3u <-> sm %3u): %.0f cycles one-way\n", peer, smids[0], smids[1], one_way);
        }
        std::sort(lat.begin(), lat.end());
        printf("  min %.0f  median %.0f  max %.0f\n", lat.front(), lat[lat.size() / 2], lat.back());
    }

    // ---- B
    printf("\n[B] N-way all-to-all LL exchange: cycles per exchange (mean over CTAs), W words per CTA, R replicas\n");
    {
        const int iters = 2000;
        int Ns[] = {2, 8, 16, 32, 64, 128};
        int Ws[] = {1, 2, 7};
        int Rs[] = {1, 4, 32, 128};
        for (int N : Ns)
            for (int W : Ws)
                for (int R : Rs) {
                    if (R > N) continue;
                    if ((size_t)R * 2 * N * W > BUF_WORDS) continue;
                    CK(cudaMemset(buf, 0, (size_t)R * 2 * N * W * 8));
                    int n = N, w = W, r = R, it = iters, work = 0;
                    void* a[] = {&buf, &n, &w, &r, &it, &work, &out};
                    CK(cudaLaunchCooperativeKernel((void*)alltoall_kernel<28>, dim3(N), dim3(32), a, 0, 0));
                    CK(cudaDeviceSynchronize());
                    double mean = 0;
                    for (int c = 0; c < N; ++c) mean += (double)out[c];
                    mean /= N * (double)iters;
                    printf("  N=%3d W=%d R=%3d : %7.0f cycles/exchange\n", N, W, R, mean);
                }
    }

    // ---- B2
    printf("\n[B2] all-to-all, separate publisher / poller warps: cycles per exchange\n");
    {
        const int iters = 2000;
        int Ns[] = {2, 32, 128};
        int Ws[] = {1, 2, 7};
        int Rs[] = {1, 4, 32};
        for (int N : Ns)
            for (int W : Ws)
                for (int R : Rs) {
                    if (R > N) continue;
                    CK(cudaMemset(buf, 0, (size_t)R * 2 * N * W * 8));
                    int n = N, w = W, r = R, it = iters;
                    void* a[] = {&buf, &n, &w, &r, &it, &out};
                    CK(cudaLaunchCooperativeKernel((void*)alltoall_split_kernel<28>, dim3(N), dim3(64), a, 0, 0));
                    CK(cudaDeviceSynchronize());
                    double mean = 0;
                    for (int c = 0; c < N; ++c) mean += (double)out[c];
                    mean /= N * (double)iters;
                    printf("  N=%3d W=%d R=%3d : %7.0f cycles/exchange\n", N, W, R, mean);
                }
    }
    // ---- E2
    printf("\n[E2] cluster exchange with st.async + mbarrier complete_tx: cycles per exchange\n");
    {
        const int iters = 2000;
        for (int cs : {2, 8, 16}) {
            for (int W : {1, 4, 32}) {
                if (cs * W > 512) continue;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs);
                cfg.blockDim = dim3(32);
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = cs;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                if (cs > 8) CK(cudaFuncSetAttribute(cluster_async_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                int w = W, it = iters;
                cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_async_kernel, w, it, out);
                if (e != cudaSuccess) {
                    printf("  cluster %2d W=%2d: launch failed (%s)\n", cs, W, cudaGetErrorString(e));
                    cudaGetLastError();
                    continue;
                }
                CK(cudaDeviceSynchronize());
                double mean = 0;
                for (int c = 0; c < cs; ++c) mean += (double)out[c];
                mean /= cs * (double)iters;
                printf("  cluster %2d W=%2d : %6.0f cycles/exchange\n", cs, W, mean);
            }
        }
    }

    // ---- E
    printf("\n[E] cluster DSMEM exchange (st.shared::cluster + local poll): cycles per exchange\n");
    {
        const int iters = 2000;
        for (int cs : {2, 4, 8, 16}) {
            for (int W : {1, 2, 7, 32}) {
                if (cs * W > 512) continue;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs);
                cfg.blockDim = dim3(32);
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = cs;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                if (cs > 8) CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                int w = W, it = iters;
                cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_kernel, w, it, out);
                if (e != cudaSuccess) {
                    printf("  cluster %2d W=%2d: launch failed (%s)\n", cs, W, cudaGetErrorString(e));
                    cudaGetLastError();
                    continue;
                }
                CK(cudaDeviceSynchronize());
                double mean = 0;
                for (int c = 0; c < cs; ++c) mean += (double)out[c];
                mean /= cs * (double)iters;
                printf("  cluster %2d W=%2d : %6.0f cycles/exchange\n", cs, W, mean);
            }
        }
        // how many 8-/16-CTA clusters can be co-resident with 1 CTA/SM (256 threads, 100 KB smem)?
        for (int cs : {8, 16}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 16);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = 0;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cs;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int ncl = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, cluster_kernel, &cfg);
            printf("  max active clusters of %d (256 thr): %d (%s)\n", cs, ncl, cudaGetErrorString(e));
        }
    }
    return 0;
}
