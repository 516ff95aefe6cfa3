// Round-2 signalling microbenchmarks behind the cluster design of the B = 1 sample loop.
//   T1  L2 ping-pong from one SM to every other SM, several line addresses  -> die map of the GPU
//   T2  DSMEM ping-pong inside a cluster: push (remote store + local poll), pull (local store + remote poll),
//       st.async + mbarrier
//   T3  16-way (and 8-way) cluster all-gather of 16 LL words per CTA, push and pull, lock-step loop
//   T4  grid exchange between the same-rank CTAs of the 8 clusters (one 128-byte line each), for the whole
//       GPU, per die and per pair
//   T5  how many 16- / 8-CTA clusters are co-resident at 1 CTA per SM
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hop_microbench tools/hop_microbench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
typedef unsigned long long u64;
#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr long long TIMEOUT = 3000000000LL;   // ~1.5 s: every spin is bounded

__device__ __forceinline__ void st_ll(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_ll(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_ll2(const u64* p, u64& a, u64& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned smid() {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}
__device__ __forceinline__ unsigned mapa(unsigned local, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_dsmem(unsigned addr, u64 v) { asm volatile("st.volatile.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_dsmem(unsigned addr) {
    u64 v;
    asm volatile("ld.volatile.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void ld_dsmem2(unsigned addr, u64& a, u64& b) {
    asm volatile("ld.volatile.shared::cluster.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void st_smem(unsigned addr, u64 v) { asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_smem(unsigned addr) {
    u64 v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void ld_smem2(unsigned addr, u64& a, u64& b) {
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}

// ------------------------------------------------------------------------------------------------ T1
// CTA `base` plays ping-pong with every other CTA in turn, on n_addr different pairs of lines.
__global__ void l2_pingpong_kernel(u64* buf, int base, int iters, int n_addr, long long addr_stride_words, float* out,
                                   unsigned* smids, int* abort_flag) {
    const int cta = blockIdx.x, n = gridDim.x;
    if (threadIdx.x != 0) return;
    smids[cta] = smid();
    u64* turn = buf + n_addr * addr_stride_words + 4096;   // its own line, far from the ping-pong lines
    if (cta == base) {
        for (int k = 0; k < n; ++k) {
            if (k == base) continue;
            st_ll(turn, (u64)k + 1);
            for (int a = 0; a < n_addr; ++a) {
                u64* X = buf + a * addr_stride_words;
                u64* Y = X + 16;
                const u64 tb = (u64)(k * n_addr + a) * (iters + 1);
                long long t0 = 0;
                for (int it = 0; it <= iters; ++it) {   // iteration 0 = rendezvous, untimed
                    if (it == 1) t0 = clock64();
                    st_ll(X, tb + it + 1);
                    const long long ts = clock64();
                    while (ld_ll(Y) != tb + it + 1) {
                        if (clock64() - ts > TIMEOUT) { *abort_flag = 1; return; }
                    }
                }
                out[k * n_addr + a] = (float)(clock64() - t0) / iters / 2.f;
            }
        }
    } else {
        const int k = cta;
        {
            const long long ts = clock64();
            while (ld_ll(turn) != (u64)k + 1) {
                if (*(volatile int*)abort_flag || clock64() - ts > 8 * TIMEOUT) return;
                __nanosleep(4000);
            }
        }
        for (int a = 0; a < n_addr; ++a) {
            u64* X = buf + a * addr_stride_words;
            u64* Y = X + 16;
            const u64 tb = (u64)(k * n_addr + a) * (iters + 1);
            for (int it = 0; it <= iters; ++it) {
                const long long ts = clock64();
                while (ld_ll(X) != tb + it + 1) {
                    if (*(volatile int*)abort_flag || clock64() - ts > 4 * TIMEOUT) return;
                }
                st_ll(Y, tb + it + 1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ T2
// MODE 0: push (remote st + local poll)   1: pull (local st + remote poll)   2: st.async + mbarrier
template <int MODE>
__global__ void dsmem_pingpong_kernel(int peer, int iters, float* out) {
    __shared__ __align__(16) u64 box[4];
    __shared__ __align__(8) u64 mbar;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const unsigned box_l = (unsigned)__cvta_generic_to_shared(&box[0]);
    const unsigned mb_l = (unsigned)__cvta_generic_to_shared(&mbar);
    if (threadIdx.x == 0) {
        box[0] = 0; box[1] = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_l));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    if (threadIdx.x == 0 && (rank == 0 || rank == (unsigned)peer)) {
        const unsigned other = rank == 0 ? (unsigned)peer : 0u;
        const unsigned box_r = mapa(box_l, other), mb_r = mapa(mb_l, other);
        long long t0 = 0;
        unsigned phase = 0;
        bool fail = false;
        for (int it = 0; it <= iters && !fail; ++it) {
            if (it == 1) t0 = clock64();
            const u64 tag = (u64)it + 1;
            const long long ts = clock64();
            if (MODE == 0) {
                if (rank == 0) {
                    st_dsmem(box_r, tag);
                    while (ld_smem(box_l) != tag) if (clock64() - ts > TIMEOUT) { fail = true; break; }
                } else {
                    while (ld_smem(box_l) != tag) if (clock64() - ts > TIMEOUT) { fail = true; break; }
                    st_dsmem(box_r, tag);
                }
            } else if (MODE == 1) {
                if (rank == 0) {
                    st_smem(box_l, tag);
                    while (ld_dsmem(box_r) != tag) if (clock64() - ts > TIMEOUT) { fail = true; break; }
                } else {
                    while (ld_dsmem(box_r) != tag) if (clock64() - ts > TIMEOUT) { fail = true; break; }
                    st_smem(box_l, tag);
                }
            } else {
                auto send = [&]() {
                    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];" ::"r"(box_r), "l"(tag),
                                 "r"(mb_r)
                                 : "memory");
                };
                auto recv = [&]() {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb_l), "r"(8u) : "memory");
                    unsigned done = 0;
                    while (!done) {
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(done)
                                     : "r"(mb_l), "r"(phase)
                                     : "memory");
                        if (clock64() - ts > TIMEOUT) { fail = true; break; }
                    }
                    phase ^= 1;
                };
                if (rank == 0) { send(); recv(); } else { recv(); send(); }
            }
        }
        if (rank == 0) out[0] = fail ? -1.f : (float)(clock64() - t0) / iters / 2.f;
    }
    cluster.sync();
}

// ------------------------------------------------------------------------------------------------ T3
// lock-step all-gather of W = 16 LL words per CTA inside a cluster of S CTAs; one warp per CTA.
// MODE 0 push: 8 (S=16) remote stores per lane, poll the local 256-word inbox.  MODE 1 pull: 16 local stores, poll remotely.
template <int MODE, int S>
__global__ void dsmem_allgather_kernel(int iters, float* out, int* fails) {
    __shared__ __align__(16) u64 box[2][S * 16];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), lane = threadIdx.x;
    const unsigned box_l = (unsigned)__cvta_generic_to_shared(&box[0][0]);
    for (int i = lane; i < 2 * S * 16; i += 32) (&box[0][0])[i] = 0;
    cluster.sync();
    constexpr int NPAIR = S * 16 / 2 / 32;     // 16-byte pairs per lane per poll round
    unsigned rbase[NPAIR];                     // pull: remote address of pair j ; push: local address
    unsigned sdst[S * 16 / 32];                // push: remote destination of store j
#pragma unroll
    for (int j = 0; j < NPAIR; ++j) {
        const int p = j * 32 + lane, r = p / 8, w = 2 * (p % 8);
        // pull: producer r keeps its 16 words at its own box[par][0..15]; push: they arrive at box[par][16 r ..]
        rbase[j] = (MODE == 1) ? mapa(box_l + w * 8, r) : (box_l + (r * 16 + w) * 8);
    }
#pragma unroll
    for (int j = 0; j < S * 16 / 32; ++j) {
        const int idx = j * 32 + lane, dest = idx / 16, w = idx % 16;
        sdst[j] = mapa(box_l + (rank * 16 + w) * 8, dest);
    }
    bool fail = false;
    const long long t0 = clock64();
    for (int it = 1; it <= iters && !fail; ++it) {
        const unsigned par_off = (it & 1) * S * 16 * 8;
        const u64 word = ((u64)it << 32) | lane;
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < S * 16 / 32; ++j) st_dsmem(sdst[j] + par_off, word);
        } else {
            if (lane < 16) st_smem(box_l + par_off + lane * 8, word);
        }
        const long long ts = clock64();
        for (;;) {
            u64 a[NPAIR], b[NPAIR];
#pragma unroll
            for (int j = 0; j < NPAIR; ++j) {
                if (MODE == 0) ld_smem2(rbase[j] + par_off, a[j], b[j]);
                else ld_dsmem2(rbase[j] + par_off, a[j], b[j]);
            }
            bool ok = true;
#pragma unroll
            for (int j = 0; j < NPAIR; ++j) ok = ok && (unsigned)(a[j] >> 32) == (unsigned)it && (unsigned)(b[j] >> 32) == (unsigned)it;
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - ts > TIMEOUT) { fail = true; break; }
        }
    }
    if (lane == 0) {
        out[blockIdx.x] = (float)(clock64() - t0) / iters;
        if (fail) atomicAdd(fails, 1);
    }
    cluster.sync();
}

// ------------------------------------------------------------------------------------------------ T4
// 8 clusters x 16 CTAs.  CTA (k, rho) publishes one 128-byte line (16 LL words) and polls the lines of the CTAs
// (k', rho) for every cluster k' of its group.  grp[k] = group id of cluster k.
__global__ void grid_samerank_kernel(u64* buf, const int* grp, int iters, int first_delay, float* out, unsigned* smids, int* fails) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rho = cluster.block_rank(), k = blockIdx.x / 16, lane = threadIdx.x;
    if (lane == 0) smids[blockIdx.x] = smid();
    int member[8], n = 0;
    for (int j = 0; j < 8; ++j)
        if (grp[j] == grp[k]) member[n++] = j;
    // lane l polls pair (l % 8) of producer member[l / 8 + 4 * j]
    int prod[2];
    for (int j = 0; j < 2; ++j) {
        const int m = lane / 8 + 4 * j;
        prod[j] = -1;
        for (int q = 0; q < 8; ++q) if (q == m && m < n) prod[j] = member[q];
    }
    bool fail = false;
    cluster.sync();
    const long long t0 = clock64();
    for (int it = 1; it <= iters && !fail; ++it) {
        const int par = it & 1;
        if (lane < 16) st_ll(buf + ((par * 8 + k) * 16 + rho) * 16 + lane, ((u64)it << 32) | lane);
        const long long ts = clock64();
        if (first_delay) while (clock64() - ts < first_delay) {}
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (prod[j] >= 0) {
                    u64 a, b;
                    ld_ll2(buf + ((par * 8 + prod[j]) * 16 + rho) * 16 + 2 * (lane % 8), a, b);
                    ok = ok && (unsigned)(a >> 32) == (unsigned)it && (unsigned)(b >> 32) == (unsigned)it;
                }
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - ts > TIMEOUT) { fail = true; break; }
        }
    }
    if (lane == 0) {
        out[blockIdx.x] = (float)(clock64() - t0) / iters;
        if (fail) atomicAdd(fails, 1);
    }
    cluster.sync();
}

template <typename K, typename... Args>
static cudaError_t launch_cluster(K kernel, int grid, int block, int cs, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    if (cs > 8) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e != cudaSuccess) {
        static bool told = false;
        if (!told) { printf("  (cooperative + cluster launch refused: %s; launching without the cooperative attribute)\n", cudaGetErrorString(e)); told = true; }
        cudaGetLastError();
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kernel, args...);
    }
    return e;
}

__global__ void __launch_bounds__(288, 1) occupancy_probe_kernel(float* out) {
    extern __shared__ float dyn[];
    if (out && threadIdx.x == 0) out[blockIdx.x] = dyn[0];
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d MHz\n", prop.name, nsm, prop.clockRate / 1000);

    int* d_abort; int* d_fails;
    CK(cudaMalloc(&d_abort, 4)); CK(cudaMalloc(&d_fails, 4));
    std::vector<int> die_of_smid(256, -1);

    // ---------------------------------------------------------------- T5 first: what can be co-resident
    {
        printf("\n[T5] co-resident clusters at 1 CTA/SM (288 threads, 150 KB dynamic smem)\n");
        CK(cudaFuncSetAttribute(occupancy_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024));
        CK(cudaFuncSetAttribute(occupancy_probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        for (int cs : {2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 32); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = 150 * 1024;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int ncl = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, occupancy_probe_kernel, &cfg);
            printf("  cluster size %2d: max active clusters %d (%d CTAs)  %s\n", cs, ncl, ncl * cs, cudaGetErrorString(e));
        }
    }

    // ---------------------------------------------------------------- T1
    {
        const int iters = 60, n_addr = 6;
        const long long stride_words = (12 * 1024 + 256) / 8;
        u64* buf; float* out; unsigned* smids;
        CK(cudaMalloc(&buf, 1 << 20)); CK(cudaMemset(buf, 0, 1 << 20));
        CK(cudaMalloc(&out, sizeof(float) * nsm * n_addr)); CK(cudaMemset(out, 0, sizeof(float) * nsm * n_addr));
        CK(cudaMalloc(&smids, 4 * nsm));
        CK(cudaMemset(d_abort, 0, 4));
        int base = 0;
        int it_ = iters, na = n_addr; long long sw = stride_words;
        void* args[] = {&buf, &base, &it_, &na, &sw, &out, &smids, &d_abort};
        CK(cudaLaunchCooperativeKernel((void*)l2_pingpong_kernel, dim3(nsm), dim3(32), args, 0, 0));
        CK(cudaDeviceSynchronize());
        std::vector<float> h(nsm * n_addr); std::vector<unsigned> hs(nsm);
        CK(cudaMemcpy(h.data(), out, sizeof(float) * nsm * n_addr, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hs.data(), smids, 4 * nsm, cudaMemcpyDeviceToHost));
        int ab; CK(cudaMemcpy(&ab, d_abort, 4, cudaMemcpyDeviceToHost));
        printf("\n[T1] L2 ping-pong one-way cycles from CTA0 (sm %u) to every CTA, %d line addresses (abort=%d)\n", hs[0], n_addr, ab);
        std::vector<float> mean(nsm, 0.f);
        for (int k = 1; k < nsm; ++k) {
            float s = 0;
            for (int a = 0; a < n_addr; ++a) s += h[k * n_addr + a];
            mean[k] = s / n_addr;
        }
        std::vector<float> sorted(mean.begin() + 1, mean.end());
        std::sort(sorted.begin(), sorted.end());
        float best_gap = 0, thr = 0;
        for (size_t i = 1; i < sorted.size(); ++i)
            if (sorted[i] - sorted[i - 1] > best_gap) { best_gap = sorted[i] - sorted[i - 1]; thr = 0.5f * (sorted[i] + sorted[i - 1]); }
        int n_near = 0;
        die_of_smid[hs[0]] = 0;
        for (int k = 1; k < nsm; ++k) { die_of_smid[hs[k]] = mean[k] < thr ? 0 : 1; n_near += mean[k] < thr; }
        printf("  sorted means: min %.0f  p25 %.0f  median %.0f  p75 %.0f  max %.0f ; largest gap %.0f at %.0f -> %d near, %d far\n",
               sorted.front(), sorted[sorted.size() / 4], sorted[sorted.size() / 2], sorted[3 * sorted.size() / 4], sorted.back(),
               best_gap, thr, n_near, nsm - 1 - n_near);
        for (int k = 1; k < nsm; ++k) {
            printf("  cta %3d sm %3u die %d :", k, hs[k], die_of_smid[hs[k]]);
            for (int a = 0; a < n_addr; ++a) printf(" %5.0f", h[k * n_addr + a]);
            printf("\n");
        }
        // per-address spread: is there an address effect?
        for (int a = 0; a < n_addr; ++a) {
            double sn = 0, sf = 0; int cn = 0, cf = 0;
            for (int k = 1; k < nsm; ++k) {
                if (die_of_smid[hs[k]] == 0) { sn += h[k * n_addr + a]; ++cn; } else { sf += h[k * n_addr + a]; ++cf; }
            }
            printf("  address %d: near mean %.0f  far mean %.0f\n", a, cn ? sn / cn : 0., cf ? sf / cf : 0.);
        }
        cudaFree(buf); cudaFree(out); cudaFree(smids);
    }

    // ---------------------------------------------------------------- T2
    {
        printf("\n[T2] DSMEM ping-pong, one-way cycles (rank 0 <-> rank peer)\n");
        float* out; CK(cudaMalloc(&out, 64));
        for (int cs : {2, 8, 16}) {
            for (int peer : {1, cs - 1}) {
                float r[3] = {0, 0, 0};
                cudaError_t e0 = launch_cluster(dsmem_pingpong_kernel<0>, cs, 32, cs, 0, peer, 2000, out);
                CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&r[0], out, 4, cudaMemcpyDeviceToHost));
                cudaError_t e1 = launch_cluster(dsmem_pingpong_kernel<1>, cs, 32, cs, 0, peer, 2000, out);
                CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&r[1], out, 4, cudaMemcpyDeviceToHost));
                cudaError_t e2 = launch_cluster(dsmem_pingpong_kernel<2>, cs, 32, cs, 0, peer, 2000, out);
                CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&r[2], out, 4, cudaMemcpyDeviceToHost));
                printf("  cluster %2d peer %2d : push %6.0f   pull %6.0f   st.async+mbarrier %6.0f   (%s %s %s)\n", cs, peer, r[0], r[1],
                       r[2], cudaGetErrorString(e0), cudaGetErrorString(e1), cudaGetErrorString(e2));
                if (cs == 2) break;
            }
        }
        cudaFree(out);
    }

    // ---------------------------------------------------------------- T3
    {
        printf("\n[T3] cluster all-gather of 16 LL words per CTA, cycles per lock-step exchange (mean over CTAs)\n");
        float* out; CK(cudaMalloc(&out, sizeof(float) * 256));
        auto report = [&](const char* name, int n, cudaError_t e) {
            CK(cudaDeviceSynchronize());
            std::vector<float> h(n); int f;
            CK(cudaMemcpy(h.data(), out, sizeof(float) * n, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&f, d_fails, 4, cudaMemcpyDeviceToHost));
            double s = 0; float mx = 0;
            for (float v : h) { s += v; mx = std::max(mx, v); }
            printf("  %-28s : mean %6.0f  max %6.0f  fails %d (%s)\n", name, s / n, mx, f, cudaGetErrorString(e));
        };
        CK(cudaMemset(d_fails, 0, 4));
        report("S=16 push, 1 cluster", 16, launch_cluster(dsmem_allgather_kernel<0, 16>, 16, 32, 16, 0, 3000, out, d_fails));
        report("S=16 pull, 1 cluster", 16, launch_cluster(dsmem_allgather_kernel<1, 16>, 16, 32, 16, 0, 3000, out, d_fails));
        report("S=16 push, 8 clusters", 128, launch_cluster(dsmem_allgather_kernel<0, 16>, 128, 32, 16, 0, 3000, out, d_fails));
        report("S=16 pull, 8 clusters", 128, launch_cluster(dsmem_allgather_kernel<1, 16>, 128, 32, 16, 0, 3000, out, d_fails));
        report("S=8 push, 16 clusters", 128, launch_cluster(dsmem_allgather_kernel<0, 8>, 128, 32, 8, 0, 3000, out, d_fails));
        report("S=8 pull, 16 clusters", 128, launch_cluster(dsmem_allgather_kernel<1, 8>, 128, 32, 8, 0, 3000, out, d_fails));
        cudaFree(out);
    }

    // ---------------------------------------------------------------- T4
    {
        printf("\n[T4] grid exchange between same-rank CTAs of 8 clusters x 16 (one 128-byte line each)\n");
        u64* buf; float* out; unsigned* smids; int* grp;
        CK(cudaMalloc(&buf, 2 * 8 * 16 * 16 * 8)); CK(cudaMalloc(&out, sizeof(float) * 128)); CK(cudaMalloc(&smids, 4 * 128));
        CK(cudaMalloc(&grp, 4 * 8));
        std::vector<int> cl_die(8, 0);
        auto run = [&](const char* name, std::vector<int> g, int delay) {
            CK(cudaMemset(buf, 0, 2 * 8 * 16 * 16 * 8)); CK(cudaMemset(d_fails, 0, 4));
            CK(cudaMemcpy(grp, g.data(), 4 * 8, cudaMemcpyHostToDevice));
            cudaError_t e = launch_cluster(grid_samerank_kernel, 128, 32, 16, 0, buf, (const int*)grp, 3000, delay, out, smids, d_fails);
            CK(cudaDeviceSynchronize());
            std::vector<float> h(128); std::vector<unsigned> hs(128); int f;
            CK(cudaMemcpy(h.data(), out, sizeof(float) * 128, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hs.data(), smids, 4 * 128, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&f, d_fails, 4, cudaMemcpyDeviceToHost));
            printf("  %-44s delay %4d :", name, delay);
            for (int k = 0; k < 8; ++k) {
                double s = 0;
                for (int r = 0; r < 16; ++r) s += h[k * 16 + r];
                printf(" %5.0f", s / 16);
            }
            printf("  fails %d (%s)\n", f, cudaGetErrorString(e));
            return hs;
        };
        std::vector<unsigned> hs = run("all 8 clusters in one group", {0, 0, 0, 0, 0, 0, 0, 0}, 0);
        printf("  cluster -> die (from T1's map of the SMs): ");
        for (int k = 0; k < 8; ++k) {
            int d0 = 0, d1 = 0;
            for (int r = 0; r < 16; ++r) { const int d = die_of_smid[hs[k * 16 + r]]; d0 += d == 0; d1 += d == 1; }
            cl_die[k] = d1 > d0;
            printf(" c%d:%d(%d/%d)", k, cl_die[k], d0, d1);
        }
        printf("\n  smids of cluster 0: ");
        for (int r = 0; r < 16; ++r) printf(" %u", hs[r]);
        printf("\n");
        for (int delay : {0, 200, 400}) {
            run("all 8 clusters in one group", {0, 0, 0, 0, 0, 0, 0, 0}, delay);
            run("two groups by die", cl_die, delay);
            // pairs: same-die pairs and cross-die pairs
            std::vector<int> same(8), cross(8);
            int nd[2] = {0, 0};
            std::vector<int> idx_in_die(8);
            for (int k = 0; k < 8; ++k) idx_in_die[k] = nd[cl_die[k]]++;
            for (int k = 0; k < 8; ++k) { same[k] = cl_die[k] * 4 + idx_in_die[k] / 2; cross[k] = idx_in_die[k]; }
            run("pairs on the same die", same, delay);
            run("pairs across dies", cross, delay);
            std::vector<int> alone = {0, 1, 2, 3, 4, 5, 6, 7};
            run("no partner (publish + poll own line)", alone, delay);
        }
        cudaFree(buf); cudaFree(out); cudaFree(smids); cudaFree(grp);
    }
    printf("\ndone\n");
    return 0;
}
