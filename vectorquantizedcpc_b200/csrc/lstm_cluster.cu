// Latency LSTM of Encoder.encode (/root/reference/model.py:57,69) for a few utterances (BASELINE configs[0]: one 2 s utterance,
// 100 sequential steps): one 16-CTA cluster per utterance, h_t exchanged over DSMEM.
//
// The 32-CTA kernel (lstm_kernel<1>, encoder.cu) pays one L2 all-to-all per step: 1.26 us/step.  A DSMEM hop is ~330 cycles
// (tools/hop_microbench.cu), a cluster of 16 holds all of W_hh (1 MB fp32) in registers, and clusters need no co-residency
// with each other (utterances are independent), so the grid is simply B clusters.
//   CTA r owns hidden units 16 r .. 16 r + 15: 64 gate rows, row = 4 u + g.  Thread t: row t / 4, columns 64 (t % 4) .. +63
//   (64 weights in registers).  Per step: h_{t-1} from shared memory (16 LDS.128), 32 FFMA2 in four chains, two shuffle rounds
//   over the 4 lanes of a row, the four gates of a unit gathered to one lane (3 shuffles), the cell, and every owner lane sends
//   its h value to all 16 ranks with st.async (value + transaction bytes on the receiver's mbarrier in one instruction).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {
namespace cg = cooperative_groups;

constexpr int LC_S = 16, LC_H = 256, LC_G = 4 * LC_H, LC_U = LC_H / LC_S;      // 16 units per CTA
constexpr int LC_THREADS = 256;

struct LcParams {
    const float* table;     // (512, 1024) = W_ih e + b_ih + b_hh per code
    const int64_t* idx;     // (B, Tp)
    const float* w_hh;      // (1024, 256)
    float* out;             // (B, Tp, 256)
    int* status;
    int B, Tp;
    int debug;              // VQCPC_LC_DEBUG (timing ablations, wrong results): 1 no dot product, 2 no gate math, 4 no exchange
};

__device__ __forceinline__ float2 lc_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
        "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}

constexpr int LC_RING = 8;                       // steps of input-side gate terms in flight (cp.async ring)
constexpr int LC_IDX_CHUNK = 2048;               // code indices staged in shared memory at a time

__global__ void __cluster_dims__(LC_S, 1, 1) __launch_bounds__(LC_THREADS, 1) lstm_cluster_kernel(LcParams p) {
    __shared__ __align__(16) float h_s[2][LC_H];
    __shared__ __align__(16) float x_ring[LC_RING][64];          // [slot][row]: table[idx_t][g * 256 + unit] of this CTA's 64 rows
    __shared__ short idx_s[LC_IDX_CHUNK];
    __shared__ __align__(8) unsigned long long mbar[2];
    __shared__ int abort_flag;

    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31;
    const int rank = static_cast<int>(cluster.block_rank());
    const int b = blockIdx.x / LC_S;
    const int row = tid >> 2, chunk = tid & 3;                 // row = 4 u + g of this CTA's 64; columns 16 k + 4 chunk + {0..3}, k < 16
    const int u = row >> 2, g = row & 3;
    const int unit = rank * LC_U + u;

    // 64 weights per thread, in the order the h reads below deliver their operands: the four lanes of a row read four consecutive
    // 16-byte pieces of h per instruction (one shared-memory wavefront, no bank conflicts)
    float2 w[32];
    {
        const float* src = p.w_hh + static_cast<int64_t>(g * LC_H + unit) * LC_H + 4 * chunk;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + 16 * k));
            w[2 * k] = make_float2(v.x, v.y); w[2 * k + 1] = make_float2(v.z, v.w);
        }
    }
    pdl_sync();        // W_hh is not written by any kernel of the call: its load above runs under the predecessor (table GEMM)
    const unsigned mbar_a = static_cast<unsigned>(__cvta_generic_to_shared(&mbar[0]));
    const unsigned hs_a = static_cast<unsigned>(__cvta_generic_to_shared(&h_s[0][0]));
    const unsigned ring_a = static_cast<unsigned>(__cvta_generic_to_shared(&x_ring[0][row]));
    for (int i = tid; i < 2 * LC_H; i += LC_THREADS) (&h_s[0][0])[i] = 0.f;      // h_{-1} = 0 (parity 1 is read by step 0)
    if (tid == 0) {
        abort_flag = 0;
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_a + 8 * i), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();

    // the lane that finishes a unit: lane 0 / 16 of each warp (g == 0, chunk == 0)
    const bool owner = (lane & 15) == 0;
    unsigned dst[LC_S], dmb[LC_S];
    if (owner) {
#pragma unroll
        for (int d = 0; d < LC_S; ++d) {
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst[d]) : "r"(hs_a + unit * 4), "r"(d));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dmb[d]) : "r"(mbar_a), "r"(d));
        }
    }
    const int64_t* irow = p.idx + static_cast<int64_t>(b) * p.Tp;
    const float* tcol = p.table + g * LC_H + unit;             // this row's column of the (512, 1024) table
    // input-side term of step t -> ring slot t % LC_RING (one 4-byte cp.async per row, issued by the row's chunk-0 lane; every
    // thread commits a group per step so that the group arithmetic is uniform)
    auto fetch = [&](int t, int t_chunk0) {
        if (chunk == 0 && t < p.Tp) {
            const float* src = tcol + static_cast<int>(idx_s[t - t_chunk0]) * LC_G;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ring_a + (t % LC_RING) * 64 * 4), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float cst = 0.f;

    for (int t0 = 0; t0 < p.Tp; t0 += LC_IDX_CHUNK - LC_RING) {
        // stage (and clamp) the code indices of steps t0 .. t0 + LC_IDX_CHUNK - 1; steps of this pass: t0 .. t_end - 1, the ring runs
        // LC_RING - 1 steps ahead inside the staged window
        const int t_end = min(p.Tp, t0 + LC_IDX_CHUNK - LC_RING);
        __syncthreads();
        for (int i = tid; i < LC_IDX_CHUNK && t0 + i < p.Tp; i += LC_THREADS) {
            long long id = __ldg(irow + t0 + i);
            idx_s[i] = static_cast<short>(id < 0 ? 0 : (id > 511 ? 511 : id));
        }
        __syncthreads();
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // slots of the previous pass are all consumed or re-requested below
        for (int k = 0; k < LC_RING - 1; ++k) fetch(t0 + k, t0);

        for (int t = t0; t < t_end; ++t) {
            const int cur = t & 1, prev = cur ^ 1;             // step t reads h_s[prev] (h_{t-1}) and fills h_s[cur] everywhere
            if (tid == 0 && !(p.debug & 4)) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a + 8 * cur), "r"(LC_H * 4) : "memory");
            fetch(t + LC_RING - 1, t0);
            // ---- this row's dot product over its 64 columns
            const float4* hp = reinterpret_cast<const float4*>(&h_s[prev][4 * chunk]);
            float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (p.debug & 1) break;
                const float4 hv = hp[4 * k];
                acc[k & 3] = lc_ffma2(w[2 * k + 1], make_float2(hv.z, hv.w), lc_ffma2(w[2 * k], make_float2(hv.x, hv.y), acc[k & 3]));
            }
            float s = ((acc[0].x + acc[0].y) + (acc[1].x + acc[1].y)) + ((acc[2].x + acc[2].y) + (acc[3].x + acc[3].y));
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            asm volatile("cp.async.wait_group %0;" ::"n"(LC_RING - 1) : "memory");      // this step's input-side term has landed
            if (chunk == 0) s += x_ring[t % LC_RING][row];
            // gates f, g, o of the unit sit 4, 8, 12 lanes up
            const float sf = __shfl_down_sync(0xffffffffu, s, 4), sg = __shfl_down_sync(0xffffffffu, s, 8), so = __shfl_down_sync(0xffffffffu, s, 12);
            if (owner) {
                // PyTorch gate order i, f, g, o
                float hn;
                if (p.debug & 2) {
                    hn = 0.25f * (s + sf + sg + so);
                } else {
                    const float ig = sigmoid_fast(s);
                    const float fg = sigmoid_fast(sf);
                    const float gg = tanh_fast(sg);
                    const float og = sigmoid_fast(so);
                    cst = fg * cst + ig * gg;
                    hn = og * tanh_fast(cst);
                }
                p.out[(static_cast<int64_t>(b) * p.Tp + t) * LC_H + unit] = hn;
                if (t + 1 < p.Tp && !(p.debug & 4)) {
#pragma unroll
                    for (int d = 0; d < LC_S; ++d)
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
                                     ::"r"(dst[d] + cur * LC_H * 4), "f"(hn), "r"(dmb[d] + 8 * cur) : "memory");
                }
            }
            if (t + 1 < p.Tp && !(p.debug & 4)) {
                // ---- wait for all 256 values of h_t (16 ranks x 16 units x 4 bytes on this CTA's mbarrier)
                const unsigned parity = (static_cast<unsigned>(t) >> 1) & 1u;
                const long long tw = clock64();
                unsigned n = 0;
                for (;;) {
                    unsigned done;
                    asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                                 : "=r"(done) : "r"(mbar_a + 8 * cur), "r"(parity) : "memory");
                    if (done) break;
                    if (abort_flag || ((++n & 63u) == 0 && clock64() - tw > LL_TIMEOUT_CYCLES)) {
                        abort_flag = 1;
                        if (lane == 0) atomicExch(p.status, VQCPC_ERR_TIMEOUT);
                        break;
                    }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    cluster.sync();      // nobody's shared memory disappears while a peer may still store into it
}

int lstm_cluster_supported() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (cache[dev]) return cache[dev] == 1;
    int ok = 0;
    do {
        if (cudaFuncSetAttribute(lstm_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) break;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(LC_S); cfg.blockDim = dim3(LC_THREADS);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = LC_S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, lstm_cluster_kernel, &cfg) != cudaSuccess) break;
        ok = ncl >= 1;
    } while (0);
    cudaGetLastError();
    cache[dev] = ok ? 1 : 2;
    return ok;
}

int lstm_cluster_launch(const float* table, const int64_t* idx, const float* w_hh, int B, int Tp, float* out, int* status,
                        cudaStream_t stream) {
    static const int dbg = [] { const char* e = getenv("VQCPC_LC_DEBUG"); return e ? atoi(e) : 0; }();
    LcParams p{table, idx, w_hh, out, status, B, Tp, dbg};
    VQ_CUDA(launch_pdl(true, lstm_cluster_kernel, dim3(B * LC_S), dim3(LC_THREADS), 0, stream, p));
    count_launch(1);
    return VQCPC_OK;
}

}  // namespace vqcpc
