// fp32 SIMT GEMM  C[m,n] = sum_k A(m,k) * W[n,k] (+ bias[n])  -- the exact-arithmetic ("fp32 parity") path
// behind nn.Conv1d / nn.Linear of /root/reference/model.py:43,50,54 and the vocoder's input projections.
// A is supplied by a loader functor: dense row-major, or the implicit im2col of the stride-2 conv.
//
// Tiling: BM x BN x 16 per CTA, 256 threads as a 16x16 grid, each thread a (4*RM) x (4*RN) micro-tile read
// from k-major shared tiles with LDS.128; global tiles are register-prefetched one k-step ahead (double
// buffered shared memory, one __syncthreads per k-step).
//
// Split-K (latency case, M <= 256: one 2 s utterance is M = 100, which gives 24 CTAs a 768-deep serial k loop each: 47 us per
// layer): the gridDim.z CTAs that share an output tile form a thread-block CLUSTER (sgemm_splitk_cluster_kernel): each issues
// all global loads of its K / gridDim.z slice at once (one load latency instead of one per k-step), sums the slice, parks its
// 64 x 64 partial tile in its own shared memory, and after one cluster barrier every CTA adds a sixteenth-to-fifth of the tile over
// DSMEM in fixed order s = 0, 1, ... and writes C -- deterministic, no scratch planes in global memory, no fence, no atomics
// (the first version -- scratch planes + per-tile counter + last-arriving CTA reduces -- took 17.8 us per 100 x 768 x 768 layer).
// The order of the fp32 sum differs from the unsplit kernel's (error ~1e-7 relative; the parity bar is 1e-4).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {
namespace cg = cooperative_groups;

struct DenseA {
    const float* A;
    int64_t lda;
    __device__ __forceinline__ float4 load4(int64_t m, int k) const {
        return __ldg(reinterpret_cast<const float4*>(A + m * lda + k));
    }
};

// A(m, kk) with m = b*Tp + t, kk = i*4 + k  ->  mel[(b*Cin + i)*T + 2t + k - 1], zero outside [0, T)
// (nn.Conv1d(80, C, kernel 4, stride 2, padding 1), /root/reference/model.py:43).
struct ConvA {
    const float* mel;
    int T, Tp, Cin;
    __device__ __forceinline__ float4 load4(int64_t m, int kk) const {
        const int b = static_cast<int>(m / Tp);
        const int t = static_cast<int>(m - static_cast<int64_t>(b) * Tp);
        const int i = kk >> 2;
        const float* p = mel + (static_cast<int64_t>(b) * Cin + i) * T + 2 * t - 1;
        float4 v;
        v.x = (t > 0) ? __ldg(p) : 0.0f;
        v.y = __ldg(p + 1);
        v.z = (2 * t + 1 < T) ? __ldg(p + 2) : 0.0f;
        v.w = (2 * t + 2 < T) ? __ldg(p + 3) : 0.0f;
        return v;
    }
};

constexpr int GEMM_BK = 16;

template <int RM, int RN, class ALoader>
__global__ void __launch_bounds__(256)
sgemm_tn_kernel(ALoader a, const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias,
                float* __restrict__ C, int64_t ldc, int64_t M, int N, int K, float* __restrict__ scratch, unsigned* __restrict__ counters) {
    constexpr int BM = 64 * RM, BN = 64 * RN;
    constexpr int LDA = BM + 4, LDB = BN + 4;
    __shared__ __align__(16) float As[2][GEMM_BK][LDA];
    __shared__ __align__(16) float Bs[2][GEMM_BK][LDB];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM;
    const int n0 = blockIdx.y * BN;
    pdl_sync();

    float4 ra[RM], rb[RN];
    float acc[RM][4][RN][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int p = 0; p < RN; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][j][p][q] = 0.0f;

    auto gload = [&](int k0) {
#pragma unroll
        for (int j = 0; j < RM; ++j) {
            const int f = tid + j * 256, row = f >> 2, k4 = f & 3;
            const int64_t m = m0 + row;
            ra[j] = (m < M) ? a.load4(m, k0 + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int f = tid + j * 256, row = f >> 2, k4 = f & 3;
            const int n = n0 + row;
            rb[j] = (n < N) ? __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(n) * ldw + k0 + k4 * 4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < RM; ++j) {
            const int f = tid + j * 256, row = f >> 2, k4 = f & 3;
            As[buf][k4 * 4 + 0][row] = ra[j].x;
            As[buf][k4 * 4 + 1][row] = ra[j].y;
            As[buf][k4 * 4 + 2][row] = ra[j].z;
            As[buf][k4 * 4 + 3][row] = ra[j].w;
        }
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int f = tid + j * 256, row = f >> 2, k4 = f & 3;
            Bs[buf][k4 * 4 + 0][row] = rb[j].x;
            Bs[buf][k4 * 4 + 1][row] = rb[j].y;
            Bs[buf][k4 * 4 + 2][row] = rb[j].z;
            Bs[buf][k4 * 4 + 3][row] = rb[j].w;
        }
    };

    const int ksplit = gridDim.z;
    const int nk = K / GEMM_BK / ksplit;
    const int kbase = blockIdx.z * nk * GEMM_BK;
    gload(kbase);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload(kbase + (kt + 1) * GEMM_BK);
#pragma unroll
        for (int k = 0; k < GEMM_BK; ++k) {
            float4 av[RM], bv[RN];
#pragma unroll
            for (int i = 0; i < RM; ++i) av[i] = *reinterpret_cast<const float4*>(&As[buf][k][i * 64 + ty * 4]);
#pragma unroll
            for (int p = 0; p < RN; ++p) bv[p] = *reinterpret_cast<const float4*>(&Bs[buf][k][p * 64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                const float af[4] = {av[i].x, av[i].y, av[i].z, av[i].w};
#pragma unroll
                for (int p = 0; p < RN; ++p) {
                    const float bf[4] = {bv[p].x, bv[p].y, bv[p].z, bv[p].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][j][p][q] = fmaf(af[j], bf[q], acc[i][j][p][q]);
                }
            }
        }
        if (kt + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }

    if (ksplit > 1) {
        // this split's partial sums -> scratch plane blockIdx.z (dense M x N)
        __shared__ int is_last;
        float* plane = scratch + static_cast<int64_t>(blockIdx.z) * M * N;
#pragma unroll
        for (int p = 0; p < RN; ++p) {
            const int n = n0 + p * 64 + tx * 4;
            if (n >= N) continue;
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t m = m0 + i * 64 + ty * 4 + j;
                    if (m < M) *reinterpret_cast<float4*>(plane + m * N + n) = make_float4(acc[i][j][p][0], acc[i][j][p][1], acc[i][j][p][2], acc[i][j][p][3]);
                }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            unsigned* ctr = counters + blockIdx.y * gridDim.x + blockIdx.x;
            const unsigned old = atomicAdd(ctr, 1u);
            is_last = (old == static_cast<unsigned>(ksplit - 1));
            if (is_last) *ctr = 0u;                    // ready for the next GEMM that uses these counters (stream order)
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
#pragma unroll
        for (int p = 0; p < RN; ++p) {
            const int n = n0 + p * 64 + tx * 4;
            if (n >= N) continue;
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t m = m0 + i * 64 + ty * 4 + j;
                    if (m >= M) continue;
                    float4 v = __ldcg(reinterpret_cast<const float4*>(scratch + m * N + n));
                    for (int sidx = 1; sidx < ksplit; ++sidx) {
                        const float4 w = __ldcg(reinterpret_cast<const float4*>(scratch + (static_cast<int64_t>(sidx) * M + m) * N + n));
                        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                    }
                    acc[i][j][p][0] = v.x; acc[i][j][p][1] = v.y; acc[i][j][p][2] = v.z; acc[i][j][p][3] = v.w;
                }
        }
    }

#pragma unroll
    for (int p = 0; p < RN; ++p) {
        const int n = n0 + p * 64 + tx * 4;
        if (n >= N) continue;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(bias + n));
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t m = m0 + i * 64 + ty * 4 + j;
                if (m >= M) continue;
                float4 o;
                o.x = acc[i][j][p][0] + bb.x;
                o.y = acc[i][j][p][1] + bb.y;
                o.z = acc[i][j][p][2] + bb.z;
                o.w = acc[i][j][p][3] + bb.w;
                *reinterpret_cast<float4*>(C + m * ldc + n) = o;
            }
    }
}

// Split-K over a cluster (see the header): grid (M tiles, N tiles, S), cluster (1, 1, S), 64 x 64 tiles, K / S <= 16 * SK_NKMAX.
constexpr int SK_NKMAX = 6;
template <class ALoader>
__global__ void __launch_bounds__(256)
sgemm_splitk_cluster_kernel(ALoader a, const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias,
                            float* __restrict__ C, int64_t ldc, int64_t M, int N, int K) {
    constexpr int LD = 68;
    __shared__ __align__(16) float smem[4 * GEMM_BK * LD];             // As[2][16][68] | Bs[2][16][68]; afterwards red[64][64]
    float (*As)[GEMM_BK][LD] = reinterpret_cast<float (*)[GEMM_BK][LD]>(smem);
    float (*Bs)[GEMM_BK][LD] = reinterpret_cast<float (*)[GEMM_BK][LD]>(smem + 2 * GEMM_BK * LD);
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * 64;
    const int n0 = blockIdx.y * 64;
    const int S = static_cast<int>(cluster.num_blocks()), z = static_cast<int>(cluster.block_rank());
    const int nk = K / GEMM_BK / S;
    const int kbase = z * nk * GEMM_BK;
    pdl_sync();
    // every global load of this CTA's K slice is issued here: one load latency for the whole slice
    const int lrow = tid >> 2, k4 = tid & 3;
    const int64_t lm = m0 + lrow;
    const int ln = n0 + lrow;
    float4 ra[SK_NKMAX], rb[SK_NKMAX];
#pragma unroll
    for (int kt = 0; kt < SK_NKMAX; ++kt) {
        ra[kt] = make_float4(0.f, 0.f, 0.f, 0.f); rb[kt] = ra[kt];
        if (kt < nk) {
            const int k0 = kbase + kt * GEMM_BK + k4 * 4;
            if (lm < M) ra[kt] = a.load4(lm, k0);
            if (ln < N) rb[kt] = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(ln) * ldw + k0));
        }
    }
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = 0.0f;
#pragma unroll
    for (int kt = 0; kt < SK_NKMAX; ++kt) {
        if (kt < nk) {                                 // uniform over the CTA
            const int buf = kt & 1;
            As[buf][k4 * 4 + 0][lrow] = ra[kt].x; As[buf][k4 * 4 + 1][lrow] = ra[kt].y;
            As[buf][k4 * 4 + 2][lrow] = ra[kt].z; As[buf][k4 * 4 + 3][lrow] = ra[kt].w;
            Bs[buf][k4 * 4 + 0][lrow] = rb[kt].x; Bs[buf][k4 * 4 + 1][lrow] = rb[kt].y;
            Bs[buf][k4 * 4 + 2][lrow] = rb[kt].z; Bs[buf][k4 * 4 + 3][lrow] = rb[kt].w;
            __syncthreads();                           // one barrier per step: step kt + 1 writes the buffer step kt - 1 read
#pragma unroll
            for (int k = 0; k < GEMM_BK; ++k) {
                const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                const float af[4] = {av.x, av.y, av.z, av.w}, bf[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[j][q] = fmaf(af[j], bf[q], acc[j][q]);
            }
        }
    }
    __syncthreads();                                   // everyone is done with As / Bs: the tile of partial sums goes there
    float* red = smem;                                 // [64][64]
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(red + (ty * 4 + j) * 64 + tx * 4) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    cluster.sync();
    // CTA z adds the 16-byte pieces f = z, z + S, ... of the tile over all S partial tiles, in the order s = 0, 1, ...
    for (int f = z + S * tid; f < 64 * 16; f += S * 256) {
        const int row = f >> 4, c4 = f & 15;
        float4 v = *reinterpret_cast<const float4*>(cluster.map_shared_rank(red, 0) + row * 64 + c4 * 4);
        for (int sidx = 1; sidx < S; ++sidx) {
            const float4 w = *reinterpret_cast<const float4*>(cluster.map_shared_rank(red, sidx) + row * 64 + c4 * 4);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        const int64_t m = m0 + row;
        const int n = n0 + c4 * 4;
        if (m < M && n < N) {
            if (bias != nullptr) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            *reinterpret_cast<float4*>(C + m * ldc + n) = v;
        }
    }
    cluster.sync();                                    // no CTA's shared memory goes away while a peer still reads it
}

// scratch for split-K: SPLITK_MAX_S planes of M x N floats (M <= SPLITK_MAX_M) + SPLITK_COUNTERS zeroed per-tile counters (they
// reset themselves after use; the encoder keeps them in the reserved words of its workspace header, which every call clears)
constexpr int SPLITK_MAX_M = 256, SPLITK_MAX_S = 8;
size_t gemm_splitk_ws_bytes(int64_t M, int N) {
    if (M > SPLITK_MAX_M) return 0;
    return align_up(sizeof(float) * SPLITK_MAX_S * static_cast<size_t>(M) * N, 256);
}

template <class ALoader>
static int launch_gemm(ALoader a, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
                       int64_t M, int N, int K, cudaStream_t stream, void* splitk_ws = nullptr, size_t splitk_ws_bytes = 0,
                       unsigned* splitk_counters = nullptr, bool pdl = false) {
    VQ_ARG(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
    VQ_ARG(K % GEMM_BK == 0, "gemm: K=%d must be a multiple of %d", K, GEMM_BK);
    VQ_ARG(N % 4 == 0 && ldc % 4 == 0 && ldw % 4 == 0, "gemm: N, ldc, ldw must be multiples of 4");
    if (M == 0) return VQCPC_OK;
    // Large problems: 128x128 tiles.  Small M (single-utterance latency case): 64x64 tiles for more CTAs.
    const int64_t tiles128 = ((M + 127) / 128) * ((N + 127) / 128);
    if (tiles128 >= 2 * 148) {
        dim3 grid(static_cast<unsigned>((M + 127) / 128), static_cast<unsigned>((N + 127) / 128));
        sgemm_tn_kernel<2, 2, ALoader><<<grid, 256, 0, stream>>>(a, W, ldw, bias, C, ldc, M, N, K, nullptr, nullptr);
    } else {
        dim3 grid(static_cast<unsigned>((M + 63) / 64), static_cast<unsigned>((N + 63) / 64));
        // latency case, opted into by the caller passing the split-K workspace (Encoder.encode of one utterance; everywhere else the
        // summation order must not depend on the batch size): split K so that about one CTA per SM is busy.  Slices of up to
        // 16 * SK_NKMAX columns run on the cluster kernel, longer ones through the scratch planes and counters.
        int S = 1;
        if (splitk_ws != nullptr && splitk_counters != nullptr && M <= SPLITK_MAX_M && splitk_ws_bytes >= gemm_splitk_ws_bytes(M, N) &&
            grid.x * grid.y <= static_cast<unsigned>(SPLITK_COUNTERS)) {
            const int tiles = static_cast<int>(grid.x * grid.y);
            for (int cand = SPLITK_MAX_S; cand >= 2; --cand)
                if (K % (GEMM_BK * cand) == 0 && K / cand >= 64 && tiles * cand <= 2 * device_sm_count()) { S = cand; break; }
        }
        // pdl: the caller's latency path is a chain of small kernels (programmatic dependent launches, common.cuh)
        if (S > 1 && K / S <= GEMM_BK * SK_NKMAX) {
            grid.z = S;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = static_cast<unsigned>(S);
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
            VQ_CUDA(cudaLaunchKernelEx(&cfg, sgemm_splitk_cluster_kernel<ALoader>, a, W, ldw, bias, C, ldc, M, N, K));
        } else if (S > 1) {
            grid.z = S;
            VQ_CUDA(launch_pdl(pdl, sgemm_tn_kernel<1, 1, ALoader>, grid, dim3(256), 0, stream, a, W, ldw, bias, C, ldc, M, N, K,
                               static_cast<float*>(splitk_ws), splitk_counters));
        } else {
            VQ_CUDA(launch_pdl(pdl, sgemm_tn_kernel<1, 1, ALoader>, grid, dim3(256), 0, stream, a, W, ldw, bias, C, ldc, M, N, K,
                               static_cast<float*>(nullptr), static_cast<unsigned*>(nullptr)));
        }
    }
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}

int gemm_dense(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
               int64_t M, int N, int K, cudaStream_t stream, void* splitk_ws, size_t splitk_ws_bytes, unsigned* splitk_counters, bool pdl) {
    if (M == 0) return VQCPC_OK;
    VQ_ARG(A && W && C, "gemm: null pointer");
    VQ_ARG(lda % 4 == 0, "gemm: lda must be a multiple of 4");
    DenseA a{A, lda};
    return launch_gemm(a, W, ldw, bias, C, ldc, M, N, K, stream, splitk_ws, splitk_ws_bytes, splitk_counters, pdl);
}

int gemm_conv(const float* mel, int B, int T, int Cin, const float* W, float* C, int Cout, cudaStream_t stream, void* splitk_ws,
              size_t splitk_ws_bytes, unsigned* splitk_counters) {
    VQ_ARG(mel && W && C, "conv: null pointer");
    VQ_ARG(T >= 2, "conv: T=%d must be >= 2", T);
    const int Tp = (T - 2) / 2 + 1;
    ConvA a{mel, T, Tp, Cin};
    return launch_gemm(a, W, Cin * 4, nullptr, C, Cout, static_cast<int64_t>(B) * Tp, Cout, Cin * 4, stream, splitk_ws, splitk_ws_bytes, splitk_counters);
}

}  // namespace vqcpc

extern "C" int vqcpc_linear_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C,
                                int64_t ldc, int64_t M, int32_t N, int32_t K, void* stream) {
    return vqcpc::gemm_dense(A, lda, W, ldw, bias, C, ldc, M, N, K, static_cast<cudaStream_t>(stream), nullptr, 0, nullptr);
}
