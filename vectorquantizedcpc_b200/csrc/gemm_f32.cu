// fp32 SIMT GEMM  C[m,n] = sum_k A(m,k) * W[n,k] (+ bias[n])  -- the exact-arithmetic ("fp32 parity") path
// behind nn.Conv1d / nn.Linear of /root/reference/model.py:43,50,54 and the vocoder's input projections.
// A is supplied by a loader functor: dense row-major, or the implicit im2col of the stride-2 conv.
//
// Tiling: BM x BN x 16 per CTA, 256 threads as a 16x16 grid, each thread a (4*RM) x (4*RN) micro-tile read
// from k-major shared tiles with LDS.128; global tiles are register-prefetched one k-step ahead (double
// buffered shared memory, one __syncthreads per k-step).
//
// Split-K (latency case, M <= 256: one 2 s utterance is M = 100, which gives 24 CTAs a 768-deep serial k loop each: 47 us per
// layer): gridDim.z CTAs share an output tile, each sums a K / gridDim.z slice into a scratch plane, and the LAST one to arrive
// (per-tile counter) adds the planes in fixed order s = 0, 1, ... and writes C -- deterministic, no extra launch.  The order
// of the fp32 sum differs from the unsplit kernel's (error ~1e-7 relative; the parity bar is 1e-4).
#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

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
        // latency case with a workspace: split K so that about one CTA per SM is busy
        int S = 1;
        if (splitk_ws != nullptr && splitk_counters != nullptr && M <= SPLITK_MAX_M && splitk_ws_bytes >= gemm_splitk_ws_bytes(M, N) &&
            grid.x * grid.y <= static_cast<unsigned>(SPLITK_COUNTERS)) {
            const int tiles = static_cast<int>(grid.x * grid.y);
            for (int cand = SPLITK_MAX_S; cand >= 2; --cand)
                if (K % (GEMM_BK * cand) == 0 && K / cand >= 64 && tiles * cand <= 2 * device_sm_count()) { S = cand; break; }
        }
        // pdl: the caller's latency path is a chain of small kernels (programmatic dependent launches, common.cuh)
        if (S > 1) {
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
