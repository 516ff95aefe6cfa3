// Output stage of convert.py: integrated loudness (ITU-R BS.1770-4 as implemented by pyloudnorm,
// which /root/reference/convert.py:50,57,79-80 calls: Meter(sr).integrated_loudness, normalize.loudness) and the gain.
//
//   K-weighting = high shelf (+4 dB, Q 1/sqrt 2, 1500 Hz) then high pass (Q 0.5, 38 Hz), RBJ biquads for the given rate
//   z_j   = mean square of the filtered signal over 400 ms blocks with 75 % overlap
//   l_j   = -0.691 + 10 log10 z_j ;  absolute gate -70 LUFS ;  relative gate = loudness of the abs-gated blocks - 10 LU
//   LUFS  = -0.691 + 10 log10 mean{ z_j : l_j > both gates }
//   out   = wav * 10^((target - LUFS) / 20)                                                 convert.py:80
//
// One CTA per utterance.  The two cascaded biquads are a sequential recurrence: lane 0 runs it in fp64 (the 38 Hz high
// pass at 16 kHz has poles at |z| = 0.985, fp32 state would cost ~1e-4 LU) at ~25 cycles per sample -- 0.6 ms for 3 s,
// all utterances in parallel, three orders of magnitude below the generate call it follows.  Block energies, gating and
// the final reduction use the whole CTA.
#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

struct Biquad { double b0, b1, b2, a1, a2; };

__global__ void __launch_bounds__(256) loudness_kernel(const float* __restrict__ wave, const int32_t* __restrict__ lengths, int N,
                                                       int rate, Biquad shelf, Biquad hp, float* __restrict__ filt,
                                                       double* __restrict__ zblk, int max_blocks, float* __restrict__ lufs) {
    __shared__ double red[8];
    __shared__ double gamma_r_s;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nb = lengths ? min(N, lengths[b]) : N;
    const float* x = wave + static_cast<int64_t>(b) * N;
    float* y = filt + static_cast<int64_t>(b) * N;
    if (tid == 0) {
        // transposed direct form II, the arithmetic of scipy.signal.lfilter (pyloudnorm's apply_filter)
        double s1 = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < nb; ++i) {
            const double u = static_cast<double>(__ldg(x + i));
            const double v = shelf.b0 * u + s1;
            s1 = shelf.b1 * u - shelf.a1 * v + s2;
            s2 = shelf.b2 * u - shelf.a2 * v;
            const double w = hp.b0 * v + t1;
            t1 = hp.b1 * v - hp.a1 * w + t2;
            t2 = hp.b2 * v - hp.a2 * w;
            y[i] = static_cast<float>(w);
        }
    }
    __syncthreads();
    // blocks: T_g = 0.4 s, step 0.25;  numBlocks = round((T - T_g) / (T_g * step)) + 1;  l = int(T_g * j * step * rate), u = int(T_g * (j * step + 1) * rate)
    const double Tg = 0.4, step = 0.25;
    const double T = static_cast<double>(nb) / rate;
    int n_blocks = static_cast<int>(rint((T - Tg) / (Tg * step))) + 1;
    if (n_blocks < 0) n_blocks = 0;
    if (n_blocks > max_blocks) n_blocks = max_blocks;
    double* z = zblk + static_cast<int64_t>(b) * max_blocks;
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < n_blocks; j += 8) {
        const int l = static_cast<int>(Tg * (j * step) * rate), u = min(nb, static_cast<int>(Tg * (j * step + 1.0) * rate));
        double acc = 0.0;
        for (int i = l + lane; i < u; i += 32) { const double v = y[i]; acc += v * v; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) z[j] = acc / (Tg * rate);
    }
    __syncthreads();
    auto block_sum2 = [&](double v, double c, double& sum_out, double& cnt_out) {
        // CTA-wide sums of (v, c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double sv = 0.0;
        for (int k = 0; k < 8; ++k) sv += red[k];
        __syncthreads();
        if (lane == 0) red[warp] = c;
        __syncthreads();
        double sc = 0.0;
        for (int k = 0; k < 8; ++k) sc += red[k];
        sum_out = sv; cnt_out = sc;
    };
    // pass 1: absolute gate
    double v = 0.0, c = 0.0;
    for (int j = tid; j < n_blocks; j += 256) {
        const double lj = -0.691 + 10.0 * log10(z[j]);
        if (lj >= -70.0) { v += z[j]; c += 1.0; }
    }
    double sv, sc;
    block_sum2(v, c, sv, sc);
    if (tid == 0) gamma_r_s = -0.691 + 10.0 * log10(sv / sc) - 10.0;      // NaN when nothing passes, as in pyloudnorm
    __syncthreads();
    const double gamma_r = gamma_r_s;
    // pass 2: both gates
    v = 0.0; c = 0.0;
    for (int j = tid; j < n_blocks; j += 256) {
        const double lj = -0.691 + 10.0 * log10(z[j]);
        if (lj > gamma_r && lj > -70.0) { v += z[j]; c += 1.0; }
    }
    block_sum2(v, c, sv, sc);
    if (tid == 0) {
        const double mean = sc > 0.0 ? sv / sc : 0.0;                       // np.nan_to_num(np.mean([])) = 0
        lufs[b] = static_cast<float>(-0.691 + 10.0 * log10(mean));          // -inf for silence, as in pyloudnorm
    }
}

__global__ void gain_kernel(const float* __restrict__ wave, const int32_t* __restrict__ lengths, const float* __restrict__ measured,
                            const float* __restrict__ target, float* __restrict__ out, int B, int N) {
    const int64_t total = static_cast<int64_t>(B) * N;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / N), n = static_cast<int>(i % N);
        const int nb = lengths ? min(N, lengths[b]) : N;
        const float g = exp10f((target[b] - measured[b]) / 20.0f);          // normalize.loudness: gain = 10^(delta / 20)
        out[i] = n < nb ? g * wave[i] : 0.f;
    }
}

static Biquad make_biquad(double G, double Q, double fc, double rate, bool high_shelf) {
    // pyloudnorm IIRfilter.generate_coefficients (RBJ cookbook forms), normalised by a0
    const double PI = 3.14159265358979323846;
    const double A = pow(10.0, G / 40.0), w0 = 2.0 * PI * (fc / rate), alpha = sin(w0) / (2.0 * Q), cw = cos(w0);
    double b0, b1, b2, a0, a1, a2;
    if (high_shelf) {
        b0 = A * ((A + 1) + (A - 1) * cw + 2 * sqrt(A) * alpha);
        b1 = -2 * A * ((A - 1) + (A + 1) * cw);
        b2 = A * ((A + 1) + (A - 1) * cw - 2 * sqrt(A) * alpha);
        a0 = (A + 1) - (A - 1) * cw + 2 * sqrt(A) * alpha;
        a1 = 2 * ((A - 1) - (A + 1) * cw);
        a2 = (A + 1) - (A - 1) * cw - 2 * sqrt(A) * alpha;
    } else {
        b0 = (1 + cw) / 2; b1 = -(1 + cw); b2 = (1 + cw) / 2;
        a0 = 1 + alpha; a1 = -2 * cw; a2 = 1 - alpha;
    }
    return Biquad{b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
}

static int loud_max_blocks(int N, int rate) { return static_cast<int>(static_cast<double>(N) / rate / 0.1) + 2; }
static size_t loud_ws_bytes(int B, int N, int rate) {
    return align_up(sizeof(float) * static_cast<size_t>(B) * N, 256) + align_up(sizeof(double) * static_cast<size_t>(B) * loud_max_blocks(N, rate), 256);
}

int loudness_measure(const float* wave, const int32_t* lengths, int B, int N, int rate, void* ws, size_t ws_bytes, float* out_lufs,
                     cudaStream_t stream) {
    VQ_ARG(B >= 0 && N >= 0 && rate > 0, "loudness: bad shape B=%d N=%d rate=%d", B, N, rate);
    if (B == 0) return VQCPC_OK;
    VQ_ARG(wave && ws && out_lufs, "loudness: null pointer");
    VQ_ARG(ws_bytes >= loud_ws_bytes(B, N, rate), "loudness: workspace too small");
    float* filt = static_cast<float*>(ws);
    double* z = reinterpret_cast<double*>(static_cast<unsigned char*>(ws) + align_up(sizeof(float) * static_cast<size_t>(B) * N, 256));
    const Biquad shelf = make_biquad(4.0, 1.0 / sqrt(2.0), 1500.0, rate, true), hp = make_biquad(0.0, 0.5, 38.0, rate, false);
    loudness_kernel<<<B, 256, 0, stream>>>(wave, lengths, N, rate, shelf, hp, filt, z, loud_max_blocks(N, rate), out_lufs);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}

}  // namespace vqcpc

extern "C" size_t vqcpc_loudness_workspace_bytes(int32_t B, int32_t N, int32_t rate) {
    if (B < 0 || N < 0 || rate <= 0) return 0;
    return vqcpc::loud_ws_bytes(B, N, rate);
}
extern "C" int vqcpc_integrated_loudness(const float* wave, const int32_t* lengths, int32_t B, int32_t N, int32_t rate, void* workspace,
                                         size_t workspace_bytes, float* out_lufs, void* stream) {
    return vqcpc::loudness_measure(wave, lengths, B, N, rate, workspace, workspace_bytes, out_lufs, static_cast<cudaStream_t>(stream));
}
extern "C" int vqcpc_loudness_normalize(const float* wave, const int32_t* lengths, const float* target_lufs, int32_t B, int32_t N,
                                        int32_t rate, void* workspace, size_t workspace_bytes, float* out_wave, float* out_measured_lufs,
                                        void* stream) {
    using namespace vqcpc;
    if (B == 0) return VQCPC_OK;
    VQ_ARG(target_lufs && out_wave && out_measured_lufs, "loudness_normalize: null pointer");
    int rc = loudness_measure(wave, lengths, B, N, rate, workspace, workspace_bytes, out_measured_lufs, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    const int64_t total = static_cast<int64_t>(B) * N;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    gain_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(wave, lengths, out_measured_lufs, target_lufs, out_wave, B, N);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}
