// Log-mel front-end: the step immediately before Encoder.encode.
//
// Restates /root/reference/preprocess.py:53-75 (wave_to_mel; same arithmetic inline in convert.py:54-70):
//   wave_s = wave / max|wave| * 0.999                               preprocess.py:62
//   y      = lfilter([1, -preemph], [1], wave_s)                    preprocess.py:16-17, 65
//   S      = | STFT(y; n_fft 2048, hann(win 400), hop 160, center, reflect) |      librosa.feature.melspectrogram, power=1
//   mel    = melW(80 x 1025, Slaney scale + Slaney norm, fmin 50) . S              preprocess.py:65-72
//   logmel = max(20 log10(max(1e-5, mel)), max_over_utterance - top_db) / top_db + 1      preprocess.py:73-74
//
// B200 mapping: the hann window has 400 non-zero taps inside the 2048-point frame, so the DFT of a frame is a
// (400 -> 2 x 1025) real matrix product -- frames x [cos | sin] -- and the mel projection a second one; both run on the
// fp32 GEMM of gemm_f32.cu.  Peak scaling, pre-emphasis, reflection padding, framing and windowing are fused into the
// kernel that writes the frame matrix (the padded signal is never materialised); the dB conversion, the per-utterance
// top_db clamp and the transpose to the (B, n_mels, T) layout Encoder.encode takes are fused into the last kernel.
// Ragged batches: `lengths` gives the samples of each utterance; frames beyond 1 + lengths[b] / hop are written as 0.
#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

// max |x| over an utterance: one CTA per utterance
__global__ void __launch_bounds__(256) fe_peak_kernel(const float* __restrict__ wave, const int32_t* __restrict__ lengths, int N,
                                                      float* __restrict__ peak) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const int nb = lengths ? min(N, lengths[b]) : N;
    const float* x = wave + static_cast<int64_t>(b) * N;
    float m = 0.f;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(x + i)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 8) {
        m = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
        if (threadIdx.x == 0) peak[b] = m;
    }
}

// frame matrix A[(b, t)][j] = window[j] * y_b[reflect(t * hop - win / 2 + j)],  y = pre-emphasised, peak-scaled wave
__global__ void fe_frames_kernel(const float* __restrict__ wave, const int32_t* __restrict__ lengths, const float* __restrict__ peak,
                                 const float* __restrict__ window, float* __restrict__ A, int B, int N, int T, int win, int hop,
                                 float preemph) {
    const int64_t total = static_cast<int64_t>(B) * T * win;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(i % win);
        const int64_t bt = i / win;
        const int t = static_cast<int>(bt % T), b = static_cast<int>(bt / T);
        const int nb = lengths ? min(N, lengths[b]) : N;
        float v = 0.f;
        if (t <= nb / hop) {
            int idx = t * hop - win / 2 + j;
            if (idx < 0) idx = -idx;                                  // numpy "reflect": the edge sample is not repeated
            if (idx >= nb) idx = 2 * (nb - 1) - idx;
            const float* x = wave + static_cast<int64_t>(b) * N;
            const float s = 0.999f / peak[b];
            const float cur = __ldg(x + idx) * s;
            const float prev = idx > 0 ? __ldg(x + idx - 1) * s : 0.f;   // lfilter starts from a zero state
            v = __ldg(window + j) * (cur - preemph * prev);
        }
        A[i] = v;
    }
}

// mag[m][k] = |re + i im| for k < n_freq, 0 in the padding columns; spec = [re (nfp cols) | im (nfp cols)]
__global__ void fe_mag_kernel(const float* __restrict__ spec, float* __restrict__ mag, int64_t M, int n_freq, int nfp) {
    const int64_t total = M * nfp;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % nfp);
        const int64_t m = i / nfp;
        float v = 0.f;
        if (k < n_freq) {
            const float re = spec[m * 2 * nfp + k], im = spec[m * 2 * nfp + nfp + k];
            v = sqrtf(re * re + im * im);
        }
        mag[i] = v;
    }
}

// per utterance: max of 20 log10(max(amin, mel)) over its valid frames; one CTA per utterance
__global__ void __launch_bounds__(256) fe_dbmax_kernel(const float* __restrict__ mel, const int32_t* __restrict__ lengths, int N, int T,
                                                       int hop, int n_mels, float* __restrict__ dbmax) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const int nb = lengths ? min(N, lengths[b]) : N;
    const int Tb = 1 + nb / hop;
    const float* src = mel + static_cast<int64_t>(b) * T * n_mels;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < Tb * n_mels; i += blockDim.x) m = fmaxf(m, 20.0f * log10f(fmaxf(1e-5f, src[i])));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 8) {
        m = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
        if (threadIdx.x == 0) dbmax[b] = m;
    }
}

// out[b][f][t] = max(db, dbmax[b] - top_db) / top_db + 1   (0 beyond the utterance's last frame)
__global__ void fe_finalize_kernel(const float* __restrict__ mel, const int32_t* __restrict__ lengths, const float* __restrict__ dbmax,
                                   float* __restrict__ out, int B, int N, int T, int hop, int n_mels, float top_db) {
    const int64_t total = static_cast<int64_t>(B) * n_mels * T;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i % T);
        const int64_t bf = i / T;
        const int f = static_cast<int>(bf % n_mels), b = static_cast<int>(bf / n_mels);
        const int nb = lengths ? min(N, lengths[b]) : N;
        float v = 0.f;
        if (t <= nb / hop) {
            const float db = 20.0f * log10f(fmaxf(1e-5f, mel[(static_cast<int64_t>(b) * T + t) * n_mels + f]));
            v = fmaxf(db, dbmax[b] - top_db) / top_db + 1.0f;
        }
        out[i] = v;
    }
}

static size_t logmel_ws_bytes(int B, int N, const vqcpc_logmel_config* c) {
    const int64_t T = 1 + N / c->hop_length, M = static_cast<int64_t>(B) * T;
    const int nfp = c->n_freq_padded;
    return align_up(sizeof(float) * 2 * B, 256) + align_up(sizeof(float) * M * c->win_length, 256) +
           align_up(sizeof(float) * M * 2 * nfp, 256) + align_up(sizeof(float) * M * nfp, 256) +
           align_up(sizeof(float) * M * c->n_mels, 256);
}

static int check_cfg(const vqcpc_logmel_config* c) {
    VQ_ARG(c != nullptr, "logmel: null config");
    VQ_ARG(c->n_fft >= 16 && c->n_fft % 2 == 0 && c->win_length >= 16 && c->win_length <= c->n_fft && c->win_length % 16 == 0,
           "logmel: n_fft=%d / win_length=%d unsupported (win_length must be a multiple of 16, <= n_fft)", c->n_fft, c->win_length);
    VQ_ARG(c->hop_length >= 1 && c->n_mels >= 4 && c->n_mels % 4 == 0, "logmel: hop_length=%d / n_mels=%d unsupported", c->hop_length,
           c->n_mels);
    VQ_ARG(c->n_freq_padded >= c->n_fft / 2 + 1 && c->n_freq_padded % 16 == 0, "logmel: n_freq_padded=%d must be a multiple of 16 >= %d",
           c->n_freq_padded, c->n_fft / 2 + 1);
    VQ_ARG(c->top_db > 0.f, "logmel: top_db must be positive");
    return VQCPC_OK;
}

int logmel_forward(const vqcpc_logmel_config* c, const float* wave, const int32_t* lengths, int B, int N, const float* window,
                   const float* dft, const float* melw, void* ws, size_t ws_bytes, float* out, cudaStream_t stream) {
    int rc = check_cfg(c);
    if (rc) return rc;
    if (B == 0) return VQCPC_OK;
    VQ_ARG(wave && window && dft && melw && ws && out, "logmel: null pointer");
    VQ_ARG(B > 0 && N > c->n_fft / 2, "logmel: need more than n_fft/2 = %d samples per utterance (reflect padding), got %d", c->n_fft / 2, N);
    VQ_ARG(ws_bytes >= logmel_ws_bytes(B, N, c), "logmel: workspace too small");
    const int T = 1 + N / c->hop_length, nfp = c->n_freq_padded, n_freq = c->n_fft / 2 + 1;
    const int64_t M = static_cast<int64_t>(B) * T;
    unsigned char* p = static_cast<unsigned char*>(ws);
    float* peak = reinterpret_cast<float*>(p);
    float* dbmax = peak + B;
    p += align_up(sizeof(float) * 2 * B, 256);
    float* A = reinterpret_cast<float*>(p); p += align_up(sizeof(float) * M * c->win_length, 256);
    float* spec = reinterpret_cast<float*>(p); p += align_up(sizeof(float) * M * 2 * nfp, 256);
    float* mag = reinterpret_cast<float*>(p); p += align_up(sizeof(float) * M * nfp, 256);
    float* mel = reinterpret_cast<float*>(p);
    auto blocks = [](int64_t total) { return static_cast<unsigned>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16); };

    fe_peak_kernel<<<B, 256, 0, stream>>>(wave, lengths, N, peak);
    fe_frames_kernel<<<blocks(M * c->win_length), 256, 0, stream>>>(wave, lengths, peak, window, A, B, N, T, c->win_length,
                                                                    c->hop_length, c->preemph);
    VQ_CUDA(cudaGetLastError());
    count_launch(2);
    // DFT of the windowed taps as a product with [cos | sin] (2 nfp rows x win), then the mel projection
    if ((rc = gemm_dense(A, c->win_length, dft, c->win_length, nullptr, spec, 2 * nfp, M, 2 * nfp, c->win_length, stream))) return rc;
    fe_mag_kernel<<<blocks(M * nfp), 256, 0, stream>>>(spec, mag, M, n_freq, nfp);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    if ((rc = gemm_dense(mag, nfp, melw, nfp, nullptr, mel, c->n_mels, M, c->n_mels, nfp, stream))) return rc;
    fe_dbmax_kernel<<<B, 256, 0, stream>>>(mel, lengths, N, T, c->hop_length, c->n_mels, dbmax);
    fe_finalize_kernel<<<blocks(M * c->n_mels), 256, 0, stream>>>(mel, lengths, dbmax, out, B, N, T, c->hop_length, c->n_mels, c->top_db);
    VQ_CUDA(cudaGetLastError());
    count_launch(2);
    return VQCPC_OK;
}

}  // namespace vqcpc

extern "C" size_t vqcpc_logmel_workspace_bytes(const vqcpc_logmel_config* cfg, int32_t B, int32_t N) {
    if (cfg == nullptr || B < 0 || N < 0 || cfg->hop_length < 1) return 0;
    return vqcpc::logmel_ws_bytes(B, N, cfg);
}
extern "C" int vqcpc_logmel_forward(const vqcpc_logmel_config* cfg, const float* wave, const int32_t* lengths, int32_t B, int32_t N,
                                    const float* window, const float* dft, const float* melw, void* workspace,
                                    size_t workspace_bytes, float* out, void* stream) {
    return vqcpc::logmel_forward(cfg, wave, lengths, B, N, window, dft, melw, workspace, workspace_bytes, out,
                                 static_cast<cudaStream_t>(stream));
}
