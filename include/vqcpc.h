/*
 * vqcpc.h -- C-ABI of the B200-native VQ-CPC inference hot path (libvqcpc_b200.so).
 *
 * The reference (tarepan/VectorQuantizedCPC) is pure Python and has NO plugin/FFI boundary; its
 * boundary for this path is three Python methods.  Each entry point below names the
 * reference interface it replaces.  All pointers are DEVICE pointers (plain C types only, no torch
 * types); `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success and a
 * non-zero code on failure, with a thread-local message retrievable through vqcpc_last_error().
 * Outputs are written into caller-allocated buffers; inputs are never modified.
 *
 * Built for sm_100a only.  There is no CPU fallback.
 */
#ifndef VQCPC_H
#define VQCPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQCPC_ABI_VERSION 1

/* error codes */
#define VQCPC_OK 0
#define VQCPC_ERR_ARG 1      /* bad shape / null pointer / unsupported dimension */
#define VQCPC_ERR_CUDA 2     /* a CUDA runtime call failed */
#define VQCPC_ERR_DEVICE 3   /* device is not sm_100 / cannot co-schedule the persistent grid */
#define VQCPC_ERR_TIMEOUT 4  /* a persistent kernel's exchange timed out (see vqcpc_check_status) */
#define VQCPC_ERR_INDEX 5    /* a code / speaker index was out of range (nn.Embedding would raise IndexError); the gather
                              * clamps it and vqcpc_check_status reports it -- no host synchronisation before the launch */

const char* vqcpc_last_error(void);
int vqcpc_abi_version(void);
/* Number of kernels this library has launched in the calling process so far (diagnostic counter). */
uint64_t vqcpc_launch_count(void);
/* 0 iff `device` is compute capability 10.x with >= 128 SMs (what the persistent kernels need). */
int vqcpc_device_check(int device);

/* ---- weights of Encoder: state_dict layout of /root/reference/model.py:43-57,90-101 */
typedef struct {
    int32_t in_channels;   /* 80  ConfEncoder.in_channels  model.py:27 */
    int32_t channels;      /* C: 512 or 768 (multiple of 128, <= 1024)  model.py:28 */
    int32_t n_embeddings;  /* 512 model.py:29 */
    int32_t z_dim;         /* 64  model.py:30 */
    int32_t c_dim;         /* 256 model.py:31 */
    int32_t _pad;
    const float* conv_w;      /* conv.weight (C, 80, 4) contiguous == (C, 320) */
    const float* ln_w[5];     /* encoder.{0,3,6,9,12}.weight (C,) */
    const float* ln_b[5];     /* encoder.{0,3,6,9,12}.bias   (C,) */
    const float* fc_w[4];     /* encoder.{2,5,8,11}.weight (C, C) */
    const float* proj_w;      /* encoder.14.weight (64, C) */
    const float* proj_b;      /* encoder.14.bias (64,) */
    const float* codebook;    /* codebook.embedding (512, 64) */
    const float* lstm_w_ih;   /* rnn.weight_ih_l0 (1024, 64) */
    const float* lstm_w_hh;   /* rnn.weight_hh_l0 (1024, 256) */
    const float* lstm_b;      /* rnn.bias_ih_l0 + rnn.bias_hh_l0 (1024,)  (host-side sum) */
    /* bf16 hi/lo planes of the GEMM weights for the tensor-core mode ([hi | lo], 2K columns, built once with
     * vqcpc_split_planes); may be NULL when only VQCPC_GEMM_FP32 is used. */
    const void* conv_wp;      /* (C, 2*320) */
    const void* fc_wp[4];     /* (C, 2*C) */
    const void* proj_wp;      /* (64, 2*C) */
    const void* lstm_whh_p;   /* (1024, 2*256): planes of rnn.weight_hh_l0 for the batched (B >= 33) LSTM */
    /* optional, weight-only: the LSTM input projection of every CODE, table[m] = lstm_w_ih codebook[m] + lstm_b (512, 1024),
     * exactly vqcpc_linear_f32(codebook, lstm_w_ih, lstm_b).  NULL: recomputed inside every call (one 512 x 1024 x 64 GEMM). */
    const float* lstm_table;
} vqcpc_encoder_weights;

/* GEMM arithmetic of Encoder.encode */
#define VQCPC_GEMM_FP32 0     /* fp32 FMA on the CUDA cores: exact-order parity path */
#define VQCPC_GEMM_BF16X3 1   /* tcgen05 tensor cores, bf16 hi/lo split (hi*hi + hi*lo + lo*hi), fp32 accumulate */
#define VQCPC_GEMM_BF16 2     /* speed mode: the conv / MLP / projection products in single-pass bf16 (fp32 accumulate, fp32
                               * LayerNorm); the VQ search stays exact for the z it is given and the LSTM stays bf16x3, so c is
                               * fp32-grade GIVEN the indices.  Stated bound (tests/test_gpu_parity.py::test_encoder_bf16_mode):
                               * pre-VQ z within 3e-2 * max|z| of the oracle, >= 97 % of the indices equal on trained-like
                               * weights. */

/* ---- weights of Vocoder: /root/reference/network_vocoder.py:37-39 + rnnms dims config.py:62-77,199 */
typedef struct {
    int32_t n_codes;       /* 512 size_i_codebook */
    int32_t dim_code;      /* 64  dim_i_embedding */
    int32_t n_speakers;    /* 102 */
    int32_t dim_speaker;   /* 64  dim_speaker_embedding */
    int32_t upsample_t;    /* 160 upsampling_t (hop length) */
    int32_t _pad;
    const float* code_emb;      /* code_embedding.weight (512, 64) */
    const float* spk_emb;       /* speaker_embedding.weight (n_speakers, 64) */
    /* prenet: 2-layer bidirectional GRU, hidden 128/direction; per layer, directions concatenated */
    const float* pre_w_ih[2];   /* [fwd;bwd] weight_ih (768, 128) / (768, 256) */
    const float* pre_b_ih[2];   /* [fwd;bwd] bias_ih (768,) */
    const float* pre_w_hh[2];   /* [fwd;bwd] weight_hh (2, 384, 128) */
    const float* pre_b_hh[2];   /* [fwd;bwd] bias_hh (2, 384) */
    /* autoregressive part */
    const float* ar_w_ih;       /* rnn weight_ih (2688, 512): [:, :256] acts on the sample embedding */
    const float* ar_b_ih;       /* (2688,) */
    const float* ar_w_hh;       /* (2688, 896) */
    const float* ar_b_hh;       /* (2688,) */
    const float* fc1_w;         /* (256, 896) */
    const float* fc1_b;         /* (256,) */
    const float* fc2_w;         /* (256, 256) */
    const float* fc2_b;         /* (256,) */
    const float* ar_emb;        /* embedding (256, 256) */
    const float* eprime;        /* (256, 2688) = ar_emb . ar_w_ih[:, :256]^T, filled by vqcpc_vocoder_pack */
    const float* mulaw_lut;     /* (256,) k -> wav, formula of /root/reference/preprocess.py:30-35 */
} vqcpc_vocoder_weights;

/* ------------------------------------------------------------------ building blocks (also used by tests)
 * C[m,n] = sum_k A[m*lda+k] * W[n*ldw+k] (+ bias[n]);  fp32, K % 16 == 0, N % 4 == 0, 16-byte aligned rows.
 * Replaces the implicit cuBLAS calls behind nn.Linear (model.py:50,54). */
int vqcpc_linear_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                     float* C, int64_t ldc, int64_t M, int32_t N, int32_t K, void* stream);
/* Same product on the tensor cores (tcgen05 / TMEM accumulators / TMA operand tiles).  A (M, K) and W (N, K) fp32 are
 * split into bf16 hi/lo planes (a_planes: M x 2K bf16, w_planes: N x 2K bf16, caller-provided scratch) and multiplied
 * as hi*hi + hi*lo + lo*hi with fp32 accumulation (mode 3; error ~2^-16 relative).  K % 64 == 0, N % 64 == 0.
 * err_flag: device int, set non-zero if a pipeline wait timed out. */
int vqcpc_linear_tc(const float* A, const float* W, const float* bias, float* C, int64_t M, int32_t N, int32_t K,
                    int32_t mode, void* a_planes, void* w_planes, int32_t* err_flag, void* stream);
/* fp32 (rows, K) with leading dimension ld -> bf16 planes (rows, 2K): [bf16(x) | bf16(x - bf16(x))]. */
int vqcpc_split_planes(const float* x, int64_t ld, void* out_planes, int64_t rows, int32_t K, void* stream);
/* In-place relu(LayerNorm(x)) over rows of width C (biased variance, eps 1e-5) -- model.py:47-48,51-52. */
int vqcpc_layernorm_relu_f32(float* x, const float* w, const float* b, int64_t rows, int32_t C, void* stream);

/* ------------------------------------------------------------------ VQEmbeddingEMA.encode -- model.py:103-115
 * x (n_frames, 64) fp32 -> out_q (n_frames, 64) = codebook[idx], out_idx (n_frames,) int64 = first argmin of
 * |e|^2 - 2 x.e (fp32; the |x|^2 term of model.py:107-110 is argmin-invariant and is not added). */
int vqcpc_vq_lookup(const float* x, const float* codebook, int64_t n_frames, int32_t n_codes, int32_t dim,
                    float* out_q, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);
/* workspace == NULL: the exact fp32 SIMT search.  With a workspace of vqcpc_vq_workspace_bytes() bytes (caller-owned, one per
 * stream in flight -- the library keeps no hidden state) and n_frames >= 8192 the search runs on the tensor cores: tcgen05
 * coarse pass over bf16 hi/lo planes, and an exact fp32 rescan of all 512 codes for every frame whose coarse minimum is
 * not the only score within twice the coarse error bound of it (any other frame provably has its coarse winner as exact
 * winner), so results are identical.  vqcpc_check_status(workspace, stream) returns VQCPC_ERR_TIMEOUT if that pipeline timed out. */
size_t vqcpc_vq_workspace_bytes(void);

/* ------------------------------------------------------------------ Encoder.encode -- model.py:59-70
 * mel (B, 80, T) fp32 -> out_z (B, T', 64) quantised, out_c (B, T', 256), out_idx (B, T') int64,
 * T' = (T-2)/2+1.  Optional (nullable): out_prevq (B, T', 64) = output of encoder.encoder[-1] (what the
 * forward hook of encode.py:34-40 observes); out_hidden (B, T', C) = its input; out_c (NULL: the recurrence is skipped,
 * see vqcpc_lstm_forward_ex). */
size_t vqcpc_encoder_workspace_bytes(int32_t B, int32_t T, int32_t channels);
int vqcpc_encoder_forward(const vqcpc_encoder_weights* w, const float* mel, int32_t B, int32_t T,
                          void* workspace, size_t workspace_bytes,
                          float* out_z, float* out_c, int64_t* out_idx,
                          float* out_prevq, float* out_hidden, void* stream);
/* Same with an explicit GEMM arithmetic (VQCPC_GEMM_*); workspace from vqcpc_encoder_workspace_bytes_ex. */
size_t vqcpc_encoder_workspace_bytes_ex(int32_t B, int32_t T, int32_t channels, int32_t gemm_mode);
int vqcpc_encoder_forward_ex(const vqcpc_encoder_weights* w, const float* mel, int32_t B, int32_t T,
                             void* workspace, size_t workspace_bytes, float* out_z, float* out_c, int64_t* out_idx,
                             float* out_prevq, float* out_hidden, int32_t gemm_mode, void* stream);
/* nn.LSTM(64,256) over quantised codes only (model.py:57,69): idx (B, T') -> out_c (B, T', 256). */
size_t vqcpc_lstm_workspace_bytes(int32_t B, int32_t Tp);
int vqcpc_lstm_forward(const vqcpc_encoder_weights* w, const int64_t* idx, int32_t B, int32_t Tp,
                       void* workspace, size_t workspace_bytes, float* out_c, void* stream);
/* Same with an explicit arithmetic (VQCPC_GEMM_*): the tensor-core modes run batches of >= 33 utterances as one persistent
 * tcgen05 launch (lstm_whh_p planes present), otherwise as the fp32 entry above.  Together with out_c == NULL in
 * vqcpc_encoder_forward_ex (front part only: conv .. VQ, no recurrence) this lets a caller encode a large batch in chunks
 * -- overlapping each chunk's host->device copy with the previous chunk's GEMMs -- and run the recurrence, whose latency chain
 * does not shrink with the batch, once over all utterances (Encoder.encode_from_host). */
int vqcpc_lstm_forward_ex(const vqcpc_encoder_weights* w, const int64_t* idx, int32_t B, int32_t Tp,
                          void* workspace, size_t workspace_bytes, float* out_c, int32_t gemm_mode, void* stream);

/* ------------------------------------------------------------------ Vocoder -- network_vocoder.py:41-78
 * vqcpc_vocoder_pack: weight-only precompute, eprime_out (256, 2688). */
int vqcpc_vocoder_pack(const vqcpc_vocoder_weights* w, float* eprime_out, void* stream);
/* Conditioning: code/speaker embedding, x2 nearest, concat (network_vocoder.py:73-77), prenet biGRU, and the
 * hoisted input projection  G[b,f,:] = p[b,f,:] . ar_w_ih[:, 256:]^T + ar_b_ih   (B, 2Tc, 2688).
 * The x160 upsample of rnnms is the index map t -> t / 160 and is never materialised.
 * Optional out_p (B, 2Tc, 256): the prenet output. */
size_t vqcpc_vocoder_workspace_bytes(int32_t B, int32_t Tc);
int vqcpc_vocoder_condition(const vqcpc_vocoder_weights* w, const int64_t* codes, const int64_t* speaker,
                            int32_t B, int32_t Tc, void* workspace, size_t workspace_bytes,
                            float* out_G, float* out_p, void* stream);
/* Ragged batches (the reference sidesteps them with batch 1, vocoder.py:69, datamodule.py:93):
 * codes (B, Tc) padded to the longest utterance, code_lengths (B,) int32 device array with 1 <= code_lengths[b] <= Tc.
 * Utterance b's bidirectional prenet runs over its own 2 * code_lengths[b] frames, so its first 320 * code_lengths[b]
 * samples are exactly those of an unpadded run; the padded tail of out_G / out_p is zero. */
int vqcpc_vocoder_condition_ragged(const vqcpc_vocoder_weights* w, const int64_t* codes, const int64_t* speaker,
                                   const int32_t* code_lengths, int32_t B, int32_t Tc, void* workspace,
                                   size_t workspace_bytes, float* out_G, float* out_p, void* stream);
/* Vocoder.generate (network_vocoder.py:69-78): one persistent kernel launch per utterance, L <= 320*Tc steps.
 * uniforms (B, L) in [0,1): the injected per-(utterance, step) randomness of the inverse-CDF sampler.
 * out_wav (B, L) fp32; nullable out_codes (B, L) int32 mu-law codes; nullable out_logits (B, L, 256). */
int vqcpc_vocoder_generate(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms,
                           int32_t B, int32_t T2, int32_t L, void* workspace, size_t workspace_bytes,
                           float* out_wav, int32_t* out_codes, float* out_logits, void* stream);
/* Vocoder.forward (network_vocoder.py:41-67; contract vocoder.py:62-63): teacher-forced energies.
 * x_in (B, L) int64 mu-law codes (AR input at step t is x_in[:, t]) -> out_logits (B, L, 256). */
int vqcpc_vocoder_logits_tf(const vqcpc_vocoder_weights* w, const float* G, const int64_t* x_in,
                            int32_t B, int32_t T2, int32_t L, void* workspace, size_t workspace_bytes,
                            float* out_logits, void* stream);
/* Diagnostics: while device_buf is non-NULL, sample-loop launches record 8 clock64() phase timestamps per step of
 * CTA `cta` for steps [first_step, first_step + n_steps) into device_buf[n_steps][8] (see profiles/). */
int vqcpc_debug_set_ar_trace(long long* device_buf, int32_t cta, int32_t first_step, int32_t n_steps);
/* Tuning of the sample loop's exchanges: bits 0..11 cycles before the first poll round (default 400), bits 12..23
 * cycles between failed rounds (default 0), bits 24..27 cap on utterances interleaved per launch (0 = default 4),
 * bit 28 disables the batched (B >= 8) kernel, bit 29 its two-group (65..128 utterances per launch) variant. */
int vqcpc_debug_set_ar_poll_gap(int32_t packed);
/* The single-utterance sample loop runs by default on the cluster kernel (csrc/vocoder_cluster.cu: 7 clusters of 16 CTAs,
 * one grid-scope exchange per step, fc1 / fc2 / sampling over DSMEM) whenever the device can co-schedule that grid.
 * enable = 0 forces the round-1 128-CTA kernel (three grid-scope exchanges per step) for A/B measurements;
 * first_poll_delay = cycles between a CTA's own publish of h_t and its first L2 poll; poll_mode is reserved (0): probes in flight beyond one were measured slower.
 * The cluster kernel's trace (vqcpc_debug_set_ar_trace) has 32 slots per step instead of 8. */
int vqcpc_debug_set_ar_cluster(int32_t enable, int32_t first_poll_delay, int32_t poll_mode);
/* Which kernel single-utterance generate uses on the current device: 1 = cluster kernel, 0 = round-1 128-CTA kernel.  The
 * latter is chosen when the device cannot co-schedule 7 clusters of 16 CTAs, when VQCPC_AR_CLUSTER=0 is set in the environment,
 * and when Nsight Compute is attached to the process (ncu cannot launch the cluster grid -- a cooperative launch that takes
 * every 16-CTA cluster slot of the device -- in any replay mode; a profiled run therefore shows ar_kernel).  The two kernels agree to
 * summation-order noise (logits within 1e-5, tests/test_gpu_parity.py); both are held to the oracle. */
int vqcpc_ar_cluster_active(void);
/* Measures the bare grid-scope exchange of the sample loop that generate runs on this device, with no compute in between:
 * mean SM cycles per exchange over `iters` exchanges.  Cluster kernel (default): 112 CTAs publish 8 LL words each and poll
 * all 896 with the kernel's own first-probe delay -- ONE of these per step (floor = t_smem + one exchange).
 * Round-1 kernel (vqcpc_debug_set_ar_cluster(0, ..)): the 128-way padded-slot exchange, three per step.  workspace >= 64 KiB. */
int vqcpc_debug_exchange_floor(void* workspace, size_t workspace_bytes, int32_t iters, double* mean_cycles, void* stream);
/* ---- log-mel front-end: the step before Encoder.encode.
 * Replaces wave_to_mel, /root/reference/preprocess.py:53-75 (the same arithmetic is inline in convert.py:54-70):
 * peak scaling to 0.999, pre-emphasis lfilter([1, -preemph], [1]), |STFT| (hann(win_length) centred in n_fft, hop_length,
 * center=True with reflect padding), mel projection, 20 log10 with amin 1e-5, per-utterance top_db clamp, / top_db + 1.
 * wave (B, N) fp32, nullable lengths (B,) int32 = samples per utterance of a padded batch (each > n_fft / 2);
 * window (win_length,), dft (2 * n_freq_padded, win_length) = [cos rows | sin rows] of 2 pi k j / n_fft (padding rows
 * zero), melw (n_mels, n_freq_padded) = the mel filterbank (padding columns zero) -- all fp32 device arrays built once
 * by the host (vectorquantizedcpc_b200/frontend.py).  out (B, n_mels, T) fp32, T = 1 + N / hop_length: the layout
 * Encoder.encode takes; frames beyond 1 + lengths[b] / hop_length are 0. */
typedef struct {
    int32_t n_fft, win_length, hop_length, n_mels;
    int32_t n_freq_padded;   /* n_fft / 2 + 1 rounded up to a multiple of 16 */
    float preemph, top_db;
} vqcpc_logmel_config;
size_t vqcpc_logmel_workspace_bytes(const vqcpc_logmel_config* cfg, int32_t B, int32_t N);
int vqcpc_logmel_forward(const vqcpc_logmel_config* cfg, const float* wave, const int32_t* lengths, int32_t B, int32_t N,
                         const float* window, const float* dft, const float* melw, void* workspace, size_t workspace_bytes,
                         float* out, void* stream);
/* ---- text dump of encode.py: np.savetxt(file, z, fmt="%.16f") of encode.py:48-52,57-67 on the GPU.
 * x (rows, cols) fp32 -> the exact bytes numpy writes: every value as "%.16f" % float(v) (the finite decimal expansion of the
 * fp32 value rounded to 16 fractional digits, ties to even; "nan", "inf", "-inf"), values of a row separated by ' ', every
 * row ended by '\n'.  out_text == NULL: only *out_len (bytes) is computed; otherwise out_capacity >= *out_len is required
 * (VQCPC_ERR_ARG with *out_len set if not).  One host synchronisation per call (the length). */
size_t vqcpc_textdump_workspace_bytes(int64_t rows, int32_t cols);
int vqcpc_textdump_f16(const float* x, int64_t rows, int32_t cols, unsigned char* out_text, size_t out_capacity,
                       int64_t* out_len, void* workspace, size_t workspace_bytes, void* stream);

/* ---- output stage of convert.py: pyloudnorm's integrated loudness and gain.
 * Replaces pyloudnorm.Meter(sr).integrated_loudness (convert.py:50,57,79) and pyloudnorm.normalize.loudness
 * (convert.py:80): K-weighting biquads for `rate`, 400 ms blocks / 75 % overlap, -70 LUFS absolute and -10 LU relative
 * gates.  wave (B, N) fp32 mono, nullable lengths (B,) int32; out_lufs (B,) fp32 (-inf for silence, as pyloudnorm).
 * vqcpc_loudness_normalize writes out_wave = wave * 10^((target_lufs[b] - measured[b]) / 20) (0 beyond lengths[b]). */
size_t vqcpc_loudness_workspace_bytes(int32_t B, int32_t N, int32_t rate);
int vqcpc_integrated_loudness(const float* wave, const int32_t* lengths, int32_t B, int32_t N, int32_t rate, void* workspace,
                              size_t workspace_bytes, float* out_lufs, void* stream);
int vqcpc_loudness_normalize(const float* wave, const int32_t* lengths, const float* target_lufs, int32_t B, int32_t N, int32_t rate,
                             void* workspace, size_t workspace_bytes, float* out_wave, float* out_measured_lufs, void* stream);
/* Reads (and clears) the device-side status word of the persistent kernels in `workspace` after the stream
 * has been synchronised by the caller: 0 ok, VQCPC_ERR_TIMEOUT if an exchange timed out. */
int vqcpc_check_status(void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQCPC_H */
