// Internal (non-exported) launchers shared between translation units of libvqcpc_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vqcpc {

// gemm_f32.cu
// split-K for the latency case M <= 256 (optional): splitk_ws = gemm_splitk_ws_bytes(M, N) bytes of scratch planes,
// splitk_counters = SPLITK_COUNTERS zeroed words (self-resetting after each GEMM)
constexpr int SPLITK_COUNTERS = 60;
int gemm_dense(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
               int64_t M, int N, int K, cudaStream_t stream, void* splitk_ws = nullptr, size_t splitk_ws_bytes = 0,
               unsigned* splitk_counters = nullptr, bool pdl = false);      // pdl: programmatic dependent launch (common.cuh)
int gemm_conv(const float* mel, int B, int T, int Cin, const float* W, float* C, int Cout, cudaStream_t stream,
              void* splitk_ws = nullptr, size_t splitk_ws_bytes = 0, unsigned* splitk_counters = nullptr);
size_t gemm_splitk_ws_bytes(int64_t M, int N);

// gemm_tc.cu  (tcgen05 / TMEM / TMA)
int gemm_tc(const void* a_planes, const void* w_planes, const float* bias, float* C, long long ldc, int M, int N, int K,
            int nseg, int* err_flag, cudaStream_t stream);
struct TcPlan {                      // a tcgen05 GEMM bound to fixed operand buffers (tensor maps encoded once)
    alignas(64) unsigned char map_a[128];
    alignas(64) unsigned char map_w[128];
    float* C; const float* bias; int* err; long long ldc; int M, N, K, nseg, bn;
};
int gemm_tc_plan(TcPlan* plan, const void* a_planes, const void* w_planes, const float* bias, float* C, long long ldc, int M,
                 int N, int K, int nseg, int* err_flag);
int gemm_tc_run(const TcPlan* plan, cudaStream_t stream, bool pdl = false);   // pdl: programmatic dependent launch
int gemm_tc_plan_lstm(TcPlan* plan, const void* h_planes, const void* whh_planes, int M, int hidden, int* err_flag);
int gemm_tc_run_lstm(const TcPlan* plan, const float* table, const int64_t* idx, float* cstate, float* out, void* planes_out, int t,
                     int Tp, cudaStream_t stream, bool pdl);
size_t gemm_tc_ln_scratch_bytes(int N);
bool gemm_tc_ln_supported(int N);
int gemm_tc_ln(const void* a_planes, const void* w_planes, const float* ln_w, const float* ln_b, void* out_planes,
               float* out_f32, void* scratch, int M, int N, int K, int nseg, int* err_flag, cudaStream_t stream);
// gemm_pair.cu: the same fused layer on CTA pairs (tcgen05 cta_group::2); nseg = 1 is the single-pass bf16 mode
size_t gemm_ln_pair_scratch_bytes(int N);
bool gemm_ln_pair_supported(int N);
int gemm_ln_pair(const void* a_planes, long long a_pitch, int a_lo_col, const void* w_planes, const float* ln_w, const float* ln_b,
                 void* out_planes, int out_pitch, float* out_f32, void* scratch, int M, int N, int K, int nseg, int* err_flag,
                 cudaStream_t stream);
// lstm_persist.cu: the whole batched LSTM recurrence in one persistent tcgen05 launch (W_hh slices resident in shared memory)
int lstm_persist(const float* table, const int64_t* idx, const void* whh_planes, int B, int Tp, void* planes0, void* planes1,
                 unsigned* counters, float* table_perm, float* out, int* err_flag, cudaStream_t stream);
size_t lstm_persist_table_bytes();
// lstm_cluster.cu: latency LSTM for a few utterances, one 16-CTA cluster each, h_t over DSMEM
int lstm_cluster_supported();
int lstm_cluster_launch(const float* table, const int64_t* idx, const float* w_hh, int B, int Tp, float* out, int* status,
                        cudaStream_t stream);      // launched as a programmatic dependent of the kernel before it
int split_planes(const float* x, long long ld, void* out, long long rows, int K, cudaStream_t stream);

// encoder.cu
int layernorm_relu(float* x, const float* w, const float* b, int64_t rows, int C, cudaStream_t stream, bool pdl = false);
int vq_lookup(const float* x, const float* codebook, int64_t n, int n_codes, int dim, float* q, int64_t* idx,
              cudaStream_t stream, bool pdl = false);
// tensor-core search (vq_tc.cu); vq_lookup_auto picks it for n >= VQ_TC_MIN_FRAMES when scratch is provided
constexpr int64_t VQ_TC_MIN_FRAMES = 8192;
constexpr int VQ_TC_FLAG_CAP = 32768;           // frames the main kernel can list for the exact rescan before it falls back to a sweep
// bf16 hi/lo planes of -2e | |e|^2 fp32 | |e|^2 MMA block | flagged-frame list (count + 3 pad words + VQ_TC_FLAG_CAP frame numbers)
constexpr size_t VQ_TC_PLANES_BYTES = 512 * 128 * 2 + 512 * 4 + 512 * 32 + 16 + 4 * VQ_TC_FLAG_CAP;
int vq_lookup_tc(const float* x, const float* codebook, int64_t n, float* q, int64_t* idx, void* planes_ws, int* err,
                 cudaStream_t stream);
int vq_lookup_auto(const float* x, const float* codebook, int64_t n, float* q, int64_t* idx, void* planes_ws, int* err,
                   cudaStream_t stream, bool pdl = false);

// vocoder_batch.cu: batched sample loop (up to 64 utterances per launch, grid-barrier phases)
size_t ar_batch_workspace_bytes();
int ensure_dyn_smem(const void* func, int bytes);   // per-(kernel, device) opt-in to large dynamic shared memory

// persistent-kernel workspace header (first bytes of every workspace handed to a persistent kernel)
constexpr int INDEX_ERROR_MAGIC = 0x1DE7BAD5;
struct WorkspaceHeader {
    int status;       // 0 ok, else VQCPC_ERR_* (persistent kernels / tensor-core pipelines; reset by the call that launches them)
    int index_error;  // == INDEX_ERROR_MAGIC: a code / speaker index was out of range (set by the conditioning gather, reset by
                      // vocoder_condition).  A magic value, not a flag: entry points that do not run the gather leave this word
                      // alone, and a workspace they are handed may hold uninitialised memory here.
    int reserved[62];
};

void count_launch(int n);
int device_sm_count();
int device_cc_major();

}  // namespace vqcpc
