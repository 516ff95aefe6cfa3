// Error reporting, device queries and status read-back for libvqcpc_b200 (C ABI in include/vqcpc.h).
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

static thread_local char g_error[1024] = "";
static unsigned long long g_launches = 0;   // kernels launched by this library (diagnostic, not thread-safe-exact)

void count_launch(int n) { g_launches += static_cast<unsigned long long>(n); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// attribute of the CURRENT device, cached per device (a process may drive several GPUs)
static int cached_attr(cudaDeviceAttr attr, int (&cache)[64]) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64 && cache[dev] != 0) return cache[dev];
    if (cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64) cache[dev] = v;
    return v;
}
int device_sm_count() {
    static int cache[64] = {0};
    return cached_attr(cudaDevAttrMultiProcessorCount, cache);
}
// One-time opt-in to > 48 KB of dynamic shared memory, remembered per (kernel, device): a function attribute is a
// per-device property, so a process that drives several GPUs must set it on each of them.
int ensure_dyn_smem(const void* func, int bytes) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> done;
    int dev = 0;
    VQ_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& e : done)
        if (e.first == func && e.second == dev) return VQCPC_OK;
    VQ_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.emplace_back(func, dev);
    return VQCPC_OK;
}
int device_cc_major() {
    static int cache[64] = {0};
    return cached_attr(cudaDevAttrComputeCapabilityMajor, cache);
}

}  // namespace vqcpc

extern "C" const char* vqcpc_last_error(void) { return vqcpc::g_error; }
extern "C" int vqcpc_abi_version(void) { return VQCPC_ABI_VERSION; }
extern "C" uint64_t vqcpc_launch_count(void) { return vqcpc::g_launches; }

extern "C" int vqcpc_device_check(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) {
        vqcpc::set_error("no CUDA device %d (%s)", device, e == cudaSuccess ? "out of range" : cudaGetErrorString(e));
        return VQCPC_ERR_DEVICE;
    }
    int major = 0, sms = 0, coop = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (major != 10 || sms < 128 || !coop) {
        vqcpc::set_error("device %d is not a B200-class GPU (cc major %d, %d SMs, cooperative launch %d); "
                         "this library is built for sm_100a only", device, major, sms, coop);
        return VQCPC_ERR_DEVICE;
    }
    return VQCPC_OK;
}

extern "C" int vqcpc_check_status(void* workspace, void* stream) {
    if (workspace == nullptr) {
        vqcpc::set_error("check_status: null workspace");
        return VQCPC_ERR_ARG;
    }
    int hdr[2] = {0, 0};            // WorkspaceHeader: status, index_error
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VQ_CUDA(cudaMemcpyAsync(hdr, workspace, sizeof(hdr), cudaMemcpyDeviceToHost, s));
    VQ_CUDA(cudaStreamSynchronize(s));
    int status = hdr[0];
    if (status == 0 && hdr[1] == vqcpc::INDEX_ERROR_MAGIC) status = VQCPC_ERR_INDEX;
    if (status != 0) {
        VQ_CUDA(cudaMemsetAsync(workspace, 0, sizeof(hdr), s));
        vqcpc::set_error("device reported status %d (%s)", status,
                         status == VQCPC_ERR_TIMEOUT ? "an exchange of a persistent kernel timed out"
                         : status == VQCPC_ERR_INDEX ? "code index or speaker id out of range"
                         : status == VQCPC_ERR_ARG ? "argument out of range" : "unknown");
    }
    return status;
}
