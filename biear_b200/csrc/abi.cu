// C-ABI plumbing of libbiear_b200.so: version, thread-local error text, launch accounting and the
// per-device FFT twiddle table.  No compute lives here.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace biear {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail_invalid(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return BIEAR_EINVAL;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// One table per device, built in float64 on the host and uploaded synchronously the first time a
// device is used.  (First use must therefore happen outside CUDA-graph capture: biear_init does it.)
const float2* twiddle_table(cudaStream_t, int* err) {
    static std::mutex mu;
    static float2* tables[64] = {nullptr};
    *err = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= 64) {
        *err = check_cuda(e == cudaSuccess ? cudaErrorInvalidDevice : e, "cudaGetDevice");
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(mu);
    if (tables[dev]) return tables[dev];
    std::vector<float2> host(1024);
    for (int m = 0; m < 1024; ++m) {
        const double a = -2.0 * M_PI * (double)m / 1024.0;
        host[m] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    float2* d = nullptr;
    if ((*err = check_cuda(cudaMalloc(&d, sizeof(float2) * 1024), "cudaMalloc(twiddles)"))) return nullptr;
    if ((*err = check_cuda(cudaMemcpy(d, host.data(), sizeof(float2) * 1024, cudaMemcpyHostToDevice),
                           "cudaMemcpy(twiddles)"))) {
        cudaFree(d);
        return nullptr;
    }
    tables[dev] = d;
    return d;
}

}  // namespace biear

extern "C" {

int biear_abi_version(void) { return BIEAR_ABI_VERSION; }

const char* biear_last_error(void) { return biear::g_err; }

int64_t biear_launch_count(void) { return biear::g_launches.load(std::memory_order_relaxed); }

void biear_reset_launch_count(void) { biear::g_launches.store(0, std::memory_order_relaxed); }

int biear_init(void) {
    int err = 0;
    biear::twiddle_table(nullptr, &err);
    return err;
}

}  // extern "C"
