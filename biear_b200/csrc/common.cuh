// Shared host-side plumbing for the C ABI: error reporting, launch accounting, device constants.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/biear_b200.h"

namespace biear {

constexpr int kSmCountB200 = 148;

// thread-local error text behind biear_last_error()
void set_error(const char* fmt, ...);
int fail_invalid(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n = 1);

// Per-device table tw1024[m] = exp(-2 pi i m / 1024), uploaded on first use (float64 -> float32).
const float2* twiddle_table(cudaStream_t stream, int* err);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define BIEAR_REQUIRE(cond, ...)                      \
    do {                                              \
        if (!(cond)) return biear::fail_invalid(__VA_ARGS__); \
    } while (0)

#define BIEAR_LAUNCH_CHECK(what)                                        \
    do {                                                                \
        biear::count_launch();                                          \
        int _e = biear::check_cuda(cudaGetLastError(), what);          \
        if (_e) return _e;                                              \
    } while (0)

__device__ __forceinline__ float warp_sum_8(float v) {   // sum over the 8 lanes sharing lane/8
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef BIEAR_EXACT_EXP   // diagnostic builds only
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

__device__ __forceinline__ float sanitize(float v) {   // torch.nan_to_num(v, 0, 0, 0)
    return (fabsf(v) <= 3.402823466e+38f) ? v : 0.0f;
}

}  // namespace biear
