// Broadband interaural cross-correlation over a small lag range, as device-side building blocks.
//
// c[k] = sum_n (L[n+k] - mean L)(R[n] - mean R),  k = k_min..k_max  (np.correlate(l, r, "full")
// cropped to the lags with abs(k)/fs <= max_lag), then c /= max abs(c) + 1e-8 and np.interp onto num_lags points.
// Replaces utils.py:390-420 (compute_cross_correlation_feature) of the reference, which evaluates all
// 2N-1 lags in float64 and throws away all but ~97 of them.
//
// Work split inside one CTA (one clip): a thread owns one block of 16 consecutive lags and a set of
// 16-sample time blocks; per (time block, lag block) it keeps a 31-sample window of L and 16 samples
// of R in registers and issues 256 FMAs for 12 128-bit shared loads.  The clip is streamed through
// shared memory in chunks; the chunk buffers hold the de-meaned, zero-padded signals in a float4
// layout swizzled so that threads working on neighbouring time blocks hit different banks.
//
// HOST_DEVICE so that tests/host_emu/cc_emu.cpp can run the same index arithmetic on the CPU.
#pragma once
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define BIEAR_HD __host__ __device__ __forceinline__
#else
#define BIEAR_HD inline
#endif

namespace biear {

constexpr int kCcThreads = 256;
constexpr int kCcLagBlock = 16;
constexpr int kCcMaxLagBlocks = 16;            // up to 256 lags
constexpr int kCcMaxChunk = 3072;              // samples of R per chunk (upper bound)

struct CcPlan {
    int nlags;        // k_max - k_min + 1
    int lag_blocks;   // ceil(nlags / 16)
    int strips;       // threads per lag block
    int m;            // time blocks per strip per chunk
    int chunk;        // 16 * strips * m samples of R per chunk
    int n_chunks;
};

inline CcPlan cc_make_plan(long long nsamp, int k_min, int k_max) {
    CcPlan p;
    p.nlags = k_max - k_min + 1;
    p.lag_blocks = (p.nlags + kCcLagBlock - 1) / kCcLagBlock;
    p.strips = kCcThreads / p.lag_blocks;
    long long best_total = -1;
    p.m = 1;
    for (int m = 1; m <= 16; ++m) {
        const long long chunk = 16LL * p.strips * m;
        if (chunk > kCcMaxChunk && m > 1) break;
        const long long total = (nsamp + chunk - 1) / chunk * chunk;
        if (best_total < 0 || total < best_total || (total == best_total)) {
            best_total = total;
            p.m = m;
        }
    }
    p.chunk = 16 * p.strips * p.m;
    p.n_chunks = (int)((nsamp + p.chunk - 1) / p.chunk);
    if (p.n_chunks < 1) p.n_chunks = 1;
    return p;
}

// physical float4 slot of logical float4 index g (4 float4 per 16-sample block)
BIEAR_HD int cc_slot4(int g) {
    const int b = g >> 2;
    return (b << 2) + ((g & 3) ^ ((b >> 1) & 3));
}
// physical float index of logical sample index e inside a chunk buffer
BIEAR_HD int cc_slot(int e) { return (cc_slot4(e >> 2) << 2) + (e & 3); }

BIEAR_HD int cc_r_floats(const CcPlan& p) { return p.chunk; }
BIEAR_HD int cc_l_floats(const CcPlan& p) { return p.chunk + kCcLagBlock * p.lag_blocks; }

// 16x16 MACs of one (time block tb, lag block lb) unit.
BIEAR_HD void cc_unit(const float4* sL, const float4* sR, int tb, int lb, float acc[16]) {
    float r[16], w[32];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = sR[cc_slot4(tb * 4 + q)];
        r[4 * q + 0] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 v = sL[cc_slot4((tb + lb) * 4 + q)];
        w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int t = 0; t < 16; ++t)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(w[t + i], r[t], acc[i]);
}

}  // namespace biear
