// Broadband interaural cross-correlation feature: one CTA per clip, everything (means, 97-lag
// correlation, max-abs normalisation, np.interp resampling) in one launch.
// Replaces utils.py:390-420 (compute_cross_correlation_feature) of the reference.
#include "common.cuh"
#include "cc_dev.cuh"

namespace biear {

struct CcArgs {
    const float* wavL;
    const float* wavR;
    long long B, nsamp, row_stride;
    int k_min;
    CcPlan plan;
    const int32_t* interp_idx;
    const float* interp_frac;
    int num_lags;
    float* cc;
};

__device__ __forceinline__ float block_sum(float v, float* s_red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kCcThreads / 32; ++w) t += s_red[w];
    return t;
}

__global__ void __launch_bounds__(kCcThreads) cc_fwd_kernel(const CcArgs a) {
    extern __shared__ float4 s_dyn[];
    __shared__ float s_red[kCcThreads / 32];
    __shared__ float s_c[kCcLagBlock * kCcMaxLagBlocks];

    const CcPlan& p = a.plan;
    float* sR = reinterpret_cast<float*>(s_dyn);
    float* sL = sR + cc_r_floats(p);
    const int tid = threadIdx.x;

    for (long long clip = blockIdx.x; clip < a.B; clip += gridDim.x) {
        const float* L = a.wavL + clip * a.row_stride;
        const float* R = a.wavR + clip * a.row_stride;

        // ---- means ------------------------------------------------------------------------
        float sl = 0.f, sr = 0.f;
        for (long long i = tid; i < a.nsamp; i += kCcThreads) {
            sl += __ldg(L + i);
            sr += __ldg(R + i);
        }
        const float inv_n = a.nsamp > 0 ? 1.0f / (float)a.nsamp : 0.f;
        const float mean_l = block_sum(sl, s_red) * inv_n;
        const float mean_r = block_sum(sr, s_red) * inv_n;

        // ---- chunked correlation --------------------------------------------------------------
        const int lb = tid / p.strips;
        const int s = tid - lb * p.strips;
        const bool worker = lb < p.lag_blocks;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;

        for (int c = 0; c < p.n_chunks; ++c) {
            const long long c0 = (long long)c * p.chunk;
            __syncthreads();   // previous chunk fully consumed
            for (int e = tid; e < cc_r_floats(p); e += kCcThreads) {
                const long long n = c0 + e;
                sR[cc_slot(e)] = (n < a.nsamp) ? __ldg(R + n) - mean_r : 0.f;
            }
            for (int e = tid; e < cc_l_floats(p); e += kCcThreads) {
                const long long n = c0 + a.k_min + e;
                sL[cc_slot(e)] = (n >= 0 && n < a.nsamp) ? __ldg(L + n) - mean_l : 0.f;
            }
            __syncthreads();
            if (worker) {
                for (int i = 0; i < p.m; ++i)
                    cc_unit(reinterpret_cast<const float4*>(sL), reinterpret_cast<const float4*>(sR),
                            s + p.strips * i, lb, acc);
            }
        }

        // ---- reduce the strips' partial sums: partial[s][lag] lives in the (now free) chunk buffers ----
        __syncthreads();
        float* partial = reinterpret_cast<float*>(s_dyn);
        const int lag_span = kCcLagBlock * p.lag_blocks;
        if (worker) {
#pragma unroll
            for (int i = 0; i < 16; ++i) partial[s * lag_span + lb * kCcLagBlock + i] = acc[i];
        }
        __syncthreads();
        float ck = 0.f;
        if (tid < p.nlags) {
            for (int ss = 0; ss < p.strips; ++ss) ck += partial[ss * lag_span + tid];
        }
        // ---- normalise by max abs(c) + 1e-8 ---------------------------------------------------------
        float mx = warp_max(tid < p.nlags ? fabsf(ck) : 0.f);
        __syncthreads();
        if ((tid & 31) == 0) s_red[tid >> 5] = mx;
        __syncthreads();
        mx = 0.f;
#pragma unroll
        for (int w = 0; w < kCcThreads / 32; ++w) mx = fmaxf(mx, s_red[w]);
        if (tid < p.nlags) s_c[tid] = ck / (mx + 1e-8f);
        __syncthreads();
        // ---- np.interp onto num_lags points -------------------------------------------------------
        for (int j = tid; j < a.num_lags; j += kCcThreads) {
            const int i0 = a.interp_idx[j];
            const float fr = a.interp_frac[j];
            const float c0v = s_c[i0];
            const float c1v = s_c[min(i0 + 1, p.nlags - 1)];
            a.cc[clip * a.num_lags + j] = fmaf(fr, c1v - c0v, c0v);
        }
    }
}

}  // namespace biear

extern "C" int biear_cc_fwd(const float* wavL, const float* wavR, int64_t B, int64_t nsamp, int64_t row_stride,
                            int k_min, int k_max, const int32_t* interp_idx, const float* interp_frac,
                            int num_lags, float* cc, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(B >= 0 && nsamp >= 1 && row_stride >= nsamp && num_lags >= 1,
                  "biear_cc_fwd: bad shape B=%lld nsamp=%lld stride=%lld num_lags=%d", (long long)B,
                  (long long)nsamp, (long long)row_stride, num_lags);
    BIEAR_REQUIRE(k_max >= k_min && k_max - k_min + 1 <= kCcLagBlock * kCcMaxLagBlocks,
                  "biear_cc_fwd: lag range [%d,%d] unsupported (at most %d lags)", k_min, k_max,
                  kCcLagBlock * kCcMaxLagBlocks);
    if (B == 0) return 0;
    BIEAR_REQUIRE(wavL && wavR && interp_idx && interp_frac && cc, "biear_cc_fwd: null pointer");
    CcArgs a;
    a.wavL = wavL; a.wavR = wavR; a.B = B; a.nsamp = nsamp; a.row_stride = row_stride;
    a.k_min = k_min;
    a.plan = cc_make_plan(nsamp, k_min, k_max);
    a.interp_idx = interp_idx; a.interp_frac = interp_frac; a.num_lags = num_lags; a.cc = cc;
    size_t smem = sizeof(float) * (size_t)(cc_r_floats(a.plan) + cc_l_floats(a.plan));
    const size_t partial = sizeof(float) * (size_t)a.plan.strips * kCcLagBlock * a.plan.lag_blocks;
    if (partial > smem) smem = partial;
    BIEAR_REQUIRE(smem <= 48 * 1024, "biear_cc_fwd: internal plan needs %zu B of shared memory", smem);
    const long long cap = (long long)kSmCountB200 * 4;
    const int grid = (int)(B < cap ? B : cap);
    cc_fwd_kernel<<<grid, kCcThreads, smem, as_stream(stream)>>>(a);
    BIEAR_LAUNCH_CHECK("cc_fwd_kernel");
    return 0;
}
