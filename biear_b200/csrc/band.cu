// Band stage kernels: Gaussian band weighting (width fc/Q) + row normalisation + contraction with one
// frame's spectrum, producing in ONE pass the band energy Y, the sub-band phase and the exact
// Jacobians dY/dQ, dphase/dQ -- or, in the backward (recompute) form, dL/dQ directly.
//
// Replaces, per (row, frame):
//   model_torch.py:340-346   bw, W = exp(-0.5 u^2), W /= sum, nan_to_num, einsum("bf,bnf->bn")
//   model_torch.py:1050-1060 the second W build + complex einsum + atan2 (_subband_phase_from_X)
//   and what autograd derives from both for dL/dQ (closed form).
//
// The (B,100,513) weight tensor of the reference is never materialised: one CTA owns one (row, frame)
// item, stages its 513-bin spectrum once in shared memory as {1, abs, re, im}, and each group of 8 lanes
// owns one band, walking only the bins with abs(u) <= cutoff.  Reductions are 3-step warp shuffles; every
// output element is written by exactly one lane (no atomics).
#include "band_dev.cuh"

namespace biear {

constexpr int kBandWarps = 5;
constexpr int kBandThreads = kBandWarps * 32;

struct BandArgs {
    const float* X; long long x_stride;       // floats
    const float* Q; long long q_stride;
    const float* fc;
    long long items;
    int N, F;
    float df, cutoff;
    // forward outputs
    float* Y; long long y_stride;
    float* phase; long long phase_stride;
    float* dYdQ; float* dPdQ; long long jac_stride;
    // backward inputs / output
    const float* gY; long long gy_stride;
    const float* gP; long long gp_stride;
    float* dQ; long long dq_stride;
    int accumulate;
};

enum { kBandForward = 0, kBandBackward = 1 };

template <int MODE>
__global__ void __launch_bounds__(kBandThreads) band_kernel(const BandArgs a) {
    extern __shared__ float4 s_spec[];
    const int tile = spec_tile_len(a.F);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int quads = (a.N + 3) >> 2;

    for (long long item = blockIdx.x; item < a.items; item += gridDim.x) {
        const float2* x = reinterpret_cast<const float2*>(a.X + item * a.x_stride);
        __syncthreads();   // previous item's readers are done with the tile
        for (int k = threadIdx.x; k < tile; k += kBandThreads) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < a.F) {
                v = spec_entry(__ldg(x + k));
            }
            s_spec[k] = v;
        }
        __syncthreads();

        const float* q_row = a.Q + item * a.q_stride;
        // quads are ordered by frequency = by cost; deal them to the warps in snake order
        for (int round = 0, q0 = 0; q0 < quads; ++round, q0 += kBandWarps) {
            const int slot = (round & 1) ? (kBandWarps - 1 - warp) : warp;
            const int quad = q0 + slot;
            if (quad >= quads) continue;
            const int n = (quad << 2) + (lane >> 3);
            const bool active = n < a.N;
            const float fc = active ? __ldg(a.fc + n) : 1.0f;
            const float q = active ? q_row[n] : 1.0f;
            const BandParams p = band_params(fc, q, a.df, a.cutoff, a.F, active);
            const BandSums s = band_accumulate(s_spec, a.F, p, lane);
            const BandResult r = band_finish(s);
            if (!active || (lane & 7) != 0) continue;

            const float qe = q + 1e-8f;
            const float kappa = -fc / (qe * qe * p.bw);
            const float dydq = kappa * (r.a2 - r.Yraw * r.m2);
            const float mag2 = r.Zr * r.Zr + r.Zi * r.Zi;
            // Re Z Im dZ - Im Z Re dZ with dZ = kappa (z2 - Z m2): the Z m2 terms cancel exactly
            const float dpdq = mag2 > 0.0f ? kappa * (r.Zr * r.z2i - r.Zi * r.z2r) / mag2 : 0.0f;
            if (MODE == kBandForward) {
                a.Y[item * a.y_stride + n] = r.Y;
                if (a.phase) a.phase[item * a.phase_stride + n] = atan2f(r.Zi, r.Zr);
                if (a.dYdQ) a.dYdQ[item * a.jac_stride + n] = dydq;
                if (a.dPdQ) a.dPdQ[item * a.jac_stride + n] = dpdq;
            } else {
                float d = 0.0f;
                if (a.gY) d = __ldg(a.gY + item * a.gy_stride + n) * dydq;
                if (a.gP) d = fmaf(__ldg(a.gP + item * a.gp_stride + n), dpdq, d);
                float* out = a.dQ + item * a.dq_stride + n;
                *out = a.accumulate ? *out + d : d;
            }
        }
    }
}

static int launch_band(const BandArgs& a, int mode, cudaStream_t st) {
    const size_t smem = sizeof(float4) * spec_tile_len(a.F);
    BIEAR_REQUIRE(smem <= 48 * 1024, "band stage: F=%d too large for the spectrum tile", a.F);
    const long long cap = (long long)kSmCountB200 * 12;
    const int grid = (int)(a.items < cap ? a.items : cap);
    if (mode == kBandForward)
        band_kernel<kBandForward><<<grid, kBandThreads, smem, st>>>(a);
    else
        band_kernel<kBandBackward><<<grid, kBandThreads, smem, st>>>(a);
    BIEAR_LAUNCH_CHECK(mode == kBandForward ? "band_fwd_kernel" : "band_bwd_kernel");
    return 0;
}

}  // namespace biear

extern "C" int biear_band_fwd(const float* X, int64_t x_stride, const float* Q, int64_t q_stride,
                              const float* fc, int64_t items, int N, int F, float df, float cutoff, float* Y,
                              int64_t y_stride, float* phase, int64_t phase_stride, float* dYdQ, float* dPdQ,
                              int64_t jac_stride, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(items >= 0 && N >= 1 && F >= 2 && df > 0.f, "biear_band_fwd: bad shape items=%lld N=%d F=%d df=%g",
                  (long long)items, N, F, (double)df);
    if (items == 0) return 0;
    BIEAR_REQUIRE(X && Q && fc && Y, "biear_band_fwd: null pointer");
    BIEAR_REQUIRE((x_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 7) == 0,
                  "biear_band_fwd: X must be 8-byte aligned with an even stride");
    BandArgs a = {};
    a.X = X; a.x_stride = x_stride; a.Q = Q; a.q_stride = q_stride; a.fc = fc;
    a.items = items; a.N = N; a.F = F; a.df = df; a.cutoff = cutoff;
    a.Y = Y; a.y_stride = y_stride; a.phase = phase; a.phase_stride = phase_stride;
    a.dYdQ = dYdQ; a.dPdQ = dPdQ; a.jac_stride = jac_stride;
    return launch_band(a, kBandForward, as_stream(stream));
}

extern "C" int biear_band_bwd(const float* X, int64_t x_stride, const float* Q, int64_t q_stride,
                              const float* fc, int64_t items, int N, int F, float df, float cutoff,
                              const float* gY, int64_t gy_stride, const float* gP, int64_t gp_stride, float* dQ,
                              int64_t dq_stride, int accumulate, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(items >= 0 && N >= 1 && F >= 2 && df > 0.f, "biear_band_bwd: bad shape items=%lld N=%d F=%d df=%g",
                  (long long)items, N, F, (double)df);
    if (items == 0) return 0;
    BIEAR_REQUIRE(X && Q && fc && dQ && (gY || gP), "biear_band_bwd: null pointer");
    BIEAR_REQUIRE((x_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 7) == 0,
                  "biear_band_bwd: X must be 8-byte aligned with an even stride");
    BandArgs a = {};
    a.X = X; a.x_stride = x_stride; a.Q = Q; a.q_stride = q_stride; a.fc = fc;
    a.items = items; a.N = N; a.F = F; a.df = df; a.cutoff = cutoff;
    a.gY = gY; a.gy_stride = gy_stride; a.gP = gP; a.gp_stride = gp_stride;
    a.dQ = dQ; a.dq_stride = dq_stride; a.accumulate = accumulate;
    return launch_band(a, kBandBackward, as_stream(stream));
}
