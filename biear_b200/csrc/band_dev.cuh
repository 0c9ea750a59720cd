// Band stage: Gaussian-in-frequency weighting with width fc/Q, row normalisation and contraction
// against one frame's spectrum, for four adjacent bands per warp (8 lanes per band).
//
// Replaces, per (row, frame):   model_torch.py:340-346  (W build, normalise, einsum -> Y)
//                               model_torch.py:1050-1060 (second W build, complex einsum -> phase)
// and produces in the same pass the three u^2-moments that make the backward into Q closed-form
// (SURVEY.md Appendix A.3), so that W is never materialised and never rebuilt.
//
// Spectrum tile layout in shared memory: float4 {1, abs(X), Re X, Im X} per bin, all-zero beyond bin F-1 up to
// a multiple of 8 bins, 16-byte aligned.  The 8 lanes of a band read 8 consecutive bins (128 B = one
// conflict-free wavefront); the 4 band groups of the warp read the same addresses (broadcast).  The leading 1
// makes {sum G, sum G abs(X)} and {sum G Re X, sum G Im X} two packed fp32x2 FMAs (FFMA2, sm_100) per bin, and the
// zero padding removes the bin-range select from the loop.
#pragma once
#include "common.cuh"

namespace biear {

constexpr float kHalfLog2e = 0.72134752044448170368f;      // 0.5 * log2(e)
constexpr float kSqrtHalfLog2e = 0.84932180028801904272f;  // sqrt(0.5 * log2(e))
constexpr float kTwoLn2 = 1.38629436111989061883f;         // 1 / (0.5 * log2(e))

__host__ __device__ constexpr int spec_tile_len(int F) { return (F + 7) & ~7; }

struct BandSums {
    float S, Y, Zr, Zi, m2, a2, z2r, z2i;
};

// Per-lane Gaussian parameters of the lane's band.
struct BandParams {
    float bw;      // fc/(Q+1e-8)+1e-8
    float a, b;    // u*sqrt(.5 log2 e) = a*(k-kc) + b  for bin index k (kc = bin nearest fc, so that
                   // a*(k-kc) is an exact product of small integers and b is small: no cancellation)
    int kc;
    int k_lo, k_hi;
};

__device__ __forceinline__ BandParams band_params(float fc, float q, float df, float cutoff, int F, bool active) {
    BandParams p;
    p.bw = fc / (q + 1e-8f) + 1e-8f;
    const float inv = 1.0f / p.bw;
    const float inv_df = 1.0f / df;
    p.kc = min(F - 1, max(0, __float2int_rn(fc * inv_df)));
    p.a = df * inv * kSqrtHalfLog2e;
    p.b = fmaf((float)p.kc, df, -fc) * (inv * kSqrtHalfLog2e);
    if (!active) {
        p.k_lo = F;      // empty window; contributes nothing to the quad's union
        p.k_hi = -1;
    } else if (cutoff > 0.0f) {
        const float half = cutoff * p.bw;
        // floor / ceil so the two bins bracketing fc are always inside; NaN -> 0 (then sanitised)
        p.k_lo = max(0, __float2int_rd((fc - half) * inv_df));
        p.k_hi = min(F - 1, __float2int_ru((fc + half) * inv_df));
        if (!(half == half)) { p.k_lo = 0; p.k_hi = F - 1; }
    } else {
        p.k_lo = 0;
        p.k_hi = F - 1;
    }
    return p;
}

__device__ __forceinline__ float4 spec_entry(float2 c) {   // one bin of the shared spectrum tile
    return make_float4(1.0f, sqrtf(fmaf(c.x, c.x, c.y * c.y)), c.x, c.y);
}

// Accumulate the 8 sums of the lane's band over the quad's bin window.  `spec` is the padded tile.
__device__ __forceinline__ BandSums band_accumulate(const float4* __restrict__ spec, int F, const BandParams& p,
                                                    int lane) {
    const int j = lane & 7;
    // union window of the 4 bands of this warp, aligned to 8 bins
    int k0 = p.k_lo, k1 = p.k_hi;
    k0 = min(k0, __shfl_xor_sync(0xffffffffu, k0, 8));
    k1 = max(k1, __shfl_xor_sync(0xffffffffu, k1, 8));
    k0 = min(k0, __shfl_xor_sync(0xffffffffu, k0, 16));
    k1 = max(k1, __shfl_xor_sync(0xffffffffu, k1, 16));
    k0 &= ~7;

    float2 sy = make_float2(0.f, 0.f), zz = make_float2(0.f, 0.f);     // {S, Y}, {Zr, Zi}
    float2 ma = make_float2(0.f, 0.f), z2 = make_float2(0.f, 0.f);     // {m2, a2}, {z2r, z2i}
    int k = k0 + j;
    float kf = (float)(k - p.kc);
#pragma unroll 8
    for (; k - j <= k1; k += 8, kf += 8.0f) {
        const float4 x = spec[k];                       // k < spec_tile_len(F) always; zeros beyond F-1
        const float u = fmaf(kf, p.a, p.b);
        const float e = u * u;
        const float g = ex2_approx(-e);
        const float ge = g * e;
        const float2 oa = make_float2(x.x, x.y), ri = make_float2(x.z, x.w);
        const float2 gg = make_float2(g, g), gege = make_float2(ge, ge);
        sy = __ffma2_rn(oa, gg, sy);
        zz = __ffma2_rn(ri, gg, zz);
        ma = __ffma2_rn(oa, gege, ma);
        z2 = __ffma2_rn(ri, gege, z2);
    }
    BandSums s;
    s.S = warp_sum_8(sy.x);
    s.Y = warp_sum_8(sy.y);
    s.Zr = warp_sum_8(zz.x);
    s.Zi = warp_sum_8(zz.y);
    s.m2 = warp_sum_8(ma.x);
    s.a2 = warp_sum_8(ma.y);
    s.z2r = warp_sum_8(z2.x);
    s.z2i = warp_sum_8(z2.y);
    return s;
}

// Normalised results of one band (every one of the band's 8 lanes holds the same values).
struct BandResult {
    float Y;             // nan_to_num(sum abs(X) W)
    float Yraw;          // the same before nan_to_num (what the Jacobian is built from)
    float Zr, Zi;        // sum W X
    float m2, a2, z2r, z2i;   // u^2 moments (true u, not the scaled one)
    float S;
};

__device__ __forceinline__ BandResult band_finish(const BandSums& s) {
    BandResult r;
    const float inv_s = 1.0f / (s.S + 1e-8f);
    const float inv_s2 = inv_s * kTwoLn2;
    r.S = s.S;
    r.Yraw = s.Y * inv_s;
    r.Y = sanitize(r.Yraw);
    r.Zr = s.Zr * inv_s;
    r.Zi = s.Zi * inv_s;
    r.m2 = s.m2 * inv_s2;
    r.a2 = s.a2 * inv_s2;
    r.z2r = s.z2r * inv_s2;
    r.z2i = s.z2i * inv_s2;
    return r;
}

}  // namespace biear
