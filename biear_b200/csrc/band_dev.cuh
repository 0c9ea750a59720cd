// Band stage: Gaussian-in-frequency weighting with width fc/Q, row normalisation and contraction
// against one frame's spectrum, for four adjacent bands (a "quad") per warp.
//
// Replaces, per (row, frame):   model_torch.py:340-346  (W build, normalise, einsum -> Y)
//                               model_torch.py:1050-1060 (second W build, complex einsum -> phase)
// and produces in the same pass the three u^2-moments that make the backward into Q closed-form
// (SURVEY.md Appendix A.3), so that W is never materialised and never rebuilt.
//
// Spectrum tile layout in shared memory: float4 {1, abs(X), Re X, Im X} per bin, all-zero from bin F up to the end of
// the tile (at least 31 zero bins, 16-byte aligned).  Work split inside the warp: LANE = BIN, every lane evaluates all
// FOUR bands of the quad on its bin -- one 128-bit shared-memory load feeds four Gaussian evaluations (the first
// version gave every band its own 8 lanes and re-read the spectrum four times per warp: shared-memory wavefronts were
// the busiest pipe of the recurrence kernel, ncu profiles/r1_*), the arithmetic that is common to two bands runs as
// packed fp32x2 instructions (FFMA2 / FMUL2, sm_100), and the loop has a hoisted trip count, so consecutive iterations
// overlap.  The leading 1 makes {sum G, sum G abs(X)} and {sum G Re X, sum G Im X} two FFMA2 per bin and masks the
// zero padding out of sum G; 32 partial sums per lane are reduced with a transposing butterfly (48 shuffles).
#pragma once
#include "common.cuh"

namespace biear {

// diagnostic builds (-DBIEAR_PHASE_PROF): cycle accounting inside the band stage, thread 0 of block 0 (read by seq.cu)
#ifdef BIEAR_PHASE_PROF
static __device__ unsigned long long g_band_prof[8];
#define BAND_PROF_INIT() long long _bp_last = clock64()
#define BAND_PROF_MARK(i)                                           \
    do {                                                            \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                  \
            const long long _n = clock64();                         \
            g_band_prof[i] += (unsigned long long)(_n - _bp_last);  \
            _bp_last = _n;                                          \
        }                                                           \
    } while (0)
#else
#define BAND_PROF_INIT() do {} while (0)
#define BAND_PROF_MARK(i) do {} while (0)
#endif

constexpr float kHalfLog2e = 0.72134752044448170368f;      // 0.5 * log2(e)
constexpr float kSqrtHalfLog2e = 0.84932180028801904272f;  // sqrt(0.5 * log2(e))
constexpr float kTwoLn2 = 1.38629436111989061883f;         // 1 / (0.5 * log2(e))

// F real bins + at least 31 zero bins (a warp walks the window 32 bins at a time from an unaligned start), multiple of 8
__host__ __device__ constexpr int spec_tile_len(int F) { return (F + 31 + 7) & ~7; }

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

// Accumulate the 8 sums of the four bands of the warp's quad over the union of their bin windows.  `p` holds the
// parameters of band (lane >> 3) of the quad (all 8 lanes of a group pass the same values); on return every lane of a
// group holds the 8 sums of its band.  `spec` is the padded tile.
__device__ __forceinline__ BandSums band_accumulate(const float4* __restrict__ spec, int F, const BandParams& p,
                                                    int lane) {
    constexpr unsigned kFull = 0xffffffffu;
    BAND_PROF_INIT();
    // union window of the 4 bands of this warp
    int k0 = p.k_lo, k1 = p.k_hi;
    k0 = min(k0, __shfl_xor_sync(kFull, k0, 8));
    k1 = max(k1, __shfl_xor_sync(kFull, k1, 8));
    k0 = min(k0, __shfl_xor_sync(kFull, k0, 16));
    k1 = max(k1, __shfl_xor_sync(kFull, k1, 16));
    // all four bands' parameters into every lane, re-centred on ONE integer bin kq so that the lane's bin offset is
    // shared: u_j(k) = a_j (k - kc_j) + b_j = a_j (k - kq) + [a_j (kq - kc_j) + b_j]   (the bracket: one rounding)
    const int kq = __shfl_sync(kFull, p.kc, 0);
    float a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j] = __shfl_sync(kFull, p.a, 8 * j);
        const float bj = __shfl_sync(kFull, p.b, 8 * j);
        const int kcj = __shfl_sync(kFull, p.kc, 8 * j);
        b[j] = fmaf(a[j], (float)(kq - kcj), bj);
    }
    float2 sy[4], zz[4], ma[4], z2[4];       // per band: {S, Y}, {Zr, Zi}, {m2, a2}, {z2r, z2i}
#pragma unroll
    for (int j = 0; j < 4; ++j) sy[j] = zz[j] = ma[j] = z2[j] = make_float2(0.f, 0.f);

    BAND_PROF_MARK(0);                        // window union + parameter shuffles
    const int n_it = (k1 - k0 + 32) >> 5;     // <= 0 for an empty window (k_lo = F, k_hi = -1)
    int k = k0 + lane;                        // k <= k1 + 31 <= F + 30 < spec_tile_len(F): zeros beyond F - 1
    float kf = (float)(k - kq);
#pragma unroll 2
    for (int it = 0; it < n_it; ++it, k += 32, kf += 32.0f) {
        const float4 x = spec[k];
        const float2 oa = make_float2(x.x, x.y), ri = make_float2(x.z, x.w);
        const float2 kk = make_float2(kf, kf);
#pragma unroll
        for (int h = 0; h < 2; ++h) {         // two bands at a time through the packed pipes
            const float2 u = __ffma2_rn(make_float2(a[2 * h], a[2 * h + 1]), kk, make_float2(b[2 * h], b[2 * h + 1]));
            const float2 e = __fmul2_rn(u, u);
            const float2 g = make_float2(ex2_approx(-e.x), ex2_approx(-e.y));
            const float2 ge = __fmul2_rn(g, e);
            const float2 g0 = make_float2(g.x, g.x), g1 = make_float2(g.y, g.y);
            const float2 ge0 = make_float2(ge.x, ge.x), ge1 = make_float2(ge.y, ge.y);
            sy[2 * h] = __ffma2_rn(oa, g0, sy[2 * h]);
            zz[2 * h] = __ffma2_rn(ri, g0, zz[2 * h]);
            ma[2 * h] = __ffma2_rn(oa, ge0, ma[2 * h]);
            z2[2 * h] = __ffma2_rn(ri, ge0, z2[2 * h]);
            sy[2 * h + 1] = __ffma2_rn(oa, g1, sy[2 * h + 1]);
            zz[2 * h + 1] = __ffma2_rn(ri, g1, zz[2 * h + 1]);
            ma[2 * h + 1] = __ffma2_rn(oa, ge1, ma[2 * h + 1]);
            z2[2 * h + 1] = __ffma2_rn(ri, ge1, z2[2 * h + 1]);
        }
    }
    BAND_PROF_MARK(1);                        // the bin loop
#ifdef BIEAR_PHASE_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0) g_band_prof[4] += (unsigned long long)max(n_it, 0);
#endif
    // Transposing reduction: 4 bands x 8 sums per lane -> the 8 sums of band (lane >> 3) in every lane of its group.
    //   round 1 (xor 16): lanes 0-15 keep bands 0,1 and hand bands 2,3 over, lanes 16-31 the other way round;
    //   round 2 (xor 8) : keep band (lane >> 3); then a 3-step butterfly over the 8 lanes of the group.
    float v[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j][0] = sy[j].x; v[j][1] = sy[j].y; v[j][2] = zz[j].x; v[j][3] = zz[j].y;
        v[j][4] = ma[j].x; v[j][5] = ma[j].y; v[j][6] = z2[j].x; v[j][7] = z2[j].y;
    }
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    float w[2][8];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float give = up16 ? v[j][i] : v[j + 2][i];
            const float keep = up16 ? v[j + 2][i] : v[j][i];
            w[j][i] = keep + __shfl_xor_sync(kFull, give, 16);
        }
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float give = up8 ? w[0][i] : w[1][i];
        const float keep = up8 ? w[1][i] : w[0][i];
        t[i] = warp_sum_8(keep + __shfl_xor_sync(kFull, give, 8));
    }
    BandSums s;
    s.S = t[0]; s.Y = t[1]; s.Zr = t[2]; s.Zi = t[3];
    s.m2 = t[4]; s.a2 = t[5]; s.z2r = t[6]; s.z2i = t[7];
    BAND_PROF_MARK(2);                        // transposing reduction
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
