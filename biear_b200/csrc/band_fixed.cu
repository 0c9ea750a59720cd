// Fixed-Q band stage as a dense contraction: when every (row, frame) item uses the SAME Q vector (FIXED_FRONTEND_Q,
// the AuralNet filterbank, the passive / precompute path), the Gaussian weights W (N x F) are one shared matrix and
//     Y[m][n] = sum_k abs(X[m][k]) W[n][k],      Z[m][n] = sum_k X[m][k] W[n][k],      phase = atan2(Im Z, Re Z)
// is a GEMM of the three planes {abs X, Re X, Im X} (M x F each) against W^T.
//
// Replaces model_torch.py:451-487 (FramewiseFixedGammatoneFB: rebuilds the identical W for each of the 19 frames and
// contracts with a batched mat-vec), :161-195 (AuralNetGammatoneFB's einsum) and :1039-1063 (phase) for fixed Q.
//
//   fixed_weights_kernel : W^T (Fp x 128, zero padded) from (fc, Q): exp2-based Gaussian on the |u| <= cutoff window,
//                          row-normalised with the reference's 1e-8 terms, NaN/Inf -> 0 (nan_to_num)
//   band_fixed_kernel    : 32 items x 128 bands per CTA, K chunks of 16 bins double-buffered in shared memory, the
//                          planes built on the fly from the complex spectrum (X is read exactly once), 4 x 4 x 3
//                          register tiles with packed FFMA2, atan2 / nan_to_num epilogue
// fp32 on purpose: a TF32 tensor-core version needs the 3-way split to hold the 1e-4 parity contract.
#include "band_dev.cuh"

namespace biear {

constexpr int kFxM = 32;            // items per CTA
constexpr int kFxN = 128;           // bands per CTA (N <= 128, padded)
constexpr int kFxK = 16;            // bins per chunk
constexpr int kFxThreads = 256;     // 8 (4 items each) x 32 (4 bands each)
constexpr int kFxAPitch = 3 * kFxM + 4;   // floats per bin row of the plane tile: {abs[32], re[32], im[32]} + pad

__global__ void __launch_bounds__(128) fixed_weights_kernel(const float* __restrict__ fc, const float* __restrict__ Q, int N,
                                                            int F, int Fp, float df, float cutoff, float* __restrict__ Wt) {
    const int n = blockIdx.x;                       // one band per CTA; n in [N, 128) writes zeros
    __shared__ float s_part[4];
    float fcn = 1.f, q = 1.f;
    BandParams bp;
    if (n < N) {
        fcn = fc[n];
        q = Q[n];
    }
    bp = band_params(fcn, q, df, cutoff, F, n < N);
    float sum = 0.f;
    for (int k = threadIdx.x; k < Fp; k += blockDim.x) {
        float g = 0.f;
        if (k >= bp.k_lo && k <= bp.k_hi && k < F) {
            const float u = fmaf((float)(k - bp.kc), bp.a, bp.b);
            g = ex2_approx(-u * u);
        }
        sum += g;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    const float S = s_part[0] + s_part[1] + s_part[2] + s_part[3];
    const float inv = 1.0f / (S + 1e-8f);
    for (int k = threadIdx.x; k < Fp; k += blockDim.x) {
        float w = 0.f;
        if (k >= bp.k_lo && k <= bp.k_hi && k < F) {
            const float u = fmaf((float)(k - bp.kc), bp.a, bp.b);
            w = sanitize(ex2_approx(-u * u) * inv);
        }
        Wt[(long long)k * kFxN + n] = w;
    }
}

struct FixedArgs {
    const float* X; long long x_stride;      // floats per item
    const float* Wt;                          // (Fp, 128)
    long long items;
    int N, F, Fp;
    float* Y; long long y_stride;
    float* phase; long long p_stride;         // nullable
};

__global__ void __launch_bounds__(kFxThreads, 2) band_fixed_kernel(const FixedArgs a) {
    __shared__ __align__(16) float As[2][kFxK * kFxAPitch];
    __shared__ __align__(16) float Ws[2][kFxK * kFxN];
    const int tid = threadIdx.x;
    const int tm = tid >> 5, tn = tid & 31;               // items 4tm..4tm+3, bands 4tn..4tn+3
    const long long m0 = (long long)blockIdx.x * kFxM;
    // loader roles: 32 items x 16 bins = 512 complex values, 2 per thread (consecutive threads -> consecutive bins)
    const int li = tid >> 4, lk = tid & 15;               // items li and li + 16, bin lk of the chunk
    const int chunks = a.Fp / kFxK;
    const bool want_phase = a.phase != nullptr;

    float2 acc[3][4][2];                                  // [plane][item][band pair]
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[pl][i][0] = acc[pl][i][1] = make_float2(0.f, 0.f);

    float2 xa[2];
    float4 wv[2];
    auto fetch = [&](int c) {
        const int k = c * kFxK + lk;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long m = m0 + li + 16 * h;
            xa[h] = (m < a.items && k < a.F) ? __ldg(reinterpret_cast<const float2*>(a.X + m * a.x_stride) + k)
                                             : make_float2(0.f, 0.f);
        }
        // W chunk: 16 x 128 floats = 512 float4, 2 per thread
#pragma unroll
        for (int h = 0; h < 2; ++h)
            wv[h] = __ldg(reinterpret_cast<const float4*>(a.Wt + (long long)c * kFxK * kFxN) + tid + kFxThreads * h);
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float* row = &As[buf][lk * kFxAPitch + li + 16 * h];
            row[0] = sqrtf(fmaf(xa[h].x, xa[h].x, xa[h].y * xa[h].y));
            row[kFxM] = xa[h].x;
            row[2 * kFxM] = xa[h].y;
            reinterpret_cast<float4*>(Ws[buf])[tid + kFxThreads * h] = wv[h];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int c = 0; c < chunks; ++c) {
        if (c + 1 < chunks) fetch(c + 1);
        const float* as = As[buf] + tm * 4;
        const float* ws = Ws[buf] + tn * 4;
#pragma unroll
        for (int kk = 0; kk < kFxK; ++kk) {
            const float4 w = *reinterpret_cast<const float4*>(ws + kk * kFxN);
            const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
                if (pl > 0 && !want_phase) continue;
                const float4 x = *reinterpret_cast<const float4*>(as + kk * kFxAPitch + pl * kFxM);
                const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 xx = make_float2(xs[i], xs[i]);
                    acc[pl][i][0] = __ffma2_rn(xx, w01, acc[pl][i][0]);
                    acc[pl][i][1] = __ffma2_rn(xx, w23, acc[pl][i][1]);
                }
            }
        }
        if (c + 1 < chunks) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + tm * 4 + i;
        if (m >= a.items) continue;
        const float y[4] = {acc[0][i][0].x, acc[0][i][0].y, acc[0][i][1].x, acc[0][i][1].y};
        const float zr[4] = {acc[1][i][0].x, acc[1][i][0].y, acc[1][i][1].x, acc[1][i][1].y};
        const float zi[4] = {acc[2][i][0].x, acc[2][i][0].y, acc[2][i][1].x, acc[2][i][1].y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = tn * 4 + j;
            if (n >= a.N) continue;
            a.Y[m * a.y_stride + n] = sanitize(y[j]);
            if (want_phase) a.phase[m * a.p_stride + n] = atan2f(zi[j], zr[j]);
        }
    }
}

}  // namespace biear

extern "C" int64_t biear_band_fixed_workspace_floats(int F) {
    if (F < 2) return 0;
    const int Fp = (F + biear::kFxK - 1) / biear::kFxK * biear::kFxK;
    return (int64_t)Fp * biear::kFxN;
}

extern "C" int biear_band_fixed_fwd(const float* X, int64_t x_stride, const float* Q, const float* fc, int64_t items, int N,
                                    int F, float df, float cutoff, float* Y, int64_t y_stride, float* phase,
                                    int64_t phase_stride, float* workspace, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(items >= 0 && N >= 1 && N <= kFxN && F >= 2 && df > 0.f,
                  "biear_band_fixed_fwd: bad shape items=%lld N=%d (<= %d) F=%d df=%g", (long long)items, N, kFxN, F, (double)df);
    if (items == 0) return 0;
    BIEAR_REQUIRE(X && Q && fc && Y && workspace, "biear_band_fixed_fwd: null pointer");
    BIEAR_REQUIRE((x_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 7) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  "biear_band_fixed_fwd: X must be 8-byte aligned with an even stride, workspace 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int Fp = (F + kFxK - 1) / kFxK * kFxK;
    fixed_weights_kernel<<<kFxN, 128, 0, st>>>(fc, Q, N, F, Fp, df, cutoff, workspace);
    BIEAR_LAUNCH_CHECK("fixed_weights_kernel");
    FixedArgs a;
    a.X = X; a.x_stride = x_stride; a.Wt = workspace; a.items = items; a.N = N; a.F = F; a.Fp = Fp;
    a.Y = Y; a.y_stride = y_stride; a.phase = phase; a.p_stride = phase_stride;
    const long long grid = (items + kFxM - 1) / kFxM;
    BIEAR_REQUIRE(grid <= 0x7fffffffLL, "biear_band_fixed_fwd: too many items");
    band_fixed_kernel<<<(unsigned)grid, kFxThreads, 0, st>>>(a);
    BIEAR_LAUNCH_CHECK("band_fixed_kernel");
    return 0;
}
