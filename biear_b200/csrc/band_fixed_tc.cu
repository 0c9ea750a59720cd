// Fixed-Q band stage on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate through the 3xTF32 split.
//
// Same contraction as band_fixed.cu (model_torch.py:451-487, :161-195 and :1039-1063 for fixed Q):
//     Y[m][n] = sum_k abs(X[m][k]) W[n][k],   Z[m][n] = sum_k X[m][k] W[n][k],   phase = atan2(Im Z, Re Z)
// i.e. three GEMMs {abs X, Re X, Im X} (M x F) . W^T (F x N) that share the B operand.  TF32 alone (10-bit mantissa)
// cannot hold the 1e-4 parity contract, so both operands are split x = hi + lo (hi = the TF32-representable top bits,
// lo = x - hi exactly) and every product is formed as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (the
// dropped lo*lo term is ~2^-22 relative).
//
//   fixed_weights_tc_kernel : W (hi and lo parts) in the exact shared-memory image of a B tile, one 16 KB block per
//                             16-bin k-block, so that a stage's B operand is ONE bulk async copy (cp.async.bulk)
//   band_fixed_tc_kernel    : one CTA = 128 items.  Warps 0-7 build the A operand of every k-block on the fly from the
//                             complex spectrum (X is read once: abs, split, six 128 x 16 TF32 tiles in the UMMA
//                             K-major no-swizzle core-matrix layout), warp 8 issues tcgen05.mma (128 x 128 x 8, kind::tf32,
//                             nine per 8 bins) into three fp32 accumulators in TMEM (384 columns); a 3-stage mbarrier
//                             ring decouples them; the epilogue reads TMEM with tcgen05.ld (nan_to_num / atan2).
#include "band_dev.cuh"
#include "tc_dev.cuh"

namespace biear {

constexpr int kTcM = 128;             // items per CTA (UMMA M)
constexpr int kTcN = 128;             // bands per CTA (UMMA N; N <= 128 zero-padded)
constexpr int kTcStages = 3;
constexpr int kTcProducerWarps = 8;
constexpr int kTcThreads = (kTcProducerWarps + 1) * 32;
constexpr int kTcABytes = 6 * kTcTileBytes;                    // {abs, re, im} x {hi, lo}
constexpr int kTcBBytes = 2 * kTcTileBytes;                    // W {hi, lo}
constexpr int kTcStageBytes = kTcABytes + kTcBBytes;           // 64 KB
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 1024; // + barriers / TMEM pointer (and alignment slack)

__host__ __device__ constexpr int tc_kblocks(int F) { return (F + kTcKB - 1) / kTcKB; }

// sqrt.approx: <= 1 ulp-level error (2^-23 relative), a fraction of the instructions of the IEEE sqrtf sequence
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// W^T image: for k-block kb, part p (0 hi, 1 lo): tile at ((kb * 2 + p) * kTcTileBytes / 4) floats.
__global__ void __launch_bounds__(128) fixed_weights_tc_kernel(const float* __restrict__ fc, const float* __restrict__ Q, int N,
                                                               int F, float df, float cutoff, float* __restrict__ Wimg) {
    const int n = blockIdx.x;                       // one band per CTA; n in [N, 128) writes zeros
    __shared__ float s_part[4];
    float fcn = 1.f, q = 1.f;
    if (n < N) {
        fcn = fc[n];
        q = Q[n];
    }
    const BandParams bp = band_params(fcn, q, df, cutoff, F, n < N);
    const int Kp = tc_kblocks(F) * kTcKB;
    float sum = 0.f;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
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
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        float w = 0.f;
        if (k >= bp.k_lo && k <= bp.k_hi && k < F) {
            const float u = fmaf((float)(k - bp.kc), bp.a, bp.b);
            w = sanitize(ex2_approx(-u * u) * inv);
        }
        const float hi = tf32_hi(w);
        const int kb = k / kTcKB, kk = k % kTcKB;
        float* tile = Wimg + (long long)kb * (kTcBBytes / 4);
        tile[tc_tile_off(n, kk)] = hi;
        tile[kTcTileBytes / 4 + tc_tile_off(n, kk)] = w - hi;
    }
}

struct FixedTcArgs {
    const float* X; long long x_stride;      // floats per item
    const float* Wimg;                        // tc_kblocks(F) x 16 KB
    long long items;
    int N, F;
    float* Y; long long y_stride;
    float* phase; long long p_stride;         // nullable
};

__global__ void __launch_bounds__(kTcThreads, 1) band_fixed_tc_kernel(const FixedTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // stage s: [A: abs.hi abs.lo re.hi re.lo im.hi im.lo][B: W.hi W.lo], every tile 8 KB
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + kTcStages * kTcStageBytes;      // full[3], empty[3], acc_full, then the TMEM pointer
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kTcStages, bar_acc = bars + 16 * kTcStages;
    volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kTcStages * kTcStageBytes + 16 * kTcStages + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool want_phase = a.phase != nullptr;
    const int planes = want_phase ? 3 : 1;
#ifdef BIEAR_TC_ONE_KB               // (timing experiments only)
    const int nkb = 1;
#else
    const int nkb = tc_kblocks(a.F);
#endif
    const long long m0 = (long long)blockIdx.x * kTcM;

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(bar_full + 8 * s, kTcProducerWarps + 1); // one arrival per producer warp + the B copy's expect_tx
            mbar_init(bar_empty + 8 * s, 1);                   // tcgen05.commit
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcProducerWarps) {    // the MMA warp owns the TMEM allocation: 512 columns (3 x 128 fp32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_s)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;

    if (warp < kTcProducerWarps) {
        // ================= producers: A tiles from the spectrum =================
        const int r = tid & (kTcM - 1);              // item row of the tile
        const int half = tid >> 7;                    // bins 8*half .. 8*half+7 of the k-block
        const long long m = m0 + r;
        const bool row_ok = m < a.items;
        const float2* xrow = reinterpret_cast<const float2*>(a.X + (row_ok ? m : 0) * a.x_stride);
        float2 x[8];
        auto fetch = [&](int kb) {
            const int k0 = kb * kTcKB + half * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                x[i] = (row_ok && k0 + i < a.F) ? __ldg(xrow + k0 + i) : make_float2(0.f, 0.f);
        };
        fetch(0);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kTcStages;
            const uint32_t ph = (kb / kTcStages) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);              // the MMAs that read this stage have completed
            unsigned char* stage = smem + s * kTcStageBytes;
            if (tid == 0) {                                     // B operand of this k-block: one 16 KB bulk copy
#ifdef BIEAR_TC_SKIP_BCOPY           // (timing experiments only)
                mbar_arrive(bar_full + 8 * s);
            }
            if (false) {
#endif
                mbar_arrive_expect_tx(bar_full + 8 * s, kTcBBytes);
                bulk_g2s(smem_base + s * kTcStageBytes + kTcABytes, a.Wimg + (long long)kb * (kTcBBytes / 4), kTcBBytes,
                         bar_full + 8 * s);
            }
#ifndef BIEAR_TC_SKIP_PRODUCER     // (timing experiments only)
            float v[3][8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[0][i] = sqrt_approx(fmaf(x[i].x, x[i].x, x[i].y * x[i].y));
                v[1][i] = x[i].x;
                v[2][i] = x[i].y;
            }
            if (kb + 1 < nkb) fetch(kb + 1);                    // next k-block's loads in flight behind the stores
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
                if (pl >= planes) break;
#pragma unroll
                for (int c = 0; c < 2; ++c) {                   // the thread's two 4-bin chunks
                    float hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        hi[j] = tf32_hi(v[pl][c * 4 + j]);
                        lo[j] = v[pl][c * 4 + j] - hi[j];
                    }
                    float* t_hi = reinterpret_cast<float*>(stage + (2 * pl) * kTcTileBytes) + tc_tile_off(r, half * 8 + c * 4);
                    float* t_lo = reinterpret_cast<float*>(stage + (2 * pl + 1) * kTcTileBytes) + tc_tile_off(r, half * 8 + c * 4);
                    *reinterpret_cast<float4*>(t_hi) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(t_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
#endif
            fence_proxy_async();                                // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }
        // ================= epilogue: TMEM -> registers -> HBM =================
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int q = warp & 3, colh = warp >> 2;               // TMEM lane quarter (rows 32q..32q+31), band half
        const long long me = m0 + 32 * q + lane;
        const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
        const bool vec_ok = ((a.y_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.Y) & 15) == 0) &&
                            (!want_phase || (((a.p_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.phase) & 15) == 0)));
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            const int n0 = colh * 64 + cb * 16;
            if (n0 >= a.N) break;                               // warp-uniform
            float y[16], zr[16], zi[16];
            tc_ld16(trow + n0, y);
            if (want_phase) {
                tc_ld16(trow + kTcN + n0, zr);
                tc_ld16(trow + 2 * kTcN + n0, zi);
            }
            if (me < a.items) {
                float ph[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    y[j] = sanitize(y[j]);
                    ph[j] = want_phase ? atan2f(zi[j], zr[j]) : 0.f;
                }
                float* yrow = a.Y + me * a.y_stride + n0;
                float* prow = want_phase ? a.phase + me * a.p_stride + n0 : nullptr;
                if (vec_ok) {       // 128-bit stores: 4 instead of 16 store instructions per 16 bands and plane
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        if (n0 + j + 3 < a.N) {
                            *reinterpret_cast<float4*>(yrow + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                            if (want_phase)
                                *reinterpret_cast<float4*>(prow + j) = make_float4(ph[j], ph[j + 1], ph[j + 2], ph[j + 3]);
                        } else {
#pragma unroll
                            for (int i = j; i < j + 4; ++i) {
                                if (n0 + i < a.N) {
                                    yrow[i] = y[i];
                                    if (want_phase) prow[i] = ph[i];
                                }
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (n0 + j < a.N) {
                            yrow[j] = y[j];
                            if (want_phase) prow[j] = ph[j];
                        }
                    }
                }
            }
        }
    } else {
        // ================= MMA issuer (one elected lane) =================
        // instruction descriptor: D fp32, A/B TF32, both K-major, N = 128, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kTcStages;
            const uint32_t ph = (kb / kTcStages) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_base + s * kTcStageBytes;
                const uint32_t sb = sa + kTcABytes;
#pragma unroll
                for (int j = 0; j < 2; ++j) {                   // the two K = 8 steps of the k-block (2 chunks each)
                    const uint32_t koff = j * 2 * kTcLBO;
                    const uint64_t b_hi = tc_smem_desc(sb + koff), b_lo = tc_smem_desc(sb + kTcTileBytes + koff);
                    for (int pl = 0; pl < planes; ++pl) {
                        const uint64_t a_hi = tc_smem_desc(sa + (2 * pl) * kTcTileBytes + koff);
                        const uint64_t a_lo = tc_smem_desc(sa + (2 * pl + 1) * kTcTileBytes + koff);
                        const uint32_t d = tmem + pl * kTcN;
#ifndef BIEAR_TC_SKIP_MMA          // (timing experiments only)
                        tc_mma_tf32(d, a_lo, b_hi, idesc, (kb | j) != 0);   // small terms first
                        tc_mma_tf32(d, a_hi, b_lo, idesc, 1u);
                        tc_mma_tf32(d, a_hi, b_hi, idesc, 1u);
#endif
                    }
                }
                tc_commit(bar_empty + 8 * s);                   // stage free once these MMAs have read it
                if (kb == nkb - 1) tc_commit(bar_acc);          // accumulators complete
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTcProducerWarps)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace biear

extern "C" int64_t biear_band_fixed_tc_workspace_floats(int F) {
    if (F < 2) return 0;
    return (int64_t)biear::tc_kblocks(F) * (biear::kTcBBytes / 4);
}

extern "C" int biear_band_fixed_fwd_tc(const float* X, int64_t x_stride, const float* Q, const float* fc, int64_t items, int N,
                                       int F, float df, float cutoff, float* Y, int64_t y_stride, float* phase,
                                       int64_t phase_stride, float* workspace, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(items >= 0 && N >= 1 && N <= kTcN && F >= 2 && df > 0.f,
                  "biear_band_fixed_fwd_tc: bad shape items=%lld N=%d (<= %d) F=%d df=%g", (long long)items, N, kTcN, F, (double)df);
    if (items == 0) return 0;
    BIEAR_REQUIRE(X && Q && fc && Y && workspace, "biear_band_fixed_fwd_tc: null pointer");
    BIEAR_REQUIRE((x_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 7) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  "biear_band_fixed_fwd_tc: X must be 8-byte aligned with an even stride, workspace 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    static bool configured[64] = {false};
    int dev = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        if (int e = check_cuda(cudaFuncSetAttribute(band_fixed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes),
                               "cudaFuncSetAttribute(band_fixed_tc_kernel)"))
            return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    fixed_weights_tc_kernel<<<kTcN, 128, 0, st>>>(fc, Q, N, F, df, cutoff, workspace);
    BIEAR_LAUNCH_CHECK("fixed_weights_tc_kernel");
    FixedTcArgs a;
    a.X = X; a.x_stride = x_stride; a.Wimg = workspace; a.items = items; a.N = N; a.F = F;
    a.Y = Y; a.y_stride = y_stride; a.phase = phase; a.p_stride = phase_stride;
    const long long grid = (items + kTcM - 1) / kTcM;
    BIEAR_REQUIRE(grid <= 0x7fffffffLL, "biear_band_fixed_fwd_tc: too many items");
    band_fixed_tc_kernel<<<(unsigned)grid, kTcThreads, kTcSmemBytes, st>>>(a);
    BIEAR_LAUNCH_CHECK("band_fixed_tc_kernel");
    return 0;
}
