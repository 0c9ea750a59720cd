// Weight gradients of the Q controllers from tile-layout operands (see include/biear_b200.h), ALL layers in two
// launches:
//   matrix job    dW[g][o][i] = sum_{k < chunks} sum_{r < R} A[g][k][o][r] * Bm[g][k][i][r],   db[g][o] = sum_{k,r} A[g][k][o][r]
//   diagonal job  dW[g][o]    = sum_{k,r} A[g][k][o][r] * Bm[g][k][o][r]     (LayerNorm gamma; db = LayerNorm beta)
// A = per-sample pre-activation gradients written by the backward recurrence, Bm = the layer inputs saved by the
// forward recurrence; R = tile rows (16 or 32; 32/R consecutive chunks form one 32-sample slab).  Both operands are
// "K-major in chunks", so a 64 x 64 output tile streams [64][32] slabs of each operand straight into shared memory
// with 128-bit loads.  The contraction ((T-1)*B samples) is far longer than the outputs are wide, so every job is
// split along K over the grid; partials go to scratch and ONE second kernel sums them in a fixed order
// (deterministic, no atomics).  fp32 FFMA on purpose: TF32 cannot hold the 1e-4 gradient parity contract.
//
// Replaces what autograd + cuBLAS do for the weight gradients of model_torch.py:256-267 (GRU weight_ih / weight_hh,
// three Linear layers, two LayerNorms) in the reference: ~30 library launches per ear there.
#include "common.cuh"
#include "tc_dev.cuh"

namespace biear {

constexpr int kWgTile = 64;          // output tile (o and i)
constexpr int kWgThreads = 256;      // 16 x 16 threads, 4 x 4 outputs each
constexpr int kWgPitch = 36;         // padded slab row (floats): 16-byte aligned, conflict-free column reads
constexpr int kWgMaxJobs = BIEAR_WGRAD_MAX_JOBS;

struct WgradJobPlan {
    const float* A; long long a_group, a_chunk; int Do;
    const float* B; long long b_group, b_chunk; int Di;      // Di == 0: diagonal job
    long long tile_chunks;                                    // operand chunks (of tile_w samples)
    long long slabs;                                          // 32-sample slabs
    int tiles_i, tiles, splits; long long slabs_per_split;
    long long diag_chunks_per_split;                          // diagonal jobs: operand chunks per split
    int cta_begin;                                            // first CTA of this job in the partial grid
    long long part_off, bpart_off;                            // offsets into scratch (floats); bpart_off < 0: no bias
    long long out_begin;                                      // first element of this job in the reduce index space
    float* dW; float* db;
    long long dw_group, dw_row, db_group;                    // output strides (floats)
    float* dW2; float scale2;                                 // optional scaled second copy (matrix jobs)
};

struct WgradPlan {
    WgradJobPlan job[kWgMaxJobs];
    int n_jobs, G, tile_w;
    int total_ctas;
    long long total_out;
    float* scratch;
};

template <int TILE_W>
__device__ __forceinline__ void wgrad_matrix(const WgradPlan& pl, const WgradJobPlan& a, int tile, int split, int g,
                                             float (*As)[kWgTile * kWgPitch], float (*Bs)[kWgTile * kWgPitch]) {
    static_assert(TILE_W == 16 || TILE_W == 32, "tile rows");
    const int to = tile / a.tiles_i, ti = tile % a.tiles_i;
    const int o0 = to * kWgTile, i0 = ti * kWgTile;
    const long long c0 = (long long)split * a.slabs_per_split;
    const long long c1 = min(a.slabs, c0 + a.slabs_per_split);
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;          // outputs o0 + ty + 16*{0..3}, i0 + tx + 16*{0..3}
    const int lr = tid / 8, lc = tid % 8;            // slab loader: rows lr and lr + 32, float4 column lc
    // this thread's two rows of each operand inside a slab: operand chunk (lc*4)/TILE_W of the slab, offset (lc*4)%TILE_W
    constexpr int kPerSlab = 32 / TILE_W;
    const int sub = (lc * 4) / TILE_W, off = (lc * 4) % TILE_W;
    const float* ap[2];
    const float* bp[2];
    bool aok[2], bok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = lr + 32 * h;
        aok[h] = o0 + r < a.Do;
        bok[h] = i0 + r < a.Di;
        ap[h] = a.A + (long long)g * a.a_group + (long long)(o0 + r) * TILE_W + off;
        bp[h] = a.B + (long long)g * a.b_group + (long long)(i0 + r) * TILE_W + off;
    }

    float2 acc[4][4];                                // {even-sample, odd-sample} partial sums of output (y, x)
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) acc[y][x] = make_float2(0.f, 0.f);
    // bias (sum of A over the samples): lane l of warp w sums 8 samples of row 8w + l/4 per slab; combined at the end
    const bool want_bias = a.bpart_off >= 0 && ti == 0;
    const int brow = (tid >> 5) * 8 + ((tid & 31) >> 2), bcol = (tid & 3) * 8;
    float bsum = 0.f;
    float4 ra[2], rb[2];
    auto fetch = [&](long long c) {
        const long long cc = c * kPerSlab + sub;                             // operand chunk of this float4 column
        const bool live = cc < a.tile_chunks;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            ra[h] = (live && aok[h]) ? __ldg(reinterpret_cast<const float4*>(ap[h] + cc * a.a_chunk))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[h] = (live && bok[h]) ? __ldg(reinterpret_cast<const float4*>(bp[h] + cc * a.b_chunk))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 32 * h;
            *reinterpret_cast<float4*>(&As[buf][r * kWgPitch + lc * 4]) = ra[h];
            *reinterpret_cast<float4*>(&Bs[buf][r * kWgPitch + lc * 4]) = rb[h];
        }
    };
    if (c0 < c1) {
        fetch(c0);
        stash(0);
    }
    __syncthreads();
    int buf = 0;
    for (long long c = c0; c < c1; ++c) {
        if (c + 1 < c1) fetch(c + 1);
        const float* as = As[buf] + ty * kWgPitch;      // rows ty + 16 j: the 8 lanes of a quarter warp hit 8 distinct
        const float* bs = Bs[buf] + tx * kWgPitch;      // 16-byte bank groups (pitch 36 words)
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
            float4 av[4], bv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                av[j] = *reinterpret_cast<const float4*>(as + j * 16 * kWgPitch + r4 * 4);
                bv[j] = *reinterpret_cast<const float4*>(bs + j * 16 * kWgPitch + r4 * 4);
            }
            // the two updates of one accumulator are 16 instructions apart (no back-to-back dependent FFMA2)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const float2 alo = make_float2(av[y].x, av[y].y);
#pragma unroll
                for (int x = 0; x < 4; ++x) acc[y][x] = __ffma2_rn(alo, make_float2(bv[x].x, bv[x].y), acc[y][x]);
            }
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const float2 ahi = make_float2(av[y].z, av[y].w);
#pragma unroll
                for (int x = 0; x < 4; ++x) acc[y][x] = __ffma2_rn(ahi, make_float2(bv[x].z, bv[x].w), acc[y][x]);
            }
        }
        if (want_bias) {
            const float4 u = *reinterpret_cast<const float4*>(&As[buf][brow * kWgPitch + bcol]);
            const float4 v = *reinterpret_cast<const float4*>(&As[buf][brow * kWgPitch + bcol + 4]);
            bsum += ((u.x + u.y) + (u.z + u.w)) + ((v.x + v.y) + (v.z + v.w));
        }
        if (c + 1 < c1) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    float* out = pl.scratch + a.part_off + ((long long)(g * a.splits + split) * a.Do) * a.Di;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        const int o = o0 + ty + 16 * y;
        if (o >= a.Do) continue;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int i = i0 + tx + 16 * x;
            if (i < a.Di) out[(long long)o * a.Di + i] = acc[y][x].x + acc[y][x].y;
        }
    }
    if (want_bias) {
        bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);
        bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);
        if ((tid & 3) == 0 && o0 + brow < a.Do)
            pl.scratch[a.bpart_off + (long long)(g * a.splits + split) * a.Do + o0 + brow] = bsum;
    }
}

// Diagonal job: one CTA per (g, split) covers all Do <= 128 features; 2 threads per feature (8 of the 16 / 16 of the
// 32 tile rows each).
__device__ __forceinline__ void wgrad_diagonal(const WgradPlan& pl, const WgradJobPlan& a, int split, int g) {
    const int tile_w = pl.tile_w;
    const int f = threadIdx.x >> 1, half = threadIdx.x & 1;
    const int per = tile_w / 2;
    const long long per_split = a.diag_chunks_per_split;               // in operand chunks
    const long long c0 = (long long)split * per_split;
    const long long c1 = min(a.tile_chunks, c0 + per_split);
    float dot = 0.f, sum = 0.f;
    if (f < a.Do) {
        const float* Ag = a.A + (long long)g * a.a_group + (long long)f * tile_w + half * per;
        const float* Bg = a.B + (long long)g * a.b_group + (long long)f * tile_w + half * per;
#pragma unroll 4      // independent loads of several chunks in flight: the job is pure latency
        for (long long c = c0; c < c1; ++c) {
#pragma unroll
            for (int q = 0; q < per; q += 4) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(Ag + c * a.a_chunk + q));
                const float4 y = __ldg(reinterpret_cast<const float4*>(Bg + c * a.b_chunk + q));
                dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot); dot = fmaf(x.z, y.z, dot); dot = fmaf(x.w, y.w, dot);
                sum += (x.x + x.y) + (x.z + x.w);
            }
        }
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    if (f < a.Do && half == 0) {
        pl.scratch[a.part_off + (long long)(g * a.splits + split) * a.Do + f] = dot;
        if (a.bpart_off >= 0) pl.scratch[a.bpart_off + (long long)(g * a.splits + split) * a.Do + f] = sum;
    }
}

__global__ void __launch_bounds__(kWgThreads) wgrad_partial_kernel(const WgradPlan pl) {
    int j = 0;
    while (j + 1 < pl.n_jobs && (int)blockIdx.x >= pl.job[j + 1].cta_begin) ++j;
    const WgradJobPlan& a = pl.job[j];
    int local = blockIdx.x - a.cta_begin;
    const int per_g = a.tiles * a.splits;
    const int g = local / per_g;
    local -= g * per_g;
    const int split = local / a.tiles, tile = local % a.tiles;
    __shared__ __align__(16) float As[2][kWgTile * kWgPitch];
    __shared__ __align__(16) float Bs[2][kWgTile * kWgPitch];
    if (a.Di > 0) {
        if (pl.tile_w == 16)
            wgrad_matrix<16>(pl, a, tile, split, g, As, Bs);
        else
            wgrad_matrix<32>(pl, a, tile, split, g, As, Bs);
    } else
        wgrad_diagonal(pl, a, split, g);
}

// ==================================================================================================
// tensor-core variant (tcgen05 + TMEM, 3xTF32 split: fp32-accurate)
// ==================================================================================================
// One CTA = one 128 x Di output tile of one (job, controller, K split).  Warps 0-7 stream the two operands' 16-sample
// slabs (both are K-major: [feature][sample]) from HBM/L2, split every value x = hi + lo and store the four TF32 tiles
// {A.hi, A.lo, B.hi, B.lo} in the UMMA core-matrix layout (tc_dev.cuh); warp 8 issues six tcgen05.mma per slab
// (lo*hi, hi*lo, hi*hi for each of the two K = 8 steps) into one 128 x 128 fp32 accumulator in TMEM; a 4-stage mbarrier
// ring decouples them.  The bias sums (sum over the samples of A) are accumulated by the producers on the way.
constexpr int kWtStages = 4;
constexpr int kWtProducerWarps = 8;
constexpr int kWtThreads = (kWtProducerWarps + 1) * 32;
constexpr int kWtStageBytes = 4 * kTcTileBytes;                   // 32 KB
constexpr int kWtSmemBytes = kWtStages * kWtStageBytes + 1024;    // + barriers / TMEM pointer (and alignment slack)

__device__ __forceinline__ void wgrad_matrix_tc(const WgradPlan& pl, const WgradJobPlan& a, int tile, int split, int g) {
    extern __shared__ __align__(128) unsigned char wt_smem_raw[];
    const uint32_t smem_base = (smem_u32(wt_smem_raw) + 127u) & ~127u;
    unsigned char* smem = wt_smem_raw + (smem_base - smem_u32(wt_smem_raw));
    const uint32_t bars = smem_base + kWtStages * kWtStageBytes;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kWtStages, bar_acc = bars + 16 * kWtStages;
    volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kWtStages * kWtStageBytes + 16 * kWtStages + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int o0 = tile * kTcRows;
    const int tile_w = pl.tile_w, per = tile_w / kTcKB;           // 16-sample slabs per operand chunk
    const long long c0 = (long long)split * a.slabs_per_split;
    const long long c1 = min(a.slabs, c0 + a.slabs_per_split);
    const int n_slabs = (int)(c1 - c0);
    const bool want_bias = a.bpart_off >= 0;

    if (tid == 0) {
        for (int s = 0; s < kWtStages; ++s) {
            mbar_init(bar_full + 8 * s, kWtProducerWarps);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWtProducerWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_s)),
                     "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;

    if (warp < kWtProducerWarps) {
        const int r = tid & (kTcRows - 1), half = tid >> 7;       // operand row; samples 8*half .. 8*half+7 of the slab
        const bool a_ok = o0 + r < a.Do, b_ok = r < a.Di;
        const float* ap = a.A + (long long)g * a.a_group + (long long)(o0 + r) * tile_w + half * 8;
        const float* bp = a.B + (long long)g * a.b_group + (long long)r * tile_w + half * 8;
        // register prefetch ring, kWtAhead slabs deep: a slab is little work, so the loads of several slabs must be in
        // flight at once to cover the L2 / HBM latency
        constexpr int kWtAhead = 6;
        float4 xa[kWtAhead][2], xb[kWtAhead][2];
        // Slabs are fetched strictly in order, so the producers carry running pointers instead of recomputing
        // (chunk, offset) -> address per slab: 16-row tiles advance one operand chunk per slab, 32-row tiles alternate
        // between the two halves of a chunk.
        const float* pa = ap + (c0 / per) * a.a_chunk + (int)(c0 % per) * kTcKB;
        const float* pb = bp + (c0 / per) * a.b_chunk + (int)(c0 % per) * kTcKB;
        long long q_next = c0;                                   // slab the next fetch() loads
        const long long q_end = a.tile_chunks * per;             // slabs that exist (the last split may run past it)
        auto fetch = [&](long long, float4 (&ra)[2], float4 (&rb)[2]) {
            const bool live = q_next < q_end;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                ra[c] = (live && a_ok) ? __ldg(reinterpret_cast<const float4*>(pa) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                rb[c] = (live && b_ok) ? __ldg(reinterpret_cast<const float4*>(pb) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const bool wrap = per == 1 || (q_next & 1);           // next slab starts a new operand chunk
            pa += wrap ? a.a_chunk - (per - 1) * kTcKB : kTcKB;
            pb += wrap ? a.b_chunk - (per - 1) * kTcKB : kTcKB;
            ++q_next;
        };
        float bsum = 0.f;
#pragma unroll
        for (int u = 0; u < kWtAhead; ++u)
            if (u < n_slabs) fetch(c0 + u, xa[u], xb[u]);
        for (int i0 = 0; i0 < n_slabs; i0 += kWtAhead) {
#pragma unroll
            for (int u = 0; u < kWtAhead; ++u) {
                const int i = i0 + u;
                if (i >= n_slabs) continue;          // (no break: the ring index u must stay a compile-time constant)
                const int s = i % kWtStages;
                const uint32_t ph = (i / kWtStages) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                float* stage = reinterpret_cast<float*>(smem + s * kWtStageBytes);
                const float4 va[2] = {xa[u][0], xa[u][1]}, vb[2] = {xb[u][0], xb[u][1]};
                if (i + kWtAhead < n_slabs) fetch(c0 + i + kWtAhead, xa[u], xb[u]);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int off = tc_tile_off(r, half * 8 + c * 4);
                    const float4 ah = make_float4(tf32_hi(va[c].x), tf32_hi(va[c].y), tf32_hi(va[c].z), tf32_hi(va[c].w));
                    const float4 bh = make_float4(tf32_hi(vb[c].x), tf32_hi(vb[c].y), tf32_hi(vb[c].z), tf32_hi(vb[c].w));
                    *reinterpret_cast<float4*>(stage + off) = ah;
                    *reinterpret_cast<float4*>(stage + kTcTileBytes / 4 + off) =
                        make_float4(va[c].x - ah.x, va[c].y - ah.y, va[c].z - ah.z, va[c].w - ah.w);
                    *reinterpret_cast<float4*>(stage + 2 * (kTcTileBytes / 4) + off) = bh;
                    *reinterpret_cast<float4*>(stage + 3 * (kTcTileBytes / 4) + off) =
                        make_float4(vb[c].x - bh.x, vb[c].y - bh.y, vb[c].z - bh.z, vb[c].w - bh.w);
                    bsum += (va[c].x + va[c].y) + (va[c].z + va[c].w);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * s);
            }
        }
        // ---- epilogue: TMEM -> registers -> split-K partials ----
        if (n_slabs > 0) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
        }
        float* out = pl.scratch + a.part_off + ((long long)(g * a.splits + split) * a.Do) * a.Di;
        const int q4 = warp & 3, colh = warp >> 2;
        const int o = o0 + 32 * q4 + lane;
        const uint32_t trow = tmem + ((uint32_t)(32 * q4) << 16);
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            const int n0 = colh * 64 + cb * 16;
            if (n0 >= a.Di) break;                                // warp-uniform
            float v[16];
            if (n_slabs > 0) {
                tc_ld16(trow + n0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (o < a.Do) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    if (n0 + j < a.Di)                            // Di is a multiple of 4
                        *reinterpret_cast<float4*>(out + (long long)o * a.Di + n0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        if (want_bias) {                                          // the two sample halves of a row, through shared memory
            float* bias_s = reinterpret_cast<float*>(smem);       // stage 0 is free: every MMA has completed
            asm volatile("bar.sync 1, 256;" ::: "memory");
            bias_s[tid] = bsum;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0 && a_ok)
                pl.scratch[a.bpart_off + (long long)(g * a.splits + split) * a.Do + o0 + r] = bias_s[r] + bias_s[r + kTcRows];
        }
    } else {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcRows >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
        for (int i = 0; i < n_slabs; ++i) {
            const int s = i % kWtStages;
            const uint32_t ph = (i / kWtStages) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_base + s * kWtStageBytes;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t koff = j * 2 * kTcLBO;
                    const uint64_t a_hi = tc_smem_desc(sa + koff), a_lo = tc_smem_desc(sa + kTcTileBytes + koff);
                    const uint64_t b_hi = tc_smem_desc(sa + 2 * kTcTileBytes + koff), b_lo = tc_smem_desc(sa + 3 * kTcTileBytes + koff);
                    tc_mma_tf32(tmem, a_lo, b_hi, idesc, (i | j) != 0);
                    tc_mma_tf32(tmem, a_hi, b_lo, idesc, 1u);
                    tc_mma_tf32(tmem, a_hi, b_hi, idesc, 1u);
                }
                tc_commit(bar_empty + 8 * s);
                if (i == n_slabs - 1) tc_commit(bar_acc);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWtProducerWarps)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

__global__ void __launch_bounds__(kWtThreads, 1) wgrad_tc_kernel(const WgradPlan pl) {
    int j = 0;
    while (j + 1 < pl.n_jobs && (int)blockIdx.x >= pl.job[j + 1].cta_begin) ++j;
    const WgradJobPlan& a = pl.job[j];
    int local = blockIdx.x - a.cta_begin;
    const int per_g = a.tiles * a.splits;
    const int g = local / per_g;
    local -= g * per_g;
    const int split = local / a.tiles, tile = local % a.tiles;
    if (a.Di > 0) {
        wgrad_matrix_tc(pl, a, tile, split, g);
    } else if (threadIdx.x < kWgThreads) {      // whole warps 0-7: the diagonal form is plain FFMA
        wgrad_diagonal(pl, a, split, g);
    }
}

// One pass over every output of every job: sum the split-K partials in a fixed order.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgradPlan pl) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < pl.total_out;
         idx += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (j + 1 < pl.n_jobs && idx >= pl.job[j + 1].out_begin) ++j;
        const WgradJobPlan& a = pl.job[j];
        long long e = idx - a.out_begin;
        const long long per_w = (long long)a.Do * (a.Di > 0 ? a.Di : 1);
        const long long n_w = (long long)pl.G * per_w;
        const float* src;
        float* dst;
        long long per;
        if (e < n_w) {
            per = per_w;
            src = pl.scratch + a.part_off;
            dst = a.dW;
        } else {
            e -= n_w;
            per = a.Do;
            src = pl.scratch + a.bpart_off;
            dst = a.db;
        }
        const long long g = e / per, r = e - g * per;
        src += g * a.splits * per + r;
        float s = 0.f;
        for (int k = 0; k < a.splits; ++k) s += src[(long long)k * per];
        if (dst == a.db) {
            dst[g * a.db_group + r] = s;
        } else if (a.Di > 0) {
            const long long o = r / a.Di, i = r - o * a.Di;
            dst[g * a.dw_group + o * a.dw_row + i] = s;
            if (a.dW2) a.dW2[g * a.dw_group + o * a.dw_row + i] = a.scale2 * s;
        } else {
            dst[g * a.dw_group + r] = s;
        }
    }
}

// tc == false: 64 x 64 output tiles, 32-sample slabs, 2 resident CTAs per SM (FFMA kernel)
// tc == true : 128 x Di output tiles (Di <= 128), 16-sample slabs, 1 CTA per SM (tcgen05 kernel)
static int make_plan(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows, WgradPlan* pl, long long* scratch_floats,
                     bool tc = false) {
    const int tile_o = tc ? kTcRows : kWgTile;
    const int slab = tc ? kTcKB : 32;
    BIEAR_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= kWgMaxJobs, "biear_ctrl_wgrad: need 1..%d jobs, got %d", kWgMaxJobs, n_jobs);
    BIEAR_REQUIRE(G >= 1, "biear_ctrl_wgrad: G=%d", G);
    BIEAR_REQUIRE(tile_rows == 16 || tile_rows == 32, "biear_ctrl_wgrad: tile_rows must be 16 or 32, got %d", tile_rows);
    pl->n_jobs = n_jobs; pl->G = G; pl->tile_w = tile_rows;
    // tiles of all matrix jobs decide how finely K is split
    long long tile_ctas = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const BiearWgradJob& q = jobs[j];
        BIEAR_REQUIRE(q.Do >= 1 && q.Di >= 0 && q.chunks >= 1, "biear_ctrl_wgrad: job %d bad shape Do=%d Di=%d chunks=%lld", j, q.Do,
                      q.Di, (long long)q.chunks);
        BIEAR_REQUIRE(q.A && q.Bm && q.dW, "biear_ctrl_wgrad: job %d null pointer", j);
        BIEAR_REQUIRE(q.Di > 0 || q.Do <= kWgThreads / 2, "biear_ctrl_wgrad: diagonal job %d wider than %d", j, kWgThreads / 2);
        BIEAR_REQUIRE((q.a_chunk_stride & 3) == 0 && (q.b_chunk_stride & 3) == 0 && (q.a_group_stride & 3) == 0 &&
                          (q.b_group_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(q.A) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(q.Bm) & 15) == 0,
                      "biear_ctrl_wgrad: job %d operands must be 16-byte aligned with strides that are multiples of 4 floats", j);
        BIEAR_REQUIRE(!tc || (q.Di <= kTcRows && (q.Di & 3) == 0),
                      "biear_ctrl_wgrad_tc: job %d needs Di <= %d and a multiple of 4, got %d", j, kTcRows, q.Di);
        const int tiles = q.Di > 0 ? ((q.Do + tile_o - 1) / tile_o) * (tc ? 1 : (q.Di + kWgTile - 1) / kWgTile) : 1;
        if (q.Di > 0) tile_ctas += (long long)tiles * G;      // the (short) diagonal CTAs fill in behind
    }
    // 2 CTAs per SM are resident (124 registers x 256 threads): split K so that the matrix CTAs fill one wave
    long long want = tile_ctas > 0 ? ((tc ? 1LL : 2LL) * kSmCountB200) / tile_ctas : 1;
    if (want < 1) want = 1;
    long long off = 0, out = 0;
    int cta = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const BiearWgradJob& q = jobs[j];
        WgradJobPlan& a = pl->job[j];
        a.A = q.A; a.a_group = q.a_group_stride; a.a_chunk = q.a_chunk_stride; a.Do = q.Do;
        a.B = q.Bm; a.b_group = q.b_group_stride; a.b_chunk = q.b_chunk_stride; a.Di = q.Di;
        a.tile_chunks = q.chunks;
        a.slabs = (q.chunks * tile_rows + slab - 1) / slab;
        a.tiles_i = (q.Di > 0 && !tc) ? (q.Di + kWgTile - 1) / kWgTile : 1;
        a.tiles = q.Di > 0 ? ((q.Do + tile_o - 1) / tile_o) * a.tiles_i : 1;
        long long s = q.Di > 0 ? want : 4 * want;
        if (s > a.slabs) s = a.slabs;
        if (s < 1) s = 1;
        a.slabs_per_split = (a.slabs + s - 1) / s;
        a.splits = (int)((a.slabs + a.slabs_per_split - 1) / a.slabs_per_split);
        a.diag_chunks_per_split = (a.slabs_per_split * slab + tile_rows - 1) / tile_rows;
        a.cta_begin = cta;
        cta += a.tiles * a.splits * G;
        const long long per_w = (long long)q.Do * (q.Di > 0 ? q.Di : 1);
        a.part_off = off;
        off += ((long long)G * a.splits * per_w + 3) & ~3LL;      // every region starts 16-byte aligned
        a.bpart_off = -1;
        if (q.db) {
            a.bpart_off = off;
            off += ((long long)G * a.splits * q.Do + 3) & ~3LL;
        }
        a.out_begin = out;
        out += (long long)G * (per_w + (q.db ? q.Do : 0));
        a.dW = q.dW; a.db = q.db;
        a.dW2 = q.Di > 0 ? q.dW2 : nullptr; a.scale2 = q.scale2;
        a.dw_row = q.dw_row_stride ? q.dw_row_stride : (q.Di > 0 ? q.Di : 1);
        a.dw_group = q.dw_group_stride ? q.dw_group_stride : per_w;
        a.db_group = q.db_group_stride ? q.db_group_stride : q.Do;
    }
    pl->total_ctas = cta;
    pl->total_out = out;
    *scratch_floats = off;
    return 0;
}

}  // namespace biear

extern "C" int64_t biear_wgrad_scratch_floats(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows) {
    using namespace biear;
    WgradPlan pl;
    long long n = 0;
    if (make_plan(jobs, n_jobs, G, tile_rows, &pl, &n)) return -1;
    return n;
}

extern "C" int biear_ctrl_wgrad(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows, float* scratch, void* stream) {
    using namespace biear;
    WgradPlan pl;
    long long n = 0;
    if (int e = make_plan(jobs, n_jobs, G, tile_rows, &pl, &n)) return e;
    BIEAR_REQUIRE(scratch, "biear_ctrl_wgrad: null scratch");
    pl.scratch = scratch;
    cudaStream_t st = as_stream(stream);
    wgrad_partial_kernel<<<pl.total_ctas, kWgThreads, 0, st>>>(pl);
    BIEAR_LAUNCH_CHECK("wgrad_partial_kernel");
    const int blocks = (int)((pl.total_out + 255) / 256);
    wgrad_reduce_kernel<<<blocks < 4 * kSmCountB200 ? blocks : 4 * kSmCountB200, 256, 0, st>>>(pl);
    BIEAR_LAUNCH_CHECK("wgrad_reduce_kernel");
    return 0;
}

extern "C" int64_t biear_wgrad_scratch_floats_tc(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows) {
    using namespace biear;
    WgradPlan pl;
    long long n = 0;
    if (make_plan(jobs, n_jobs, G, tile_rows, &pl, &n, true)) return -1;
    return n;
}

extern "C" int biear_ctrl_wgrad_tc(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows, float* scratch, void* stream) {
    using namespace biear;
    WgradPlan pl;
    long long n = 0;
    if (int e = make_plan(jobs, n_jobs, G, tile_rows, &pl, &n, true)) return e;
    BIEAR_REQUIRE(scratch && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "biear_ctrl_wgrad_tc: scratch must be 16-byte aligned");
    pl.scratch = scratch;
    cudaStream_t st = as_stream(stream);
    static bool configured[64] = {false};
    int dev = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        if (int e = check_cuda(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWtSmemBytes),
                               "cudaFuncSetAttribute(wgrad_tc_kernel)"))
            return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    wgrad_tc_kernel<<<pl.total_ctas, kWtThreads, kWtSmemBytes, st>>>(pl);
    BIEAR_LAUNCH_CHECK("wgrad_tc_kernel");
    const int blocks = (int)((pl.total_out + 255) / 256);
    wgrad_reduce_kernel<<<blocks < 4 * kSmCountB200 ? blocks : 4 * kSmCountB200, 256, 0, st>>>(pl);
    BIEAR_LAUNCH_CHECK("wgrad_reduce_kernel");
    return 0;
}
