// Weight gradients of the controller's Linear layers from tile-layout operands (see include/biear_b200.h):
//   dW[g][o][i] = sum_{k < chunks} sum_{r < R} A[g][k][o][r] * Bm[g][k][i][r],   db[g][o] = sum_{k,r} A[g][k][o][r]
// with R = tile_rows (16 or 32; 32 / R consecutive chunks form one 32-sample slab)
// A = per-sample pre-activation gradients written by the backward recurrence, Bm = the layer inputs saved by the
// forward recurrence.  Both are "K-major in chunks of 32", so a 64 x 64 output tile streams [64][32] slabs of each
// operand straight into shared memory with 128-bit loads.  The contraction length (chunks*32 = (T-1)*B samples) is
// far longer than the outputs are wide, so the work is split along K over the grid; partials go to scratch and
// a second kernel sums them in a fixed order (deterministic, no atomics).
//
// Replaces what autograd + cuBLAS do for the weight gradients of model_torch.py:256-267 (GRU weight_ih / weight_hh,
// three Linear layers) in the reference.
#include "common.cuh"

namespace biear {

constexpr int kWgTile = 64;          // output tile (o and i)
constexpr int kWgThreads = 256;      // 16 x 16 threads, 4 x 4 outputs each
constexpr int kWgPitch = 36;         // padded slab row (floats): 16-byte aligned, conflict-free column reads

struct WgradArgs {
    const float* A; long long a_group, a_chunk; int Do;
    const float* B; long long b_group, b_chunk; int Di;
    int G; long long chunks; int splits; long long chunks_per_split;   // "chunks" here = 32-sample slabs
    long long tile_chunks; int tile_w;                                  // chunks of tile_w samples in the operands
    float* part;     // (G, splits, Do, Di)
    float* bpart;    // (G, splits, Do) or null
};

__global__ void __launch_bounds__(kWgThreads) wgrad_partial_kernel(const WgradArgs a) {
    __shared__ __align__(16) float As[2][kWgTile * kWgPitch];
    __shared__ __align__(16) float Bs[2][kWgTile * kWgPitch];
    const int tiles_i = (a.Di + kWgTile - 1) / kWgTile;
    const int to = blockIdx.x / tiles_i, ti = blockIdx.x % tiles_i;
    const int split = blockIdx.y, g = blockIdx.z;
    const int o0 = to * kWgTile, i0 = ti * kWgTile;
    const long long c0 = (long long)split * a.chunks_per_split;
    const long long c1 = min(a.chunks, c0 + a.chunks_per_split);
    const float* Ag = a.A + (long long)g * a.a_group;
    const float* Bg = a.B + (long long)g * a.b_group;
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;          // outputs o0 + ty + 16*{0..3}, i0 + tx + 16*{0..3}
    // slab loader: 64 rows x 8 float4 = 512 float4 per operand, 2 per thread
    const int lr = tid / 8, lc = tid % 8;            // rows lr and lr + 32, float4 column lc

    float acc[4][4] = {};
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
    float4 ra[2], rb[2];
    auto fetch = [&](long long c) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 32 * h;
            const long long cc = c * (32 / a.tile_w) + (lc * 4) / a.tile_w;      // operand chunk of this float4 column
            const int off = (lc * 4) % a.tile_w;
            const bool live = cc < a.tile_chunks;
            ra[h] = (live && o0 + r < a.Do) ? __ldg(reinterpret_cast<const float4*>(Ag + cc * a.a_chunk + (long long)(o0 + r) * a.tile_w + off))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[h] = (live && i0 + r < a.Di) ? __ldg(reinterpret_cast<const float4*>(Bg + cc * a.b_chunk + (long long)(i0 + r) * a.tile_w + off))
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
#pragma unroll
            for (int y = 0; y < 4; ++y) {
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    acc[y][x] = fmaf(av[y].x, bv[x].x, acc[y][x]);
                    acc[y][x] = fmaf(av[y].y, bv[x].y, acc[y][x]);
                    acc[y][x] = fmaf(av[y].z, bv[x].z, acc[y][x]);
                    acc[y][x] = fmaf(av[y].w, bv[x].w, acc[y][x]);
                }
                if (tx == 0) bacc[y] += (av[y].x + av[y].y) + (av[y].z + av[y].w);
            }
        }
        if (c + 1 < c1) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    float* out = a.part + ((long long)(g * a.splits + split) * a.Do) * a.Di;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        const int o = o0 + ty + 16 * y;
        if (o >= a.Do) continue;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int i = i0 + tx + 16 * x;
            if (i < a.Di) out[(long long)o * a.Di + i] = acc[y][x];
        }
        if (a.bpart && tx == 0 && ti == 0) a.bpart[(long long)(g * a.splits + split) * a.Do + o] = bacc[y];
    }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int splits, long long per_group,
                                                           int G, float* __restrict__ out) {
    const long long total = (long long)G * per_group;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long g = idx / per_group, e = idx - g * per_group;
        const float* src = part + g * splits * per_group + e;
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += src[(long long)k * per_group];
        out[idx] = s;
    }
}

static int pick_splits(int G, int Do, int Di, long long chunks) {
    const int tiles = ((Do + kWgTile - 1) / kWgTile) * ((Di + kWgTile - 1) / kWgTile) * G;
    long long s = (2LL * kSmCountB200 + tiles - 1) / tiles;      // ~2 CTAs per SM over the whole grid
    if (s > chunks) s = chunks;
    if (s < 1) s = 1;
    return (int)s;
}

}  // namespace biear

extern "C" int64_t biear_wgrad_scratch_floats(int G, int Do, int Di, int64_t chunks, int tile_rows) {
    using namespace biear;
    if (G < 1 || Do < 1 || Di < 1 || chunks < 1 || (tile_rows != 16 && tile_rows != 32)) return 0;
    const int s = pick_splits(G, Do, Di, (chunks * tile_rows + 31) / 32);
    return (int64_t)G * s * ((int64_t)Do * Di + Do);
}

extern "C" int biear_ctrl_wgrad(const float* A, int64_t a_group_stride, int64_t a_chunk_stride, int Do, const float* Bm,
                                int64_t b_group_stride, int64_t b_chunk_stride, int Di, int G, int64_t chunks,
                                int tile_rows, float* dW, float* db, float* scratch, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(G >= 1 && Do >= 1 && Di >= 1 && chunks >= 1, "biear_ctrl_wgrad: bad shape G=%d Do=%d Di=%d chunks=%lld", G,
                  Do, Di, (long long)chunks);
    BIEAR_REQUIRE(A && Bm && dW && scratch, "biear_ctrl_wgrad: null pointer");
    BIEAR_REQUIRE(tile_rows == 16 || tile_rows == 32, "biear_ctrl_wgrad: tile_rows must be 16 or 32, got %d", tile_rows);
    BIEAR_REQUIRE((a_chunk_stride & 3) == 0 && (b_chunk_stride & 3) == 0 && (a_group_stride & 3) == 0 &&
                      (b_group_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(Bm) & 15) == 0,
                  "biear_ctrl_wgrad: operands must be 16-byte aligned with strides that are multiples of 4 floats");
    cudaStream_t st = as_stream(stream);
    WgradArgs a;
    a.A = A; a.a_group = a_group_stride; a.a_chunk = a_chunk_stride; a.Do = Do;
    a.B = Bm; a.b_group = b_group_stride; a.b_chunk = b_chunk_stride; a.Di = Di;
    a.G = G; a.tile_chunks = chunks; a.tile_w = tile_rows;
    a.chunks = (chunks * tile_rows + 31) / 32;
    a.splits = pick_splits(G, Do, Di, a.chunks);
    a.chunks_per_split = (a.chunks + a.splits - 1) / a.splits;
    a.part = scratch;
    a.bpart = db ? scratch + (long long)G * a.splits * Do * Di : nullptr;
    const int tiles = ((Do + kWgTile - 1) / kWgTile) * ((Di + kWgTile - 1) / kWgTile);
    wgrad_partial_kernel<<<dim3(tiles, a.splits, G), kWgThreads, 0, st>>>(a);
    BIEAR_LAUNCH_CHECK("wgrad_partial_kernel");
    const long long per_w = (long long)Do * Di;
    wgrad_reduce_kernel<<<(int)((G * per_w + 255) / 256), 256, 0, st>>>(a.part, a.splits, per_w, G, dW);
    BIEAR_LAUNCH_CHECK("wgrad_reduce_kernel(dW)");
    if (db) {
        wgrad_reduce_kernel<<<(G * Do + 255) / 256, 256, 0, st>>>(a.bpart, a.splits, Do, G, db);
        BIEAR_LAUNCH_CHECK("wgrad_reduce_kernel(db)");
    }
    return 0;
}
