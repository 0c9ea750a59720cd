// Building blocks of the fused Q-controller step kernels (ctrl.cu).
//
// One thread-block CLUSTER of CS CTAs owns R = 4*CS controller rows (clips of one ear) and runs the
// whole controller step for them:  log1p features -> GRU cell -> Linear/LayerNorm/SiLU/Dropout x2 ->
// Linear -> tanh -> Q update.  The 173 k weights of a controller are SLICED across the cluster: CTA c
// owns hidden units [c*U, (c+1)*U), U = 128/CS, of every 128-wide layer (and a 1/CS slice of the bands
// of the last layer), stages only that slice in shared memory, and the CTAs exchange the small
// (R x 128) activations through distributed shared memory between layers.  Per step a cluster
// therefore reads each weight once from L2 instead of once per CTA, and each SM holds 1/CS of them.
//
// Thread layout (256 threads): tid = ks*128 + rg*U + u
//   u  in [0,U)   output unit inside the CTA's slice
//   rg in [0,CS)  row group: rows rg*4 .. rg*4+3 of the cluster's R rows
//   ks in {0,1}   half of the contraction (k) range; the halves are summed through shared memory
// Activations live in shared memory feature-major, [feature][R], so that a thread fetches its 4 rows of
// one feature with a single 128-bit load and a warp reads at most two distinct addresses.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace biear {
namespace cg = cooperative_groups;

constexpr int kHid = 128;           // GRU / MLP width (model_torch.py:256-267)
constexpr int kCtrlThreads = 256;
constexpr int kRT = 4;              // rows per thread
constexpr float kDropP = 0.1f;      // model_torch.py:261, 265
constexpr float kLnEps = 1e-5f;

template <int CS>
struct CtrlGeom {
    static constexpr int U = kHid / CS;      // units per CTA
    static constexpr int RG = CS;            // row groups
    static constexpr int R = RG * kRT;       // rows per cluster
    static_assert(2 * RG * U == kCtrlThreads, "thread layout");
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---- Philox4x32-10 (counter-based RNG for the dropout masks; regenerated, not stored) -------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned int hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned int hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// keep-mask scale (0 or 1/(1-p)) of element (step, layer, global row, feature)
__device__ __forceinline__ float dropout_scale(unsigned long long seed, int step, int layer, long long row, int f) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)row, (unsigned)(row >> 32), (unsigned)(step * 2 + layer), (unsigned)(f >> 2)),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const unsigned int v = (f & 3) == 0 ? r.x : (f & 3) == 1 ? r.y : (f & 3) == 2 ? r.z : r.w;
    const float uni = (float)(v >> 8) * (1.0f / 16777216.0f);    // [0,1)
    return uni >= kDropP ? 1.0f / (1.0f - kDropP) : 0.0f;
}

// acc[i] += sum_{k in [k0,k1)} x_s[k*R + i] * w[k]        (x_s already offset to the thread's rows)
template <int R>
__device__ __forceinline__ void dot_rows(float acc[kRT], const float* __restrict__ x_s, const float* __restrict__ w,
                                         int k0, int k1) {
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(x_s + k * R);
        const float wk = w[k];
        acc[0] = fmaf(wk, x.x, acc[0]);
        acc[1] = fmaf(wk, x.y, acc[1]);
        acc[2] = fmaf(wk, x.z, acc[2]);
        acc[3] = fmaf(wk, x.w, acc[3]);
    }
}

// Same contraction with the weights stored [k][pitch] (column slice, used by the transposed products):
// acc[i] += sum_k x_s[k*R + i] * w[k*pitch]
template <int R>
__device__ __forceinline__ void dot_rows_strided(float acc[kRT], const float* __restrict__ x_s,
                                                 const float* __restrict__ w, int pitch, int k0, int k1) {
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(x_s + k * R);
        const float wk = w[k * pitch];
        acc[0] = fmaf(wk, x.x, acc[0]);
        acc[1] = fmaf(wk, x.y, acc[1]);
        acc[2] = fmaf(wk, x.z, acc[2]);
        acc[3] = fmaf(wk, x.w, acc[3]);
    }
}

// Sum the two k-halves: the ks == 1 threads park their accumulators, the ks == 0 threads add them.
// red_s holds NACC floats for each of the 128 (rg,u) slots.  Contains two block barriers.
template <int NACC>
__device__ __forceinline__ void reduce_halves(float* acc, float* red_s, int ks, int slot) {
    __syncthreads();
    if (ks == 1) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) red_s[i * 128 + slot] = acc[i];
    }
    __syncthreads();
    if (ks == 0) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] += red_s[i * 128 + slot];
    }
}

// Write 4 row values of one feature into the [feature][R] buffer of every CTA of the cluster.
template <int CS, int R>
__device__ __forceinline__ void broadcast_rows(cg::cluster_group& cluster, float* buf_s, int feature, int row0,
                                               const float v[kRT]) {
    const float4 val = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
    for (int dst = 0; dst < CS; ++dst) {
        float* remote = cluster.map_shared_rank(buf_s, dst);
        *reinterpret_cast<float4*>(remote + feature * R + row0) = val;
    }
}

}  // namespace biear
