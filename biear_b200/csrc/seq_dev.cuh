// Building blocks of the persistent adaptive-recurrence kernels (seq.cu).
//
// One thread-block CLUSTER of 4 CTAs owns a TILE of 16 rows (clips of one ear = one controller) for the WHOLE
// 19-frame recurrence: band stage of frame t -> log1p features -> GRU cell -> Linear/LayerNorm/SiLU/Dropout x2
// -> Linear -> tanh -> Q_{t+1} -> band stage of frame t+1 ...  Nothing on the serial chain goes through HBM:
//   * the controller's 173 k weights are SLICED across the cluster -- CTA c owns hidden units [32c, 32c+32) of
//     every 128-wide layer and a 1/4 slice of the bands of the last layer -- and stay resident in shared memory
//     for all steps (a 137 KB "image" per CTA, laid out by a pack kernel so that it is copied with 128-bit loads);
//   * the (16 x 128) activations are exchanged between the CTAs through distributed shared memory;
//   * CTA c runs the band stage for rows 4c .. 4c+3 of the tile and broadcasts log1p(Y) to its peers; the Q of
//     those rows comes back from the CTAs that own the corresponding bands of the last layer.
//
// Why 4 x 16 and not 8 x 32: a B200 fits 15 clusters of 8 CTAs with this much shared memory but 36 clusters of 4
// (cudaOccupancyMaxActiveClusters, tools/occupancy.py), and the benchmark batch (256 clips x 2 ears) needs 16 tiles
// of 32 rows -- one more than fit, i.e. two waves -- or 32 tiles of 16 rows, which run as one wave on 128 SMs.
//
// Thread layout of the GEMM phases (512 threads): tid = ks*128 + rg*32 + u
//   u  in [0,32)  output unit inside the CTA's slice
//   rg in [0,4)   row group: rows 4rg .. 4rg+3 of the tile (the rows whose band stage CTA rg runs)
//   ks in [0,4)   quarter of the contraction range; the quarters are summed through shared memory
// (16 warps per SM: the per-frame work is a chain of short dependent phases, so it is latency- not throughput-bound;
//  with 8 warps the schedulers issued on 40 % of the cycles, ncu profiles/r1_*)
// Activations live feature-major, [feature][32 rows], in shared memory AND in the tensors saved for the backward
// pass ("tile layout": (G, T-1, tiles, D, 16)), so a thread moves its 4 rows of one feature with one 128-bit access.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "tc_dev.cuh"

namespace biear {
namespace cg = cooperative_groups;

constexpr int kHid = 128;           // GRU / MLP width (model_torch.py:256-267)
constexpr int kSeqThreads = 512;
constexpr int kKS = 4;              // k-splits of every contraction
constexpr int kCS = 4;              // CTAs per cluster
constexpr int kU = kHid / kCS;      // hidden units per CTA (32)
constexpr int kR = 16;              // rows per tile
constexpr int kRT = 4;              // rows per thread = rows per CTA in the band stage
constexpr float kDropP = 0.1f;      // model_torch.py:261, 265
constexpr float kLnEps = 1e-5f;
static_assert(kKS * (kR / kRT) * kU == kSeqThreads, "thread layout");
static_assert(kR / kRT == kCS, "row groups == CTAs of the cluster");

__host__ __device__ constexpr int bands_per_cta(int N) { return (N + kCS - 1) / kCS; }

// ---- shared-memory weight images (offsets in floats) ------------------------------------------------------
// forward image of CTA (g, c):
//   wih3 [k<N][gate<3][u]   Weff[gate,u][k],  Weff = W_ih[:, :N] + 0.2 W_ih[:, N:]
//                           (feat = [yc, 0.2*yc.detach()]  =>  W_ih feat = Weff yc)
//   whh3 [k<128][gate<3][u] W_hh[gate,u][k]
//   w1 [k<128][u], w2 [k<128][u], w3 [k<128][u] (band c*NU+u; zero beyond the CTA's slice)
__host__ __device__ constexpr int fwd_img_wih(int) { return 0; }
__host__ __device__ constexpr int fwd_img_whh(int N) { return N * 3 * kU; }
__host__ __device__ constexpr int fwd_img_w1(int N) { return fwd_img_whh(N) + kHid * 3 * kU; }
__host__ __device__ constexpr int fwd_img_w2(int N) { return fwd_img_w1(N) + kHid * kU; }
__host__ __device__ constexpr int fwd_img_w3(int N) { return fwd_img_w2(N) + kHid * kU; }
__host__ __device__ constexpr int fwd_img_floats(int N) { return fwd_img_w3(N) + kHid * kU; }
// backward image of CTA (g, c): column slices for the transposed products
//   w3c [n<N][u] = W3[n][32c+u];  w2c, w1c [o<128][u] = W[o][32c+u];  whhc [o<384][u] = W_hh[o][32c+u];
//   wihc [o<384][u] = W_ih[o][c*NU+u] (first N columns only: the detached half of feat carries no gradient)
__host__ __device__ constexpr int bwd_img_w3c(int) { return 0; }
__host__ __device__ constexpr int bwd_img_w2c(int N) { return N * kU; }
__host__ __device__ constexpr int bwd_img_w1c(int N) { return bwd_img_w2c(N) + kHid * kU; }
__host__ __device__ constexpr int bwd_img_whhc(int N) { return bwd_img_w1c(N) + kHid * kU; }
__host__ __device__ constexpr int bwd_img_wihc(int N) { return bwd_img_whhc(N) + 3 * kHid * kU; }
__host__ __device__ constexpr int bwd_img_floats(int N) { return bwd_img_wihc(N) + 3 * kHid * kU; }

// ---- single-controller variant (seq_single.cu; model_torch.py:579-776): GRU(4N -> 128), one Q for both ears ------------
// Workspace: [kCS forward resident images: whh3 | w1 | w2 | w3, skewed columns][kCS streamed W_ih images [Kp][3][32],
// input order cL | cR | mL | mR][kCS backward resident images: w3c | w2c | w1c | whhc][kCS streamed W_ih^T column images:
// wihcL | wihcR, [384][32] each].
__host__ __device__ constexpr int single_kp(int N) { return (4 * N + 7) & ~7; }              // controller input width, padded
__host__ __device__ constexpr int single_chunk_rows(int N) {                                   // k-rows per W_ih ring stage
    const int m = single_kp(N) / 8;
    return 8 * (m % 5 == 0 ? 5 : (m % 4 == 0 ? 4 : (m % 3 == 0 ? 3 : (m % 2 == 0 ? 2 : 1))));
}
__host__ __device__ constexpr int single_fwd_res_floats() { return kHid * 3 * kU + 3 * kHid * kU; }
__host__ __device__ constexpr int single_bwd_res_floats(int N) { return bwd_img_wihc(N); }
__host__ __device__ constexpr long long single_off_fres() { return 0; }
__host__ __device__ constexpr long long single_off_fstr(int) { return (long long)kCS * single_fwd_res_floats(); }
__host__ __device__ constexpr long long single_off_bres(int N) { return single_off_fstr(N) + (long long)kCS * single_kp(N) * 3 * kU; }
__host__ __device__ constexpr long long single_off_bstr(int N) { return single_off_bres(N) + (long long)kCS * single_bwd_res_floats(N); }
__host__ __device__ constexpr long long single_workspace_floats(int N) { return single_off_bstr(N) + (long long)kCS * 2 * 3 * kHid * kU; }
size_t single_fwd_smem_bytes(int N, int F);
int launch_single_prepare(const BiearSeqParams* p, int want, cudaStream_t st);
int launch_single_fwd(const BiearSeqParams* p, cudaStream_t st);
int debug_phase_cycles_single(unsigned long long* out_host);

// ---- Philox4x32-10 (counter-based RNG for the dropout masks; regenerated in the backward, not stored) ------
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

__device__ __forceinline__ float keep_scale(unsigned int v) {
    const float uni = (float)(v >> 8) * (1.0f / 16777216.0f);    // [0,1)
    return uni >= kDropP ? 1.0f / (1.0f - kDropP) : 0.0f;
}

// keep-mask scales (0 or 1/(1-p)) of features 4*fq .. 4*fq+3 of (step, layer, global row)
__device__ __forceinline__ float4 dropout_scale4(unsigned long long seed, int step, int layer, long long row, int fq) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)row, (unsigned)(row >> 32), (unsigned)(step * 2 + layer), (unsigned)fq),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    return make_float4(keep_scale(r.x), keep_scale(r.y), keep_scale(r.z), keep_scale(r.w));
}

// 128-bit global -> shared copy of `n4` float4 by the whole CTA, loads batched 8 deep
__device__ __forceinline__ void copy_f4(float4* __restrict__ dst, const float4* __restrict__ src, int n4) {
    int i = threadIdx.x;
    for (; i + 7 * kSeqThreads < n4; i += 8 * kSeqThreads) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(src + i + j * kSeqThreads);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[i + j * kSeqThreads] = v[j];
    }
    for (; i < n4; i += kSeqThreads) dst[i] = __ldg(src + i);
}

// acc[i] += sum_{k in [k0,k1)} x_s[k*kR + i] * w_s[k*kU]     (x_s / w_s already offset to the thread's rows / unit)
// The 4 rows are two packed fp32x2 FMAs (FFMA2, sm_100): same rounding as four scalar FMAs, half the issue slots.
template <int RS = kR>
__device__ __forceinline__ void dot_rows(float acc[kRT], const float* __restrict__ x_s, const float* __restrict__ w_s,
                                         int k0, int k1) {
#ifdef BIEAR_SKIP_DOTS   // timing experiment only: what the phases cost without their contractions
    k1 = k0;
#endif
    float2 lo = make_float2(acc[0], acc[1]), hi = make_float2(acc[2], acc[3]);
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(x_s + k * RS);
        const float wk = w_s[k * kU];
        const float2 w2 = make_float2(wk, wk);
        lo = __ffma2_rn(w2, make_float2(x.x, x.y), lo);
        hi = __ffma2_rn(w2, make_float2(x.z, x.w), hi);
    }
    acc[0] = lo.x; acc[1] = lo.y; acc[2] = hi.x; acc[3] = hi.y;
}

// Two contractions that share the activations (rows k of x_s) but not the weights:
//   accA[i] += sum_k x_s[k*kR+i] * wa_s[k*kU],   accB[i] += sum_k x_s[k*kR+i] * wb_s[k*kU]        (one x load per k)
template <int RS = kR>
__device__ __forceinline__ void dot_rows_pair(float accA[kRT], float accB[kRT], const float* __restrict__ x_s,
                                              const float* __restrict__ wa_s, const float* __restrict__ wb_s, int k0, int k1) {
#ifdef BIEAR_SKIP_DOTS
    k1 = k0;
#endif
    float2 al = make_float2(accA[0], accA[1]), ah = make_float2(accA[2], accA[3]);
    float2 bl = make_float2(accB[0], accB[1]), bh = make_float2(accB[2], accB[3]);
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(x_s + k * RS);
        const float2 xl = make_float2(x.x, x.y), xh = make_float2(x.z, x.w);
        const float wa = wa_s[k * kU], wb = wb_s[k * kU];
        const float2 pa = make_float2(wa, wa), pb = make_float2(wb, wb);
        al = __ffma2_rn(pa, xl, al); ah = __ffma2_rn(pa, xh, ah);
        bl = __ffma2_rn(pb, xl, bl); bh = __ffma2_rn(pb, xh, bh);
    }
    accA[0] = al.x; accA[1] = al.y; accA[2] = ah.x; accA[3] = ah.y;
    accB[0] = bl.x; accB[1] = bl.y; accB[2] = bh.x; accB[3] = bh.y;
}

// Three gate rows at once: the weights of one k are [gate][kU] (w3_s already offset to the thread's unit).
template <int RS = kR>
__device__ __forceinline__ void dot_rows3(float a0[kRT], float a1[kRT], float a2[kRT], const float* __restrict__ x_s,
                                          const float* __restrict__ w3_s, int k0, int k1) {
#ifdef BIEAR_SKIP_DOTS
    k1 = k0;
#endif
    float2 l0 = make_float2(a0[0], a0[1]), h0 = make_float2(a0[2], a0[3]);
    float2 l1 = make_float2(a1[0], a1[1]), h1 = make_float2(a1[2], a1[3]);
    float2 l2 = make_float2(a2[0], a2[1]), h2 = make_float2(a2[2], a2[3]);
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(x_s + k * RS);
        const float2 xl = make_float2(x.x, x.y), xh = make_float2(x.z, x.w);
        const float w0 = w3_s[k * 3 * kU], w1 = w3_s[k * 3 * kU + kU], w2 = w3_s[k * 3 * kU + 2 * kU];
        const float2 p0 = make_float2(w0, w0), p1 = make_float2(w1, w1), p2 = make_float2(w2, w2);
        l0 = __ffma2_rn(p0, xl, l0); h0 = __ffma2_rn(p0, xh, h0);
        l1 = __ffma2_rn(p1, xl, l1); h1 = __ffma2_rn(p1, xh, h1);
        l2 = __ffma2_rn(p2, xl, l2); h2 = __ffma2_rn(p2, xh, h2);
    }
    a0[0] = l0.x; a0[1] = l0.y; a0[2] = h0.x; a0[3] = h0.y;
    a1[0] = l1.x; a1[1] = l1.y; a1[2] = h1.x; a1[3] = h1.y;
    a2[0] = l2.x; a2[1] = l2.y; a2[2] = h2.x; a2[3] = h2.y;
}

// Generic-pitch forms for the GRU layer kernels (gru.cu): activations [k][16 rows] (x_s offset to the thread's 4 rows),
// weights [k][3 gates][pitch] resp. [o][pitch] (w_s offset to the thread's unit).
__device__ __forceinline__ void dot_rows3x(float a0[kRT], float a1[kRT], float a2[kRT], const float* __restrict__ x_s,
                                           const float* __restrict__ w_s, int pitch, int k0, int k1) {
    float2 l0 = make_float2(a0[0], a0[1]), h0 = make_float2(a0[2], a0[3]);
    float2 l1 = make_float2(a1[0], a1[1]), h1 = make_float2(a1[2], a1[3]);
    float2 l2 = make_float2(a2[0], a2[1]), h2 = make_float2(a2[2], a2[3]);
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(x_s + k * kR);
        const float2 xl = make_float2(x.x, x.y), xh = make_float2(x.z, x.w);
        const float w0 = w_s[(k * 3) * pitch], w1 = w_s[(k * 3 + 1) * pitch], w2 = w_s[(k * 3 + 2) * pitch];
        const float2 p0 = make_float2(w0, w0), p1 = make_float2(w1, w1), p2 = make_float2(w2, w2);
        l0 = __ffma2_rn(p0, xl, l0); h0 = __ffma2_rn(p0, xh, h0);
        l1 = __ffma2_rn(p1, xl, l1); h1 = __ffma2_rn(p1, xh, h1);
        l2 = __ffma2_rn(p2, xl, l2); h2 = __ffma2_rn(p2, xh, h2);
    }
    a0[0] = l0.x; a0[1] = l0.y; a0[2] = h0.x; a0[3] = h0.y;
    a1[0] = l1.x; a1[1] = l1.y; a1[2] = h1.x; a1[3] = h1.y;
    a2[0] = l2.x; a2[1] = l2.y; a2[2] = h2.x; a2[3] = h2.y;
}

__device__ __forceinline__ void dot_rows_p(float acc[kRT], const float* __restrict__ x_s, const float* __restrict__ w_s,
                                           int pitch, int k0, int k1) {
    float2 lo = make_float2(acc[0], acc[1]), hi = make_float2(acc[2], acc[3]);
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(x_s + k * kR);
        const float wk = w_s[k * pitch];
        const float2 w2 = make_float2(wk, wk);
        lo = __ffma2_rn(w2, make_float2(x.x, x.y), lo);
        hi = __ffma2_rn(w2, make_float2(x.z, x.w), hi);
    }
    acc[0] = lo.x; acc[1] = lo.y; acc[2] = hi.x; acc[3] = hi.y;
}

// [k0, k1) of k-split ks over a contraction of length K
__device__ __forceinline__ void k_range(int K, int ks, int& k0, int& k1) {
    k0 = (K * ks) / kKS;
    k1 = (K * (ks + 1)) / kKS;
}

// Sum the four k-splits in two rounds (ks 2,3 -> ks 0,1; then ks 1 -> ks 0); the result lands in the ks == 0 threads.
// red_s holds 2 * NACC floats for each of the 128 (rg,u) slots.  Contains four block barriers.
template <int NACC>
__device__ __forceinline__ void reduce_ks(float* acc, float* red_s, int ks, int slot) {
    __syncthreads();
    if (ks >= 2) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) red_s[((ks - 2) * NACC + i) * 128 + slot] = acc[i];
    }
    __syncthreads();
    if (ks < 2) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] += red_s[(ks * NACC + i) * 128 + slot];
    }
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

// Same sum in ONE round: every ks > 0 thread parks its accumulators, the ks == 0 thread adds the three of them in a
// fixed order.  Needs (kKS - 1) * NACC floats per slot of scratch but only two block barriers.
template <int NACC>
__device__ __forceinline__ void reduce_ks1(float* acc, float* red_s, int ks, int slot) {
    __syncthreads();
    if (ks > 0) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) red_s[((ks - 1) * NACC + i) * 128 + slot] = acc[i];
    }
    __syncthreads();
    if (ks == 0) {
#pragma unroll
        for (int k = 0; k < kKS - 1; ++k)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] += red_s[(k * NACC + i) * 128 + slot];
    }
}

// Write 4 row values of one feature into the [feature][kR] buffer of every CTA of the cluster.
__device__ __forceinline__ void broadcast_rows(cg::cluster_group& cluster, float* buf_s, int feature, int row0,
                                               const float v[kRT]) {
    const float4 val = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
    for (int dst = 0; dst < kCS; ++dst) {
        float* remote = cluster.map_shared_rank(buf_s, dst);
        *reinterpret_cast<float4*>(remote + feature * kR + row0) = val;
    }
}

// ---- cluster exchange without the hardware cluster barrier ---------------------------------------------------------
// Every hand-over between the CTAs of a cluster (activations of one phase going to the peers) is a set of st.async
// stores: the store carries its own completion to an mbarrier in the DESTINATION CTA (complete_tx of its 16 bytes), so a
// consumer only waits on a local mbarrier until all the bytes of the phase have landed.  Compared with
// "DSMEM stores + barrier.cluster.arrive.release / wait.acquire" this removes, per hand-over, the release fence (which
// also drains the thread's outstanding GLOBAL stores -- the saved state -- and showed up as `membar` stalls in ncu), the
// ~500-cycle hardware barrier and its L1 invalidation.  One mbarrier per phase type, used once per step, parity = step & 1.
// Ordering argument (why no all-CTA barrier is needed): a CTA can only run ahead of a peer by less than one phase type --
// to send phase p of step s+1 it must have RECEIVED every peer's phase p-1 .. of step s+1 / phase last of step s, which
// the peer sends only after it finished reading the buffers of the phases before (see seq.cu for the buffer-by-buffer list).
__device__ __forceinline__ uint32_t cluster_addr(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}

__device__ __forceinline__ void st_async_f4(uint32_t remote_addr, const float4 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar)
                 : "memory");
}

__device__ __forceinline__ void st_async_f2(uint32_t remote_addr, const float2 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(remote_addr), "f"(v.x), "f"(v.y), "r"(remote_bar)
                 : "memory");
}

// 16 bytes to the same shared-memory location of every CTA of the cluster, each signalling that CTA's mbarrier `bar`.
__device__ __forceinline__ void bcast_f4_tx(const float* dst_s, const float4 v, uint32_t bar) {
    const uint32_t a = smem_u32(dst_s);
#pragma unroll
    for (uint32_t dst = 0; dst < (uint32_t)kCS; ++dst) st_async_f4(cluster_addr(a, dst), v, cluster_addr(bar, dst));
}

// Write 4 row values of one feature into the [feature][kR] buffer of every CTA of the cluster (st.async form).
__device__ __forceinline__ void broadcast_rows_tx(float* buf_s, int feature, int row0, const float v[kRT], uint32_t bar) {
    bcast_f4_tx(buf_s + feature * kR + row0, make_float4(v[0], v[1], v[2], v[3]), bar);
}

// Wait for phase `parity` of a local mbarrier (all bytes of the hand-over have landed).  Bounded: a protocol bug must
// end in a launch failure, never in a hung GPU.
// (CTA-scope acquire, like every cluster kernel that waits on a local mbarrier for remote producers: the data and its
// complete_tx arrive in this SM's own shared memory, which no cache sits in front of; a cluster-scope acquire would make
// ptxas add an L1 invalidation -- CCTL.IVALL -- to every wait.)
__device__ __forceinline__ void tx_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (unsigned spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) __trap();
    }
}

__device__ __forceinline__ void store4(float* p, const float v[kRT]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// 1 / (1 + exp(-x)) with ex2.approx and an approximate reciprocal: ~1e-6 relative, a fraction of the instructions of
// expf + an IEEE division; used identically in the forward and the backward pass (the SiLU derivative reuses it).
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.402823466e+38f; }

}  // namespace biear
