// The eight per-sector heads of the back-end as ONE forward and ONE backward launch.
//
// Replaces model_torch.py:869-906 (SubHead: shared Linear(200,100)-ReLU-Dropout(0.2), then three branches
// Linear(100,50)-ReLU-Linear(50,10)-ReLU-Linear(10,k) for presence (k = 1), angle (k = 1, sigmoid) and distance class
// (k = 5)) and the loop over the heads in model_torch.py:941-955 / 1096-1110 -- 80 tiny GEMMs + ~100 element-wise
// launches forward and three times that backward per training step in the reference formulation.
//
// One CTA = (sector head s, tile of 32 clips).  The head's 36.9 k weights (147 KB) are read once from the nn.Linear
// parameters themselves (device pointer table, torch (out, in) layout) and transposed into shared memory with an odd
// pitch, so that the forward product (lanes = output units) and the transposed product of the backward (lanes = input
// features) both read them conflict-free; activations live feature-major [feature][36] so that a thread's 16 rows of one
// input feature are four broadcast 128-bit loads.  Everything between the body features and the three outputs stays in
// shared memory; the backward recomputes the forward of its tile (cheaper than saving and re-reading the activations) and
// regenerates the dropout mask from the Philox key.  Weight gradients: thread = input feature, its 32 rows in registers,
// one dot product per output unit, written as per-(tile, head) partials that a fixed-order second pass sums
// (deterministic, no atomics); dL/dbody likewise as per-head partials.
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "seq_dev.cuh"   // philox4x32_10

namespace biear {
namespace hd {

constexpr int kThreads = 256;
constexpr int kRows = 32;          // clips per CTA
constexpr int kHP = 36;            // pitch of the [feature][row] activation buffers (16-byte aligned rows, 4-way store conflicts)
constexpr int kH1 = 100, kH2 = 50, kH3 = 10;   // model_torch.py:872-896
constexpr int kBranches = 3;
constexpr int kTensorsPerHead = 2 + kBranches * 6;
constexpr float kDropP = 0.2f;     // model_torch.py:875
constexpr int kMaxC = 8;

__host__ __device__ constexpr int branch_out(int j, int C) { return j == 2 ? C : 1; }
__host__ __device__ constexpr int odd(int n) { return n | 1; }
// flat per-head layout of the gradient buffer (== state-dict order of SubHead)
__host__ __device__ constexpr int flat_branch(int j, int D, int C) {
    int off = kH1 * D + kH1;
    for (int i = 0; i < j; ++i) off += kH2 * kH1 + kH2 + kH3 * kH2 + kH3 + branch_out(i, C) * kH3 + branch_out(i, C);
    return off;
}
__host__ __device__ constexpr int flat_floats(int D, int C) { return flat_branch(kBranches, D, C); }

struct Smem {   // offsets in floats
    int D, C;
    __host__ __device__ Smem(int D_, int C_) : D(D_), C(C_) {}
    __host__ __device__ int x() const { return 0; }                                   // [D][kHP]
    __host__ __device__ int ws() const { return x() + D * kHP; }                      // [D][odd(100)]
    __host__ __device__ int h() const { return ws() + D * odd(kH1); }                 // [100][kHP] post ReLU + dropout
    __host__ __device__ int hs() const { return h() + kH1 * kHP; }                    // [100][kHP] d out / d pre of the shared layer
    __host__ __device__ int dh() const { return hs() + kH1 * kHP; }                   // [100][kHP] dL/dh (backward)
    __host__ __device__ int w1() const { return dh() + kH1 * kHP; }                   // [100][odd(50)]
    __host__ __device__ int w2() const { return w1() + kH1 * odd(kH2); }              // [50][odd(10)]
    __host__ __device__ int w3() const { return w2() + kH2 * odd(kH3); }              // [10][odd(C)]
    __host__ __device__ int a1() const { return w3() + kH3 * odd(kMaxC); }            // [50][kHP]
    __host__ __device__ int a2() const { return a1() + kH2 * kHP; }                   // [10][kHP]
    __host__ __device__ int da1() const { return a2() + kH3 * kHP; }                  // [50][kHP]
    __host__ __device__ int da2() const { return da1() + kH2 * kHP; }                 // [10][kHP]
    __host__ __device__ int o() const { return da2() + kH3 * kHP; }                   // [kMaxC][kHP] outputs / their gradients
    __host__ __device__ int bias() const { return o() + kMaxC * kHP; }                // 100 + 50 + 10 + 8, padded to 192
    __host__ __device__ int total() const { return bias() + 192; }
};

// torch-layout W (out, in) -> dst[k * pitch + o]  (coalesced global reads; conflict-free stores for an odd pitch)
// (128-bit loads, 8 in flight per thread: the copy is bound by the latency of its loads -- 32 KB in flight per CTA)
__device__ __forceinline__ void load_wT(float* __restrict__ dst, const float* __restrict__ W, int out, int in, int pitch) {
    const int n = out * in;
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
        const float4* W4 = reinterpret_cast<const float4*>(W);
        const int n4 = n >> 2;
        for (int base = threadIdx.x; base < n4; base += 8 * kThreads) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i4 = base + j * kThreads;
                v[j] = i4 < n4 ? __ldg(W4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i4 = base + j * kThreads;
                if (i4 < n4) {
                    const float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                    int o = (4 * i4) / in, k = 4 * i4 - o * in;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        dst[k * pitch + o] = e[c];
                        if (++k == in) { k = 0; ++o; }
                    }
                }
            }
        }
        return;
    }
    for (int idx = threadIdx.x; idx < n; idx += kThreads) {
        const int o = idx / in, k = idx - o * in;
        dst[k * pitch + o] = __ldg(W + idx);
    }
}

// out_s[o][r] = act(b[o] + sum_k in_s[k][r] * WT_s[k * pitch + o])  for o < outN and the tile's 32 rows.
// thread = (unit pair = tid % 64 -> units 2p, 2p + 1; row group = tid / 64 -> 8 rows): a 2 x 8 register tile, so that per k
// the warp spends 10 shared-memory cycles (two 128-bit broadcast loads of x, two weights) on 16 FMAs -- with one unit per
// thread the products are bound by the shared-memory pipe, not by the FMA pipe.  outN <= 128.  ACT: 0 none, 1 ReLU.
template <int ACT>
__device__ __forceinline__ void dense(const float* __restrict__ in_s, const float* __restrict__ wT_s, int pitch,
                                      const float* __restrict__ b_s, float* __restrict__ out_s, int K, int outN) {
    const int o0 = (threadIdx.x & 63) * 2, rg = threadIdx.x >> 6;
    if (o0 < outN) {
        const bool two = o0 + 1 < outN;
        const int o1 = two ? o0 + 1 : o0;
        float2 a0[4], a1[4];      // 8 rows as 4 packed pairs per unit
        const float b0 = b_s[o0], b1 = b_s[o1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a0[i] = make_float2(b0, b0);
            a1[i] = make_float2(b1, b1);
        }
        const float* xin = in_s + rg * 8;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float w0 = wT_s[k * pitch + o0], w1 = wT_s[k * pitch + o1];
            const float2 p0 = make_float2(w0, w0), p1 = make_float2(w1, w1);
            const float4 xa = *reinterpret_cast<const float4*>(xin + k * kHP);
            const float4 xb = *reinterpret_cast<const float4*>(xin + k * kHP + 4);
            const float2 x0 = make_float2(xa.x, xa.y), x1 = make_float2(xa.z, xa.w), x2 = make_float2(xb.x, xb.y), x3 = make_float2(xb.z, xb.w);
            a0[0] = __ffma2_rn(p0, x0, a0[0]); a0[1] = __ffma2_rn(p0, x1, a0[1]);
            a0[2] = __ffma2_rn(p0, x2, a0[2]); a0[3] = __ffma2_rn(p0, x3, a0[3]);
            a1[0] = __ffma2_rn(p1, x0, a1[0]); a1[1] = __ffma2_rn(p1, x1, a1[1]);
            a1[2] = __ffma2_rn(p1, x2, a1[2]); a1[3] = __ffma2_rn(p1, x3, a1[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            out_s[o0 * kHP + rg * 8 + 2 * i] = ACT == 1 ? fmaxf(a0[i].x, 0.0f) : a0[i].x;
            out_s[o0 * kHP + rg * 8 + 2 * i + 1] = ACT == 1 ? fmaxf(a0[i].y, 0.0f) : a0[i].y;
            if (two) {
                out_s[o1 * kHP + rg * 8 + 2 * i] = ACT == 1 ? fmaxf(a1[i].x, 0.0f) : a1[i].x;
                out_s[o1 * kHP + rg * 8 + 2 * i + 1] = ACT == 1 ? fmaxf(a1[i].y, 0.0f) : a1[i].y;
            }
        }
    }
}

// Transposed product: din_s[k][r] (+)= sum_o dout_s[o][r] * WT_s[k * pitch + o]  for k < K.
// thread = (k pair: k0 = tid % 64 (+ 128 j), k0 + 64; row group of 8 rows): the same 2 x 8 register tile.
template <bool ACCUM>
__device__ __forceinline__ void dense_t(const float* __restrict__ dout_s, const float* __restrict__ wT_s, int pitch,
                                        float* __restrict__ din_s, int K, int outN) {
    const int rg = threadIdx.x >> 6;
    for (int k0 = threadIdx.x & 63; k0 < K; k0 += 128) {
        const bool two = k0 + 64 < K;
        const int k1 = two ? k0 + 64 : k0;
        float2 a0[4], a1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a0[i] = ACCUM ? make_float2(din_s[k0 * kHP + rg * 8 + 2 * i], din_s[k0 * kHP + rg * 8 + 2 * i + 1]) : make_float2(0.f, 0.f);
            a1[i] = ACCUM ? make_float2(din_s[k1 * kHP + rg * 8 + 2 * i], din_s[k1 * kHP + rg * 8 + 2 * i + 1]) : make_float2(0.f, 0.f);
        }
        const float* d = dout_s + rg * 8;
#pragma unroll 4
        for (int o = 0; o < outN; ++o) {
            const float w0 = wT_s[k0 * pitch + o], w1 = wT_s[k1 * pitch + o];
            const float2 p0 = make_float2(w0, w0), p1 = make_float2(w1, w1);
            const float4 xa = *reinterpret_cast<const float4*>(d + o * kHP);
            const float4 xb = *reinterpret_cast<const float4*>(d + o * kHP + 4);
            const float2 x0 = make_float2(xa.x, xa.y), x1 = make_float2(xa.z, xa.w), x2 = make_float2(xb.x, xb.y), x3 = make_float2(xb.z, xb.w);
            a0[0] = __ffma2_rn(p0, x0, a0[0]); a0[1] = __ffma2_rn(p0, x1, a0[1]);
            a0[2] = __ffma2_rn(p0, x2, a0[2]); a0[3] = __ffma2_rn(p0, x3, a0[3]);
            a1[0] = __ffma2_rn(p1, x0, a1[0]); a1[1] = __ffma2_rn(p1, x1, a1[1]);
            a1[2] = __ffma2_rn(p1, x2, a1[2]); a1[3] = __ffma2_rn(p1, x3, a1[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            din_s[k0 * kHP + rg * 8 + 2 * i] = a0[i].x;
            din_s[k0 * kHP + rg * 8 + 2 * i + 1] = a0[i].y;
            if (two) {
                din_s[k1 * kHP + rg * 8 + 2 * i] = a1[i].x;
                din_s[k1 * kHP + rg * 8 + 2 * i + 1] = a1[i].y;
            }
        }
    }
}

// dW[o][k] = sum_r dout_s[o][r] * in_s[k][r],  db[o] = sum_r dout_s[o][r]   -> global partials (torch layout (out, in)).
// thread = input feature k (its 32 rows in registers), the output units split over the thread groups of `K`-rounded size.
__device__ __forceinline__ void wgrad(const float* __restrict__ dout_s, const float* __restrict__ in_s, float* __restrict__ dW,
                                      float* __restrict__ db, int K, int outN) {
    const int kp = K <= 64 ? 64 : (K <= 128 ? 128 : 256);     // threads per group
    const int groups = kThreads / kp, grp = threadIdx.x / kp, k = threadIdx.x % kp;
    if (k < K) {
        float a[kRows];
#pragma unroll
        for (int q = 0; q < kRows / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(in_s + k * kHP + 4 * q);
            a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
        }
        for (int o = grp; o < outN; o += groups) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int q = 0; q < kRows / 4; ++q) {
                const float4 d = *reinterpret_cast<const float4*>(dout_s + o * kHP + 4 * q);
                s0 = fmaf(d.x, a[4 * q], s0);
                s1 = fmaf(d.y, a[4 * q + 1], s1);
                s0 = fmaf(d.z, a[4 * q + 2], s0);
                s1 = fmaf(d.w, a[4 * q + 3], s1);
            }
            dW[o * K + k] = s0 + s1;
        }
    }
    for (int o = threadIdx.x; o < outN; o += kThreads) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < kRows; ++r) s += dout_s[o * kHP + r];
        db[o] = s;
    }
}

__device__ __forceinline__ float keep_scale_p(unsigned int v) {
    const float uni = (float)(v >> 8) * (1.0f / 16777216.0f);
    return uni >= kDropP ? 1.0f / (1.0f - kDropP) : 0.0f;
}

// Body tile -> shared layer (h post ReLU/dropout in h_s, d h / d pre-activation in hs_s).  Leaves x_s, ws_s loaded.
__device__ __forceinline__ void shared_layer(const BiearHeadsParams& p, const float* const* __restrict__ wp, float* smem, const Smem& L,
                                             int s, int r0, unsigned long long seed) {
    float* x_s = smem + L.x();
    float* ws_s = smem + L.ws();
    float* h_s = smem + L.h();
    float* hs_s = smem + L.hs();
    float* b_s = smem + L.bias();
    const int D = p.D;
    {   // body tile, 128-bit loads (D % 4 == 0), 8 in flight per thread
        const int D4 = D >> 2, n4 = kRows * D4;
        const float4* b4 = reinterpret_cast<const float4*>(p.body + (long long)r0 * D);
        for (int base = threadIdx.x; base < n4; base += 8 * kThreads) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i4 = base + j * kThreads, r = i4 / D4;
                v[j] = (i4 < n4 && r0 + r < p.B) ? __ldg(b4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i4 = base + j * kThreads;
                if (i4 < n4) {
                    const int r = i4 / D4, k = 4 * (i4 - r * D4);
                    x_s[k * kHP + r] = v[j].x;
                    x_s[(k + 1) * kHP + r] = v[j].y;
                    x_s[(k + 2) * kHP + r] = v[j].z;
                    x_s[(k + 3) * kHP + r] = v[j].w;
                }
            }
        }
    }
    load_wT(ws_s, wp[0], kH1, D, odd(kH1));
    for (int i = threadIdx.x; i < kH1; i += kThreads) b_s[i] = __ldg(wp[1] + i);
    __syncthreads();
    dense<0>(x_s, ws_s, odd(kH1), b_s, h_s, D, kH1);
    __syncthreads();
    // ReLU + Dropout(0.2): thread = (unit quad, row); one Philox draw covers 4 units of one row
    for (int idx = threadIdx.x; idx < (kH1 / 4) * kRows; idx += kThreads) {
        const int uq = idx / kRows, r = idx - uq * kRows;
        float sc[4] = {1.f, 1.f, 1.f, 1.f};
        if (p.training) {
            const long long row = r0 + r;
            const uint4 rnd = philox4x32_10(make_uint4((unsigned)row, (unsigned)(row >> 32), 0x48454144u + (unsigned)s, (unsigned)uq),
                                            make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
            sc[0] = keep_scale_p(rnd.x); sc[1] = keep_scale_p(rnd.y); sc[2] = keep_scale_p(rnd.z); sc[3] = keep_scale_p(rnd.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int u = 4 * uq + j;
            const float pre = h_s[u * kHP + r];
            const float g = pre > 0.0f ? sc[j] : 0.0f;
            h_s[u * kHP + r] = pre * g;
            hs_s[u * kHP + r] = g;
        }
    }
    __syncthreads();
}

// One branch forward from h_s: a1_s, a2_s, o_s (pre-activation outputs) filled; the branch weights stay in shared memory.
__device__ __forceinline__ void branch_forward(const float* const* __restrict__ wp, float* smem, const Smem& L, int j, int K3) {
    float* b_s = smem + L.bias();
    const float* const* w = wp + 2 + 6 * j;
    load_wT(smem + L.w1(), w[0], kH2, kH1, odd(kH2));
    load_wT(smem + L.w2(), w[2], kH3, kH2, odd(kH3));
    load_wT(smem + L.w3(), w[4], K3, kH3, odd(kMaxC));
    for (int i = threadIdx.x; i < kH2; i += kThreads) b_s[kH1 + i] = __ldg(w[1] + i);
    for (int i = threadIdx.x; i < kH3; i += kThreads) b_s[kH1 + kH2 + i] = __ldg(w[3] + i);
    for (int i = threadIdx.x; i < K3; i += kThreads) b_s[kH1 + kH2 + kH3 + i] = __ldg(w[5] + i);
    __syncthreads();
    dense<1>(smem + L.h(), smem + L.w1(), odd(kH2), b_s + kH1, smem + L.a1(), kH1, kH2);
    __syncthreads();
    dense<1>(smem + L.a1(), smem + L.w2(), odd(kH3), b_s + kH1 + kH2, smem + L.a2(), kH2, kH3);
    __syncthreads();
    dense<0>(smem + L.a2(), smem + L.w3(), odd(kMaxC), b_s + kH1 + kH2 + kH3, smem + L.o(), kH3, K3);
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) heads_fwd_kernel(const BiearHeadsParams p) {
    extern __shared__ __align__(16) float smem[];
    const Smem L(p.D, p.C);
    const int s = blockIdx.x, r0 = blockIdx.y * kRows;
    const float* const* wp = p.wptr + (long long)s * kTensorsPerHead;
    const unsigned long long seed = p.seed_ptr ? *p.seed_ptr : p.seed;
    shared_layer(p, wp, smem, L, s, r0, seed);
    const float* o_s = smem + L.o();
    for (int j = 0; j < kBranches; ++j) {
        const int K3 = branch_out(j, p.C);
        branch_forward(wp, smem, L, j, K3);
        for (int idx = threadIdx.x; idx < K3 * kRows; idx += kThreads) {
            const int c = idx / kRows, r = idx - c * kRows;
            if (r0 + r >= p.B) continue;
            const float v = o_s[c * kHP + r];
            const long long row = r0 + r;
            if (j == 0) p.sound[row * p.S + s] = v;
            else if (j == 1) p.aoa[row * p.S + s] = 1.0f / (1.0f + expf(-v));
            else p.dist[(row * p.S + s) * p.C + c] = v;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads, 1) heads_bwd_kernel(const BiearHeadsParams p) {
    extern __shared__ __align__(16) float smem[];
    const Smem L(p.D, p.C);
    const int s = blockIdx.x, tile = blockIdx.y, r0 = tile * kRows;
    const int D = p.D, C = p.C;
    const float* const* wp = p.wptr + (long long)s * kTensorsPerHead;
    const unsigned long long seed = p.seed_ptr ? *p.seed_ptr : p.seed;
    float* gout = p.dw_part + ((long long)tile * p.S + s) * flat_floats(D, C);
    shared_layer(p, wp, smem, L, s, r0, seed);
    float* dh_s = smem + L.dh();
    float* o_s = smem + L.o();
    for (int i = threadIdx.x; i < kH1 * kHP; i += kThreads) dh_s[i] = 0.0f;
    for (int j = 0; j < kBranches; ++j) {
        const int K3 = branch_out(j, C);
        branch_forward(wp, smem, L, j, K3);
        // dL/d(output pre-activation) of this branch (rows beyond B: zero)
        for (int idx = threadIdx.x; idx < K3 * kRows; idx += kThreads) {
            const int c = idx / kRows, r = idx - c * kRows;
            const long long row = r0 + r;
            float g = 0.0f;
            if (row < p.B) {
                if (j == 0) g = p.g_sound ? __ldg(p.g_sound + row * p.S + s) : 0.0f;
                else if (j == 1) {
                    const float a = 1.0f / (1.0f + expf(-o_s[c * kHP + r]));
                    g = p.g_aoa ? __ldg(p.g_aoa + row * p.S + s) * a * (1.0f - a) : 0.0f;
                } else g = p.g_dist ? __ldg(p.g_dist + (row * p.S + s) * C + c) : 0.0f;
            }
            o_s[c * kHP + r] = g;
        }
        __syncthreads();
        float* gb = gout + flat_branch(j, D, C);
        float* gW1 = gb, *gb1 = gW1 + kH2 * kH1, *gW2 = gb1 + kH2, *gb2 = gW2 + kH3 * kH2, *gW3 = gb2 + kH3, *gb3 = gW3 + K3 * kH3;
        wgrad(o_s, smem + L.a2(), gW3, gb3, kH3, K3);
        dense_t<false>(o_s, smem + L.w3(), odd(kMaxC), smem + L.da2(), kH3, K3);
        __syncthreads();
        for (int i = threadIdx.x; i < kH3 * kHP; i += kThreads)                      // ReLU'
            if (!(smem[L.a2() + i] > 0.0f)) smem[L.da2() + i] = 0.0f;
        __syncthreads();
        wgrad(smem + L.da2(), smem + L.a1(), gW2, gb2, kH2, kH3);
        dense_t<false>(smem + L.da2(), smem + L.w2(), odd(kH3), smem + L.da1(), kH2, kH3);
        __syncthreads();
        for (int i = threadIdx.x; i < kH2 * kHP; i += kThreads)
            if (!(smem[L.a1() + i] > 0.0f)) smem[L.da1() + i] = 0.0f;
        __syncthreads();
        wgrad(smem + L.da1(), smem + L.h(), gW1, gb1, kH1, kH2);
        dense_t<true>(smem + L.da1(), smem + L.w1(), odd(kH2), dh_s, kH1, kH2);
        __syncthreads();
    }
    // through Dropout and ReLU of the shared layer, then its weight gradient and dL/dbody (this head's share)
    const float* hs_s = smem + L.hs();
    for (int i = threadIdx.x; i < kH1 * kHP; i += kThreads) dh_s[i] *= hs_s[i];
    __syncthreads();
    wgrad(dh_s, smem + L.x(), gout, gout + kH1 * D, D, kH1);
    float* dx_s = smem + L.h();                                   // h is dead: [D][kHP] needs D <= 2 * 100 rows of h | hs
    dense_t<false>(dh_s, smem + L.ws(), odd(kH1), dx_s, D, kH1);
    __syncthreads();
    float* dbody = p.d_body_part + (long long)s * p.B * D;
    for (int idx = threadIdx.x; idx < kRows * D; idx += kThreads) {
        const int r = idx / D, k = idx - r * D;
        if (r0 + r < p.B) dbody[(long long)(r0 + r) * D + k] = dx_s[k * kHP + r];
    }
}

// dst[i] = sum_{j < parts} src[j * n + i]  in the fixed order j = 0, 1, ...
__global__ void __launch_bounds__(256) heads_reduce_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n, int parts) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < parts; ++j) s += src[(long long)j * n + i];
        dst[i] = s;
    }
}

static int validate(const BiearHeadsParams* p, const char* who) {
    BIEAR_REQUIRE(p != nullptr, "%s: null parameter block", who);
    BIEAR_REQUIRE(p->B >= 1 && p->S >= 1 && p->D >= 4 && p->D <= 2 * kH1 && p->D % 4 == 0 && p->C >= 1 && p->C <= kMaxC,
                  "%s: bad geometry B=%d S=%d D=%d (4..200, multiple of 4) C=%d (1..8)", who, p->B, p->S, p->D, p->C);
    BIEAR_REQUIRE(p->body && p->wptr && p->sound && p->aoa && p->dist, "%s: null pointer", who);
    return 0;
}

template <typename Kern>
static int set_smem(Kern kern, size_t smem, const char* name) {
    BIEAR_REQUIRE(smem <= 227 * 1024, "%s needs %zu B of shared memory", name, smem);
    // per device and sticky: only the first launch (or a larger geometry) makes the attribute call
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> configured;
    int dev = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = configured[{dev, reinterpret_cast<const void*>(kern)}];
    if (have >= smem) return 0;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name)) return e;
    have = smem;
    return 0;
}

}  // namespace hd
}  // namespace biear

extern "C" int biear_heads_tile_rows(void) { return biear::hd::kRows; }
extern "C" int biear_heads_tensors_per_head(void) { return biear::hd::kTensorsPerHead; }
extern "C" int64_t biear_heads_flat_floats(int D, int C) {
    if (D < 1 || C < 1 || C > biear::hd::kMaxC) return -1;
    return biear::hd::flat_floats(D, C);
}

extern "C" int biear_heads_fwd(const BiearHeadsParams* p, void* stream) {
    using namespace biear;
    using namespace biear::hd;
    if (int e = validate(p, "biear_heads_fwd")) return e;
    const size_t smem = sizeof(float) * (size_t)Smem(p->D, p->C).total();
    if (int e = set_smem(heads_fwd_kernel, smem, "heads_fwd_kernel")) return e;
    heads_fwd_kernel<<<dim3(p->S, (p->B + kRows - 1) / kRows), kThreads, smem, as_stream(stream)>>>(*p);
    BIEAR_LAUNCH_CHECK("heads_fwd_kernel");
    return 0;
}

extern "C" int biear_heads_bwd(const BiearHeadsParams* p, void* stream) {
    using namespace biear;
    using namespace biear::hd;
    if (int e = validate(p, "biear_heads_bwd")) return e;
    BIEAR_REQUIRE(p->d_body_part && p->dw_part && p->d_body && p->dw, "biear_heads_bwd: null gradient buffer");
    const size_t smem = sizeof(float) * (size_t)Smem(p->D, p->C).total();
    if (int e = set_smem(heads_bwd_kernel, smem, "heads_bwd_kernel")) return e;
    const int tiles = (p->B + kRows - 1) / kRows;
    cudaStream_t st = as_stream(stream);
    heads_bwd_kernel<<<dim3(p->S, tiles), kThreads, smem, st>>>(*p);
    BIEAR_LAUNCH_CHECK("heads_bwd_kernel");
    const long long nw = (long long)p->S * flat_floats(p->D, p->C), nb = (long long)p->B * p->D;
    heads_reduce_kernel<<<(unsigned)((nw + 255) / 256 > 592 ? 592 : (nw + 255) / 256), 256, 0, st>>>(p->dw, p->dw_part, nw, tiles);
    BIEAR_LAUNCH_CHECK("heads_reduce_kernel");
    heads_reduce_kernel<<<(unsigned)((nb + 255) / 256 > 592 ? 592 : (nb + 255) / 256), 256, 0, st>>>(p->d_body, p->d_body_part, nb, p->S);
    BIEAR_LAUNCH_CHECK("heads_reduce_kernel");
    return 0;
}
