// Fused Q-controller step kernels (forward and backward) and the per-sequence launch loops.
//
// Replaces, per frame t, model_torch.py:351-380 of the reference:
//   Y_ctrl = log1p(clamp(Y,0)); feat = [Y_ctrl, 0.2*Y_ctrl.detach()]; GRU(200->128) one step;
//   q_out = Linear-LayerNorm-SiLU-Dropout x2 + Linear(128->N); delta = tanh; Q = clamp(Q0 (1 + dQ delta)) or
//   clamp(Q0 + dQ delta); batch-global non-finite fallback
// (~25 library launches + 1 host sync per frame and ear there) by ONE cluster kernel per frame serving both
// ears, and what autograd derives from it in the backward pass by ONE kernel per frame that also applies
// the band stage's closed-form dQ and the log1p backward.  Weight gradients are NOT formed
// on the serial chain: the backward kernel stores the per-sample pre-activation gradients and the host
// turns them into dW with a handful of large GEMMs after the 18-step chain (ops.py).
//
// See ctrl_dev.cuh for the cluster/thread layout.
#include "band_dev.cuh"
#include "ctrl_dev.cuh"

namespace biear {

// band.cu
int launch_band_frame(const BiearSeqParams& p, int t, cudaStream_t st);

template <int CS>
struct FwdSmem {
    using Gm = CtrlGeom<CS>;
    int N, NU, Kc;      // bands, bands per CTA, controller contraction width (N for the dual front-end)
    int pi;             // pitch of the staged W_ih rows
    __host__ __device__ FwdSmem(int N_, int Kc_) : N(N_), NU((N_ + CS - 1) / CS), Kc(Kc_), pi(Kc_ | 1) {}
    static constexpr int ph = kHid + 1;
    __host__ __device__ int wih() const { return 0; }
    __host__ __device__ int whh() const { return wih() + 3 * Gm::U * pi; }
    __host__ __device__ int w1() const { return whh() + 3 * Gm::U * ph; }
    __host__ __device__ int w2() const { return w1() + Gm::U * ph; }
    __host__ __device__ int w3() const { return w2() + Gm::U * ph; }
    __host__ __device__ int yc() const { return (w3() + NU * ph + 3) & ~3; }
    __host__ __device__ int hprev() const { return yc() + Kc * Gm::R; }
    __host__ __device__ int hnew() const { return hprev() + kHid * Gm::R; }
    __host__ __device__ int a1() const { return hnew() + kHid * Gm::R; }
    __host__ __device__ int a2() const { return a1() + kHid * Gm::R; }
    __host__ __device__ int red() const { return a2() + kHid * Gm::R; }
    __host__ __device__ int stat() const { return red() + 16 * 128; }
    __host__ __device__ int total() const { return stat() + 2 * kCtrlThreads; }
};

__device__ __forceinline__ long long srow(const BiearSeqParams& p, int g, int t, int b) {
    return ((long long)(g * (p.T - 1) + t)) * p.B + b;
}

// LayerNorm + SiLU + Dropout over the full rows held in buf_s ([feature][R], overwritten in place with the
// layer output); every CTA of the cluster does this redundantly and saves only its own feature slice.
template <int CS>
__device__ __forceinline__ void ln_silu_drop_fwd(const BiearSeqParams& p, float* buf_s, float* stat_s,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 int layer, int g, int t, int b0, int rank, float* xh_out,
                                                 float* d_out, float* rstd_out) {
    using Gm = CtrlGeom<CS>;
    constexpr int R = Gm::R, PARTS = kCtrlThreads / R, FPP = kHid / PARTS;
    const int row = threadIdx.x % R, part = threadIdx.x / R;
    const int f0 = part * FPP;
    float v[FPP];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        v[i] = buf_s[(f0 + i) * R + row];
        s += v[i];
    }
    stat_s[part * R + row] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int q = 0; q < PARTS; ++q) mean += stat_s[q * R + row];
    mean *= (1.0f / kHid);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        v[i] -= mean;
        ss = fmaf(v[i], v[i], ss);
    }
    stat_s[kCtrlThreads + part * R + row] = ss;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int q = 0; q < PARTS; ++q) var += stat_s[kCtrlThreads + q * R + row];
    const float rstd = rsqrtf(var * (1.0f / kHid) + kLnEps);
    const int b = b0 + row;
    const bool valid = b < p.B;
    const long long s_idx = valid ? srow(p, g, t, b) : 0;
    const long long grow = (long long)g * p.B + b;
    if (valid && rank == 0 && part == 0) rstd_out[s_idx * 2 + layer] = rstd;
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        const int f = f0 + i;
        const float xh = v[i] * rstd;
        const float y = fmaf(xh, __ldg(gamma + f), __ldg(beta + f));
        float o = y / (1.0f + expf(-y));
        if (p.training) o *= dropout_scale(p.seed, t, layer, grow, f);
        buf_s[f * R + row] = o;
        if (valid && f / Gm::U == rank) {
            xh_out[s_idx * kHid + f] = xh;
            d_out[s_idx * kHid + f] = o;
        }
    }
    __syncthreads();
}

template <int CS>
__global__ void __launch_bounds__(kCtrlThreads, 1) ctrl_fwd_kernel(const BiearSeqParams p, const int t) {
    using Gm = CtrlGeom<CS>;
    constexpr int R = Gm::R, U = Gm::U;
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cl = blockIdx.x / CS;
    const int tiles = (p.B + R - 1) / R;
    const int g = cl / tiles;
    const int b0 = (cl % tiles) * R;
    const int N = p.N;
    const FwdSmem<CS> L(N, N);
    float* wih_s = smem + L.wih();
    float* whh_s = smem + L.whh();
    float* w1_s = smem + L.w1();
    float* w2_s = smem + L.w2();
    float* w3_s = smem + L.w3();
    float* yc_s = smem + L.yc();
    float* hprev_s = smem + L.hprev();
    float* hnew_s = smem + L.hnew();
    float* a1_s = smem + L.a1();
    float* a2_s = smem + L.a2();
    float* red_s = smem + L.red();
    float* stat_s = smem + L.stat();
    const int tid = threadIdx.x;
    const int ks = tid >> 7, slot = tid & 127, rg = slot / U, u = slot % U;
    const int ug = rank * U + u;                      // global hidden unit of this thread
    const int nu_c = max(0, min(L.NU, N - rank * L.NU));   // bands of the last layer owned by this CTA

    // ---- stage this CTA's weight slices (independent of the band kernel's output) ------------------
    {
        const float* w_ih = p.w_ih + (long long)g * 3 * kHid * p.Kin;
        for (int idx = tid; idx < 3 * U * N; idx += kCtrlThreads) {
            const int o = idx / N, k = idx - o * N;
            const int og = (o / U) * kHid + rank * U + (o % U);
            // feat = [yc, 0.2*yc.detach()]  =>  W_ih feat = (W_ih[:, :N] + 0.2 W_ih[:, N:]) yc
            wih_s[o * L.pi + k] = fmaf(0.2f, __ldg(w_ih + (long long)og * p.Kin + N + k), __ldg(w_ih + (long long)og * p.Kin + k));
        }
        const float* w_hh = p.w_hh + (long long)g * 3 * kHid * kHid;
        for (int idx = tid; idx < 3 * U * kHid; idx += kCtrlThreads) {
            const int o = idx >> 7, k = idx & 127;
            const int og = (o / U) * kHid + rank * U + (o % U);
            whh_s[o * L.ph + k] = __ldg(w_hh + og * kHid + k);
        }
        const float* w1 = p.w1 + (long long)g * kHid * kHid;
        const float* w2 = p.w2 + (long long)g * kHid * kHid;
        for (int idx = tid; idx < U * kHid; idx += kCtrlThreads) {
            const int o = idx >> 7, k = idx & 127;
            w1_s[o * L.ph + k] = __ldg(w1 + (rank * U + o) * kHid + k);
            w2_s[o * L.ph + k] = __ldg(w2 + (rank * U + o) * kHid + k);
        }
        const float* w3 = p.w3 + (long long)g * N * kHid;
        for (int idx = tid; idx < nu_c * kHid; idx += kCtrlThreads) {
            const int o = idx >> 7, k = idx & 127;
            w3_s[o * L.ph + k] = __ldg(w3 + (rank * L.NU + o) * kHid + k);
        }
    }
    // ---- features and previous hidden state of the cluster's R rows ---------------------------------
    const bool h_reset = (t == 0) || (p.flags[(t - 1) * p.G + g] != 0);
    for (int idx = tid; idx < N * R; idx += kCtrlThreads) {
        const int n = idx / R, r = idx - n * R;
        const int b = b0 + r;
        float v = 0.f;
        if (b < p.B) v = log1pf(fmaxf(p.Y[(((long long)g * p.B + b) * p.T + t) * N + n], 0.0f));
        yc_s[n * R + r] = v;
    }
    for (int idx = tid; idx < kHid * R; idx += kCtrlThreads) {
        const int k = idx / R, r = idx - k * R;
        const int b = b0 + r;
        float v = 0.f;
        if (!h_reset && b < p.B) v = p.H[srow(p, g, t - 1, b) * kHid + k];
        hprev_s[k * R + r] = v;
    }
    cluster.sync();   // every CTA of the cluster is running (DSMEM is live) and local staging is visible

    // ---- GRU cell ---------------------------------------------------------------------------------
    {
        float ar[kRT] = {0.f, 0.f, 0.f, 0.f}, az[kRT] = {0.f, 0.f, 0.f, 0.f};
        float ain[kRT] = {0.f, 0.f, 0.f, 0.f}, ahn[kRT] = {0.f, 0.f, 0.f, 0.f};
        const int kh = (N + 1) >> 1;
        const int k0 = ks ? kh : 0, k1 = ks ? N : kh;
        const float* x = yc_s + rg * kRT;
        dot_rows<R>(ar, x, wih_s + (0 * U + u) * L.pi, k0, k1);
        dot_rows<R>(az, x, wih_s + (1 * U + u) * L.pi, k0, k1);
        dot_rows<R>(ain, x, wih_s + (2 * U + u) * L.pi, k0, k1);
        const float* hx = hprev_s + rg * kRT;
        const int h0 = ks ? 64 : 0, h1 = ks ? 128 : 64;
        if (!h_reset) {
            dot_rows<R>(ar, hx, whh_s + (0 * U + u) * L.ph, h0, h1);
            dot_rows<R>(az, hx, whh_s + (1 * U + u) * L.ph, h0, h1);
            dot_rows<R>(ahn, hx, whh_s + (2 * U + u) * L.ph, h0, h1);
        }
        float acc[16];
#pragma unroll
        for (int i = 0; i < kRT; ++i) {
            acc[i] = ar[i];
            acc[4 + i] = az[i];
            acc[8 + i] = ain[i];
            acc[12 + i] = ahn[i];
        }
        reduce_halves<16>(acc, red_s, ks, slot);
        if (ks == 0) {
            const float* b_ih = p.b_ih + g * 3 * kHid;
            const float* b_hh = p.b_hh + g * 3 * kHid;
            const float br = __ldg(b_ih + ug) + __ldg(b_hh + ug);
            const float bz = __ldg(b_ih + kHid + ug) + __ldg(b_hh + kHid + ug);
            const float bin = __ldg(b_ih + 2 * kHid + ug), bhn = __ldg(b_hh + 2 * kHid + ug);
            float hv[kRT];
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int r = rg * kRT + i, b = b0 + r;
                const float rr = 1.0f / (1.0f + expf(-(acc[i] + br)));
                const float zz = 1.0f / (1.0f + expf(-(acc[4 + i] + bz)));
                const float hn = acc[12 + i] + bhn;
                const float nn = tanhf(acc[8 + i] + bin + rr * hn);
                const float hp = hprev_s[ug * R + r];
                hv[i] = (1.0f - zz) * nn + zz * hp;
                if (b < p.B) {
                    const long long s = srow(p, g, t, b);
                    p.H[s * kHid + ug] = hv[i];
                    float* gt = p.gates + s * 4 * kHid;
                    gt[ug] = rr;
                    gt[kHid + ug] = zz;
                    gt[2 * kHid + ug] = nn;
                    gt[3 * kHid + ug] = hn;
                }
            }
            broadcast_rows<CS, R>(cluster, hnew_s, ug, rg * kRT, hv);
        }
    }
    cluster.sync();

    // ---- Linear 1 -> LayerNorm -> SiLU -> Dropout ----------------------------------------------------
    {
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        dot_rows<R>(acc, hnew_s + rg * kRT, w1_s + u * L.ph, ks ? 64 : 0, ks ? 128 : 64);
        reduce_halves<kRT>(acc, red_s, ks, slot);
        if (ks == 0) {
            const float bb = __ldg(p.b1 + g * kHid + ug);
#pragma unroll
            for (int i = 0; i < kRT; ++i) acc[i] += bb;
            broadcast_rows<CS, R>(cluster, a1_s, ug, rg * kRT, acc);
        }
    }
    cluster.sync();
    ln_silu_drop_fwd<CS>(p, a1_s, stat_s, p.ln1_g + g * kHid, p.ln1_b + g * kHid, 0, g, t, b0, rank, p.xh1, p.d1, p.rstd);

    // ---- Linear 2 -> LayerNorm -> SiLU -> Dropout ----------------------------------------------------
    {
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        dot_rows<R>(acc, a1_s + rg * kRT, w2_s + u * L.ph, ks ? 64 : 0, ks ? 128 : 64);
        reduce_halves<kRT>(acc, red_s, ks, slot);
        if (ks == 0) {
            const float bb = __ldg(p.b2 + g * kHid + ug);
#pragma unroll
            for (int i = 0; i < kRT; ++i) acc[i] += bb;
            broadcast_rows<CS, R>(cluster, a2_s, ug, rg * kRT, acc);
        }
    }
    cluster.sync();
    ln_silu_drop_fwd<CS>(p, a2_s, stat_s, p.ln2_g + g * kHid, p.ln2_b + g * kHid, 1, g, t, b0, rank, p.xh2, p.d2, p.rstd);

    // ---- Linear 3 -> tanh -> Q update ---------------------------------------------------------------
    {
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        const bool mine = u < nu_c;
        if (mine) dot_rows<R>(acc, a2_s + rg * kRT, w3_s + u * L.ph, ks ? 64 : 0, ks ? 128 : 64);
        reduce_halves<kRT>(acc, red_s, ks, slot);
        if (ks == 0 && mine) {
            const int n = rank * L.NU + u;
            const float bb = __ldg(p.b3 + g * N + n);
            const float q0 = __ldg(p.q0 + n), dq = __ldg(p.dq + n);
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int b = b0 + rg * kRT + i;
                if (b >= p.B) continue;
                const float delta = tanhf(acc[i] + bb);
                const float qu = p.relative ? q0 * (1.0f + dq * delta) : fmaf(dq, delta, q0);
                if (!(fabsf(qu) <= 3.402823466e+38f)) atomicOr(p.flags + t * p.G + g, 1);   // NaN / Inf: batch-global fallback
                const float q = fminf(fmaxf(qu, p.q_min), p.q_max);
                p.Q[(((long long)g * p.B + b) * p.T + (t + 1)) * N + n] = q;
                p.delta[srow(p, g, t, b) * N + n] = delta;
            }
        }
    }
}

// ==================================================================================================
// backward
// ==================================================================================================
template <int CS>
struct BwdSmem {
    using Gm = CtrlGeom<CS>;
    int N, NU;
    __host__ __device__ BwdSmem(int N_) : N(N_), NU((N_ + CS - 1) / CS) {}
    __host__ __device__ int w3c() const { return 0; }                              // [n][U]
    __host__ __device__ int w2c() const { return w3c() + N * Gm::U; }              // [o][U]
    __host__ __device__ int w1c() const { return w2c() + kHid * Gm::U; }
    __host__ __device__ int whhc() const { return w1c() + kHid * Gm::U; }          // [o<384][U]
    __host__ __device__ int wihc() const { return whhc() + 3 * kHid * Gm::U; }     // [o<384][NU]
    __host__ __device__ int dpre() const { return (wihc() + 3 * kHid * NU + 3) & ~3; }   // [n][R]
    __host__ __device__ int bufa() const { return dpre() + N * Gm::R; }            // [128][R]
    __host__ __device__ int bufb() const { return bufa() + kHid * Gm::R; }
    __host__ __device__ int gate() const { return bufb() + kHid * Gm::R; }         // [4*128][R]: drp, dzp, dnp, dhn
    __host__ __device__ int red() const { return gate() + 4 * kHid * Gm::R; }
    __host__ __device__ int stat() const { return red() + 4 * 128; }
    __host__ __device__ int total() const { return stat() + 2 * kCtrlThreads; }
};

// Backward of Dropout -> SiLU -> LayerNorm on the full rows in buf_s (holds dL/d(layer output) on entry,
// dL/d(pre-LayerNorm activation) on exit).
template <int CS>
__device__ __forceinline__ void ln_silu_drop_bwd(const BiearSeqParams& p, float* buf_s, float* stat_s,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 int layer, int g, int t, int b0, int rank, const float* xh_in,
                                                 float* gv_out, float* ga_out) {
    using Gm = CtrlGeom<CS>;
    constexpr int R = Gm::R, PARTS = kCtrlThreads / R, FPP = kHid / PARTS;
    const int row = threadIdx.x % R, part = threadIdx.x / R;
    const int f0 = part * FPP;
    const int b = b0 + row;
    const bool valid = b < p.B;
    const long long s_idx = valid ? srow(p, g, t, b) : 0;
    const long long grow = (long long)g * p.B + b;
    float xh[FPP], dxh[FPP];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        const int f = f0 + i;
        xh[i] = valid ? xh_in[s_idx * kHid + f] : 0.f;
        const float gm = __ldg(gamma + f);
        const float y = fmaf(xh[i], gm, __ldg(beta + f));
        const float sg = 1.0f / (1.0f + expf(-y));
        float d = buf_s[f * R + row];
        if (p.training) d *= dropout_scale(p.seed, t, layer, grow, f);
        const float dv = d * sg * (1.0f + y * (1.0f - sg));
        if (valid && f / Gm::U == rank) gv_out[s_idx * kHid + f] = dv;
        dxh[i] = dv * gm;
        s1 += dxh[i];
        s2 = fmaf(dxh[i], xh[i], s2);
    }
    stat_s[part * R + row] = s1;
    stat_s[kCtrlThreads + part * R + row] = s2;
    __syncthreads();
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int q = 0; q < PARTS; ++q) {
        m1 += stat_s[q * R + row];
        m2 += stat_s[kCtrlThreads + q * R + row];
    }
    m1 *= (1.0f / kHid);
    m2 *= (1.0f / kHid);
    const float rstd = valid ? p.rstd[s_idx * 2 + layer] : 0.f;
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        const int f = f0 + i;
        const float da = rstd * (dxh[i] - m1 - xh[i] * m2);
        buf_s[f * R + row] = da;
        if (valid && f / Gm::U == rank) ga_out[s_idx * kHid + f] = da;
    }
    __syncthreads();
}

template <int CS>
__global__ void __launch_bounds__(kCtrlThreads, 1) ctrl_bwd_kernel(const BiearSeqParams p, const int t) {
    using Gm = CtrlGeom<CS>;
    constexpr int R = Gm::R, U = Gm::U;
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cl = blockIdx.x / CS;
    const int tiles = (p.B + R - 1) / R;
    const int g = cl / tiles;
    const int b0 = (cl % tiles) * R;
    const int N = p.N;
    const BwdSmem<CS> L(N);
    const int NU = L.NU;
    float* w3c_s = smem + L.w3c();
    float* w2c_s = smem + L.w2c();
    float* w1c_s = smem + L.w1c();
    float* whhc_s = smem + L.whhc();
    float* wihc_s = smem + L.wihc();
    float* dpre_s = smem + L.dpre();
    float* bufa_s = smem + L.bufa();
    float* bufb_s = smem + L.bufb();
    float* gate_s = smem + L.gate();
    float* red_s = smem + L.red();
    float* stat_s = smem + L.stat();
    const int tid = threadIdx.x;
    const int ks = tid >> 7, slot = tid & 127, rg = slot / U, u = slot % U;
    const int ug = rank * U + u;
    const int nu_c = max(0, min(NU, N - rank * NU));

    // ---- stage the column slices used by the transposed products -------------------------------------
    {
        const float* w3 = p.w3 + (long long)g * N * kHid;
        for (int idx = tid; idx < N * U; idx += kCtrlThreads) {
            const int o = idx / U, j = idx - o * U;
            w3c_s[idx] = __ldg(w3 + o * kHid + rank * U + j);
        }
        const float* w2 = p.w2 + (long long)g * kHid * kHid;
        const float* w1 = p.w1 + (long long)g * kHid * kHid;
        for (int idx = tid; idx < kHid * U; idx += kCtrlThreads) {
            const int o = idx / U, j = idx - o * U;
            w2c_s[idx] = __ldg(w2 + o * kHid + rank * U + j);
            w1c_s[idx] = __ldg(w1 + o * kHid + rank * U + j);
        }
        const float* w_hh = p.w_hh + (long long)g * 3 * kHid * kHid;
        for (int idx = tid; idx < 3 * kHid * U; idx += kCtrlThreads) {
            const int o = idx / U, j = idx - o * U;
            whhc_s[idx] = __ldg(w_hh + o * kHid + rank * U + j);
        }
        const float* w_ih = p.w_ih + (long long)g * 3 * kHid * p.Kin;
        for (int idx = tid; idx < 3 * kHid * NU; idx += kCtrlThreads) {
            const int o = idx / NU, j = idx - o * NU;
            // only the non-detached half of feat = [yc, 0.2*yc.detach()] carries gradient to yc
            wihc_s[idx] = j < nu_c ? __ldg(w_ih + (long long)o * p.Kin + rank * NU + j) : 0.f;
        }
    }
    // ---- dL/dQ_{t+1} (external + band-stage Jacobians) -> dL/dpre ------------------------------------------
    const bool flagged = p.flags[t * p.G + g] != 0;       // Q_{t+1} was replaced by Q0 and h_t reset
    const bool has_ctrl_next = (t + 1) < (p.T - 1);       // a controller step consumed Y_{t+1}
    for (int idx = tid; idx < N * R; idx += kCtrlThreads) {
        const int n = idx / R, r = idx - n * R;
        const int b = b0 + r;
        float dpre = 0.f;
        if (b < p.B && !flagged) {
            const long long row = (long long)g * p.B + b;
            const long long e = (row * p.T + (t + 1)) * N + n;
            float gy = p.gY ? p.gY[e] : 0.f;
            if (has_ctrl_next) gy += p.dYc[e];
            float d = gy * p.dYdQ[e];
            if (p.gP) d = fmaf(p.gP[e], p.dPdQ[e], d);
            if (p.gQ) d += p.gQ[e];
            const long long s = srow(p, g, t, b);
            const float delta = p.delta[s * N + n];
            const float q0 = __ldg(p.q0 + n), dq = __ldg(p.dq + n);
            const float qu = p.relative ? q0 * (1.0f + dq * delta) : fmaf(dq, delta, q0);
            const float scale = p.relative ? q0 * dq : dq;
            if (qu >= p.q_min && qu <= p.q_max) dpre = d * scale * (1.0f - delta * delta);
            if (n / NU == rank) p.G_pre[s * N + n] = dpre;
        } else if (b < p.B && n / NU == rank) {
            p.G_pre[srow(p, g, t, b) * N + n] = 0.f;
        }
        dpre_s[n * R + r] = dpre;
    }
    cluster.sync();

    // ---- Linear 3 ^T ------------------------------------------------------------------------------------
    {
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        const int nh = (N + 1) >> 1;
        dot_rows_strided<R>(acc, dpre_s + rg * kRT, w3c_s + u, U, ks ? nh : 0, ks ? N : nh);
        reduce_halves<kRT>(acc, red_s, ks, slot);
        if (ks == 0) broadcast_rows<CS, R>(cluster, bufa_s, ug, rg * kRT, acc);
    }
    cluster.sync();
    ln_silu_drop_bwd<CS>(p, bufa_s, stat_s, p.ln2_g + g * kHid, p.ln2_b + g * kHid, 1, g, t, b0, rank, p.xh2, p.G_v2, p.G_a2);
    // ---- Linear 2 ^T ------------------------------------------------------------------------------------
    {
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        dot_rows_strided<R>(acc, bufa_s + rg * kRT, w2c_s + u, U, ks ? 64 : 0, ks ? 128 : 64);
        reduce_halves<kRT>(acc, red_s, ks, slot);
        if (ks == 0) broadcast_rows<CS, R>(cluster, bufb_s, ug, rg * kRT, acc);
    }
    cluster.sync();
    ln_silu_drop_bwd<CS>(p, bufb_s, stat_s, p.ln1_g + g * kHid, p.ln1_b + g * kHid, 0, g, t, b0, rank, p.xh1, p.G_v1, p.G_a1);
    // ---- Linear 1 ^T, GRU cell backward -----------------------------------------------------------------------
    float dh_direct[kRT] = {0.f, 0.f, 0.f, 0.f};
    {
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        dot_rows_strided<R>(acc, bufb_s + rg * kRT, w1c_s + u, U, ks ? 64 : 0, ks ? 128 : 64);
        reduce_halves<kRT>(acc, red_s, ks, slot);
        if (ks == 0) {
            const bool h_reset = (t == 0) || (p.flags[(t - 1) * p.G + g] != 0);
            float v0[kRT], v1[kRT], v2[kRT], v3[kRT];
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int b = b0 + rg * kRT + i;
                v0[i] = v1[i] = v2[i] = v3[i] = 0.f;
                if (b >= p.B) continue;
                const long long s = srow(p, g, t, b);
                const long long row = (long long)g * p.B + b;
                float dh = acc[i];
                if (t < p.T - 2 && !flagged) dh += p.dH[row * kHid + ug];
                const float* gt = p.gates + s * 4 * kHid;
                const float rr = gt[ug], zz = gt[kHid + ug], nn = gt[2 * kHid + ug], hn = gt[3 * kHid + ug];
                const float hp = h_reset ? 0.f : p.H[srow(p, g, t - 1, b) * kHid + ug];
                const float dn = dh * (1.0f - zz);
                const float dz = dh * (hp - nn);
                dh_direct[i] = dh * zz;
                const float dnp = dn * (1.0f - nn * nn);
                const float dhn = dnp * rr;
                const float dr = dnp * hn;
                v0[i] = dr * rr * (1.0f - rr);
                v1[i] = dz * zz * (1.0f - zz);
                v2[i] = dnp;
                v3[i] = dhn;
                float* gg = p.GG + s * 4 * kHid;
                gg[ug] = v0[i];
                gg[kHid + ug] = v1[i];
                gg[2 * kHid + ug] = v2[i];
                gg[3 * kHid + ug] = v3[i];
            }
            broadcast_rows<CS, R>(cluster, gate_s, 0 * kHid + ug, rg * kRT, v0);
            broadcast_rows<CS, R>(cluster, gate_s, 1 * kHid + ug, rg * kRT, v1);
            broadcast_rows<CS, R>(cluster, gate_s, 2 * kHid + ug, rg * kRT, v2);
            broadcast_rows<CS, R>(cluster, gate_s, 3 * kHid + ug, rg * kRT, v3);
        }
    }
    cluster.sync();
    // ---- dL/dh_{t-1} = z * dh + W_hh^T [drp, dzp, dhn] ----------------------------------------------------------
    {
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        const float* x = gate_s + rg * kRT;
        // rows o of W_hh: [0,128) r gate, [128,256) z gate, [256,384) n gate (pairs with dhn = gate_s block 3)
        if (ks == 0) {
            dot_rows_strided<R>(acc, x, whhc_s + u, U, 0, 192);
        } else {
            dot_rows_strided<R>(acc, x, whhc_s + u, U, 192, 256);
            dot_rows_strided<R>(acc, x + (3 * kHid - 2 * kHid) * R, whhc_s + u, U, 256, 384);
        }
        reduce_halves<kRT>(acc, red_s, ks, slot);
        if (ks == 0 && t > 0) {
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int b = b0 + rg * kRT + i;
                if (b < p.B) p.dH[((long long)g * p.B + b) * kHid + ug] = acc[i] + dh_direct[i];
            }
        }
    }
    // ---- dL/dY_t through the controller: W_ih[:, :N]^T [drp, dzp, dnp] * d log1p ---------------------------------
    {
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        const bool mine = u < nu_c;
        if (mine) dot_rows_strided<R>(acc, gate_s + rg * kRT, wihc_s + u, NU, ks ? 192 : 0, ks ? 384 : 192);
        reduce_halves<kRT>(acc, red_s, ks, slot);
        if (ks == 0 && mine) {
            const int n = rank * NU + u;
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int b = b0 + rg * kRT + i;
                if (b >= p.B) continue;
                const long long e = ((((long long)g * p.B + b)) * p.T + t) * N + n;
                const float y = p.Y[e];
                p.dYc[e] = y >= 0.0f ? acc[i] / (1.0f + y) : 0.0f;
            }
        }
    }
}

// ==================================================================================================
// host side
// ==================================================================================================
template <int CS>
static int launch_ctrl(const BiearSeqParams& p, int t, bool backward, cudaStream_t st) {
    using Gm = CtrlGeom<CS>;
    const size_t smem = sizeof(float) * (size_t)(backward ? BwdSmem<CS>(p.N).total() : FwdSmem<CS>(p.N, p.N).total());
    BIEAR_REQUIRE(smem <= 227 * 1024, "controller step: N=%d needs %zu B of shared memory", p.N, smem);
    auto kern = backward ? ctrl_bwd_kernel<CS> : ctrl_fwd_kernel<CS>;
    static size_t configured[2] = {0, 0};
    if (configured[backward] < smem) {
        int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                           "cudaFuncSetAttribute(ctrl kernel smem)");
        if (e) return e;
        configured[backward] = smem;
    }
    const int tiles = (p.B + Gm::R - 1) / Gm::R;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.G * tiles * CS));
    cfg.blockDim = dim3(kCtrlThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int e = check_cuda(cudaLaunchKernelEx(&cfg, kern, p, t), backward ? "ctrl_bwd_kernel" : "ctrl_fwd_kernel");
    if (e) return e;
    count_launch();
    return 0;
}

static int validate_seq(const BiearSeqParams* p, const char* who, bool backward) {
    BIEAR_REQUIRE(p != nullptr, "%s: null parameter block", who);
    BIEAR_REQUIRE(p->G >= 1 && p->B >= 1 && p->T >= 1 && p->N >= 1 && p->N <= kHid && p->F >= 2,
                  "%s: bad geometry G=%d B=%d T=%d N=%d F=%d", who, p->G, p->B, p->T, p->N, p->F);
    BIEAR_REQUIRE(p->E == p->G && p->Kin == 2 * p->N,
                  "%s: only the dual front-end (one controller per ear, Kin = 2N) is fused; got E=%d G=%d Kin=%d",
                  who, p->E, p->G, p->Kin);
    BIEAR_REQUIRE(p->fc && p->q0 && p->dq && p->w_ih && p->w_hh && p->b_ih && p->b_hh && p->w1 && p->b1 && p->ln1_g &&
                      p->ln1_b && p->w2 && p->b2 && p->ln2_g && p->ln2_b && p->w3 && p->b3,
                  "%s: null constant / weight pointer", who);
    BIEAR_REQUIRE(p->Y && p->Q && p->dYdQ && p->H && p->gates && p->xh1 && p->d1 && p->xh2 && p->d2 && p->rstd &&
                      p->delta && p->flags,
                  "%s: null output / saved-state pointer", who);
    if (backward) {
        BIEAR_REQUIRE(p->dYc && p->dH && p->GG && p->G_a1 && p->G_v1 && p->G_a2 && p->G_v2 && p->G_pre,
                      "%s: null backward buffer", who);
        BIEAR_REQUIRE(!p->gP || p->dPdQ, "%s: gphase given but the forward saved no dphase/dQ", who);
    } else {
        BIEAR_REQUIRE(p->X, "%s: null spectra", who);
    }
    return 0;
}

}  // namespace biear

extern "C" int biear_ctrl_step_fwd(const BiearSeqParams* p, int t, void* stream) {
    using namespace biear;
    if (int e = validate_seq(p, "biear_ctrl_step_fwd", false)) return e;
    BIEAR_REQUIRE(t >= 0 && t < p->T - 1, "biear_ctrl_step_fwd: step %d outside [0,%d)", t, p->T - 1);
    return launch_ctrl<8>(*p, t, false, as_stream(stream));
}

extern "C" int biear_ctrl_step_bwd(const BiearSeqParams* p, int t, void* stream) {
    using namespace biear;
    if (int e = validate_seq(p, "biear_ctrl_step_bwd", true)) return e;
    BIEAR_REQUIRE(t >= 0 && t < p->T - 1, "biear_ctrl_step_bwd: step %d outside [0,%d)", t, p->T - 1);
    return launch_ctrl<8>(*p, t, true, as_stream(stream));
}

extern "C" int biear_adaptive_fwd(const BiearSeqParams* p, void* stream) {
    using namespace biear;
    if (int e = validate_seq(p, "biear_adaptive_fwd", false)) return e;
    cudaStream_t st = as_stream(stream);
    for (int t = 0; t < p->T; ++t) {
        if (int e = launch_band_frame(*p, t, st)) return e;
        if (t < p->T - 1)
            if (int e = launch_ctrl<8>(*p, t, false, st)) return e;
    }
    return 0;
}

extern "C" int biear_adaptive_bwd(const BiearSeqParams* p, void* stream) {
    using namespace biear;
    if (int e = validate_seq(p, "biear_adaptive_bwd", true)) return e;
    cudaStream_t st = as_stream(stream);
    for (int t = p->T - 2; t >= 0; --t)
        if (int e = launch_ctrl<8>(*p, t, true, st)) return e;
    return 0;
}
