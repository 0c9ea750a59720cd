// Q regularisers of the training loss, value and gradient in ONE launch.
//
// Replaces train_biear.py:476-490 applied to model.last_Q = (QL + QR) / 2 (model_torch.py:1076-1078):
//     logQ = log(Q + 1e-8);  logQ0 = log(Q0 + 1e-8)
//     reg_q      = mean((logQ - logQ0)^2)                       over (B, T, N)
//     reg_smooth = mean((logQ[..., 1:] - logQ[..., :-1])^2)     over (B, T, N-1)
//     loss      += REG_Q_W * reg_q + REG_SMOOTH_W * reg_smooth
// (~13 elementwise / reduction launches there and ~15 more in autograd's backward).  Here: one thread per element
// computes its three logarithms (its own and its two band neighbours'), its two squared terms and the closed-form
// gradient d loss / d Q; the sums go block -> partials -> the last block to finish adds the partials in a fixed order
// (deterministic for a given grid, no floating-point atomics).
#include "common.cuh"

namespace biear {

constexpr int kQregThreads = 256;
constexpr int kQregMaxBlocks = 16 * kSmCountB200;   // one element per thread at the benchmark size: one memory round trip

struct QRegArgs {
    const float* qa;
    const float* qb;      // nullable: Q = qa
    const float* q0;
    long long rows;
    int N;
    float w_reg, w_smooth;
    float* out;           // [3]: w_reg * reg_q + w_smooth * reg_smooth, reg_q, reg_smooth
    float* gq;            // nullable: d out[0] / d qa (== d out[0] / d qb)
    float* partials;      // [2 * gridDim.x]
    unsigned int* counter;
};

__device__ __forceinline__ float q_at(const QRegArgs& a, long long e) {
    return a.qb ? 0.5f * (__ldg(a.qa + e) + __ldg(a.qb + e)) : __ldg(a.qa + e);
}

__global__ void __launch_bounds__(kQregThreads) q_reg_kernel(const QRegArgs a) {
    __shared__ float red_s[2][kQregThreads / 32];
    __shared__ bool last_s;
    const long long total = a.rows * a.N;
    const float inv1 = 1.0f / (float)total;
    const float inv2 = 1.0f / (float)(a.rows * (a.N - 1));
    const float share = a.qb ? 0.5f : 1.0f;
    float s1 = 0.f, s2 = 0.f;
    for (long long e = (long long)blockIdx.x * kQregThreads + threadIdx.x; e < total;
         e += (long long)gridDim.x * kQregThreads) {
        const int n = (int)(e % a.N);
        const float qe = q_at(a, e) + 1e-8f;
        const float lq = logf(qe);
        const float d = lq - logf(__ldg(a.q0 + n) + 1e-8f);
        s1 = fmaf(d, d, s1);
        float gl = (2.0f * a.w_reg * inv1) * d;                  // d loss / d logQ[e]
        if (n + 1 < a.N) {
            const float dn = logf(q_at(a, e + 1) + 1e-8f) - lq;
            s2 = fmaf(dn, dn, s2);
            gl = fmaf(-2.0f * a.w_smooth * inv2, dn, gl);
        }
        if (n > 0) {
            const float dp = lq - logf(q_at(a, e - 1) + 1e-8f);
            gl = fmaf(2.0f * a.w_smooth * inv2, dp, gl);
        }
        if (a.gq) a.gq[e] = share * gl / qe;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red_s[0][warp] = s1;
        red_s[1][warp] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < kQregThreads / 32; ++w) {
            t1 += red_s[0][w];
            t2 += red_s[1][w];
        }
        a.partials[2 * blockIdx.x] = t1;
        a.partials[2 * blockIdx.x + 1] = t2;
        __threadfence();
        last_s = atomicAdd(a.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    // fixed order: thread-strided over the blocks (independent loads: one L2 round trip, not one per block), then the
    // same shuffle tree / warp order as above
    float t1 = 0.f, t2 = 0.f;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kQregThreads) {
        t1 += __ldcg(a.partials + 2 * b);
        t2 += __ldcg(a.partials + 2 * b + 1);
    }
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
    __syncthreads();                                             // red_s is reused
    if (lane == 0) {
        red_s[0][warp] = t1;
        red_s[1][warp] = t2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float u1 = 0.f, u2 = 0.f;
#pragma unroll
        for (int w = 0; w < kQregThreads / 32; ++w) {
            u1 += red_s[0][w];
            u2 += red_s[1][w];
        }
        const float r1 = u1 * inv1, r2 = u2 * inv2;
        a.out[0] = fmaf(a.w_reg, r1, a.w_smooth * r2);
        a.out[1] = r1;
        a.out[2] = r2;
        *a.counter = 0u;                                         // ready for the next launch on this stream
    }
}

}  // namespace biear

extern "C" int64_t biear_q_regularizers_workspace_floats(void) { return 2 * biear::kQregMaxBlocks + 4; }

extern "C" int biear_q_regularizers(const float* QA, const float* QB, const float* Q0, int64_t rows, int N, float w_reg,
                                    float w_smooth, float* out, float* gQ, float* workspace, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(QA && Q0 && out && workspace, "biear_q_regularizers: null pointer");
    BIEAR_REQUIRE(rows >= 1 && N >= 2, "biear_q_regularizers: need rows >= 1 and N >= 2, got rows=%lld N=%d", (long long)rows, N);
    QRegArgs a;
    a.qa = QA; a.qb = QB; a.q0 = Q0; a.rows = rows; a.N = N; a.w_reg = w_reg; a.w_smooth = w_smooth;
    a.out = out; a.gq = gQ;
    a.counter = reinterpret_cast<unsigned int*>(workspace);
    a.partials = workspace + 4;
    const long long total = rows * N;
    long long blocks = (total + kQregThreads - 1) / kQregThreads;
    if (blocks > kQregMaxBlocks) blocks = kQregMaxBlocks;
    q_reg_kernel<<<(int)blocks, kQregThreads, 0, as_stream(stream)>>>(a);
    BIEAR_LAUNCH_CHECK("q_reg_kernel");
    return 0;
}
