// One GRU layer's RECURRENCE as a persistent cluster kernel, forward and backward -- used for the wide first layer of
// the back-end's ILD / IPD encoders (model_torch.py:828-867: nn.GRU(100 -> 200) over the 19 frames), which cuDNN runs
// step by step (19 x (GEMM + cell kernel) forward, 19 x (cell-gradient kernel + GEMM) backward per encoder: ~0.55 ms of
// the 2.7 ms training step on the critical path, profiles/r2m launch list).
//
// The input projection gi = x W_ih^T + b_ih does not depend on the recurrence: the caller computes it for all frames as
// one library GEMM, and likewise the weight gradients / dL/dx from the per-step gate gradients this file writes.  What
// remains is the serial part:
//   forward   gh = W_hh h_{t-1} + b_hh;  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r gh_n),
//             h_t = (1 - z) n + z h_{t-1}                                               (torch.nn.GRU, gate order r, z, n)
//   backward  dgi = [dr', dz', dn'],  dgh = [dr', dz', dn' r],  dL/dh_{t-1} = z dL/dh_t + W_hh^T dgh
// Cluster = 4 CTAs x 512 threads = a tile of 16 rows for all T steps; CTA c owns hidden units [c H/4, (c+1) H/4): its slice
// of W_hh (forward: [k][gate][unit], backward: [gate*H + j][unit] = the transposed slice) stays in shared memory, h_t
// (forward) / the gate gradients (backward) are exchanged once per step with st.async + mbarrier transaction bytes
// (seq_dev.cuh), everything else stays in registers / shared memory.  thread = (k-half, row group of 4, unit); the two
// k-halves are summed through shared memory.
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "seq_dev.cuh"

namespace biear {
namespace gru {

constexpr int kThreads = kSeqThreads;      // 512 = 2 k-halves x 4 row groups x 64 unit slots
constexpr int kRowsT = 16;                 // rows per cluster
constexpr int kSlots = 64;                 // unit slots per CTA (H / 4 <= 64)
static_assert(2 * (kRowsT / kRT) * kSlots == kThreads, "thread layout");

__host__ __device__ constexpr int pitch_of(int HU) { return (HU + 3) & ~3; }
// workspace: [kCS forward images: [k<H][gate<3][UP]][kCS backward images: [o<3H][UP]]
__host__ __device__ constexpr long long fwd_img_floats_g(int H) { return (long long)H * 3 * pitch_of(H / kCS); }
__host__ __device__ constexpr long long bwd_img_floats_g(int H) { return (long long)3 * H * pitch_of(H / kCS); }
__host__ __device__ constexpr long long workspace_floats(int H) { return kCS * (fwd_img_floats_g(H) + bwd_img_floats_g(H)); }

struct FwdSmemG {   // floats
    int H, UP;
    __host__ __device__ FwdSmemG(int H_) : H(H_), UP(pitch_of(H_ / kCS)) {}
    __host__ __device__ int img() const { return 0; }
    __host__ __device__ int h() const { return H * 3 * UP; }                    // 2 x [H][16]
    __host__ __device__ int red() const { return h() + 2 * H * kRowsT; }        // 12 x 256
    __host__ __device__ int bars() const { return red() + 12 * 256; }
    __host__ __device__ int total() const { return bars() + 16; }
};
struct BwdSmemG {
    int H, UP;
    __host__ __device__ BwdSmemG(int H_) : H(H_), UP(pitch_of(H_ / kCS)) {}
    __host__ __device__ int img() const { return 0; }
    __host__ __device__ int x() const { return 3 * H * UP; }                    // 2 x [3H][16]: dr', dz', dn' r of all units
    __host__ __device__ int red() const { return x() + 2 * 3 * H * kRowsT; }    // 4 x 256
    __host__ __device__ int bars() const { return red() + 4 * 256; }
    __host__ __device__ int total() const { return bars() + 16; }
};

__global__ void __launch_bounds__(256) gru_pack_kernel(const BiearGruParams p, float* __restrict__ ws) {
    const int H = p.H, HU = H / kCS, UP = pitch_of(HU);
    const int c = blockIdx.x % kCS, which = blockIdx.x / kCS;          // which: 0 forward image, 1 backward image
    const int tid0 = blockIdx.y * blockDim.x + threadIdx.x, stride = gridDim.y * blockDim.x;
    if (which == 0) {
        float* out = ws + (long long)c * fwd_img_floats_g(H);
        for (int idx = tid0; idx < H * 3 * UP; idx += stride) {        // [k][gate][u]: consecutive threads -> consecutive floats
            const int u = idx % UP, kg = idx / UP, gate = kg % 3, k = kg / 3;
            out[idx] = u < HU ? p.w_hh[(long long)(gate * H + c * HU + u) * H + k] : 0.f;
        }
    } else {
        float* out = ws + kCS * fwd_img_floats_g(H) + (long long)c * bwd_img_floats_g(H);
        for (int idx = tid0; idx < 3 * H * UP; idx += stride) {        // [o][u] = W_hh[o][c HU + u]
            const int u = idx % UP, o = idx / UP;
            out[idx] = u < HU ? p.w_hh[(long long)o * H + c * HU + u] : 0.f;
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) gru_fwd_kernel(const BiearGruParams p, const float* __restrict__ ws) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int H = p.H, T = p.T, B = p.B, HU = H / kCS;
    const FwdSmemG L(H);
    const int UP = L.UP;
    float* img_s = smem + L.img();
    float* hbuf_s = smem + L.h();
    float* red_s = smem + L.red();
    const uint32_t bar = smem_u32(smem + L.bars());
    const int tid = threadIdx.x;
    const int ks = tid >> 8, rg = (tid >> 6) & 3, u = tid & (kSlots - 1);
    const bool active = u < HU;
    const int unit = rank * HU + u;                                   // global hidden unit of this thread
    const int b0 = (int)(blockIdx.x / kCS) * kRowsT;
    copy_f4(reinterpret_cast<float4*>(img_s), reinterpret_cast<const float4*>(ws + (long long)rank * fwd_img_floats_g(H)),
            (int)(fwd_img_floats_g(H) / 4));
    for (int i = tid; i < 2 * H * kRowsT; i += kThreads) hbuf_s[i] = 0.f;                  // h_{-1} = 0
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();
    float bhr = 0.f, bhz = 0.f, bhn = 0.f;
    if (active) {
        bhr = __ldg(p.b_hh + unit);
        bhz = __ldg(p.b_hh + H + unit);
        bhn = __ldg(p.b_hh + 2 * H + unit);
    }
    const int k0 = ks * (H / 2), k1 = k0 + H / 2;
    for (int t = 0; t < T; ++t) {
        const float* hcur = hbuf_s + (t & 1) * H * kRowsT;
        float* hnext = hbuf_s + ((t + 1) & 1) * H * kRowsT;
        // this step's input projections of (unit, 4 rows): issued before the products, used after them
        float gir[kRT] = {0.f, 0.f, 0.f, 0.f}, giz[kRT] = {0.f, 0.f, 0.f, 0.f}, gin[kRT] = {0.f, 0.f, 0.f, 0.f};
        if (ks == 0 && active) {
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int row = b0 + rg * kRT + i;
                if (row < B) {
                    const float* g = p.gi + ((long long)row * T + t) * 3 * H + unit;
                    gir[i] = __ldg(g);
                    giz[i] = __ldg(g + H);
                    gin[i] = __ldg(g + 2 * H);
                }
            }
        }
        float ar[kRT] = {0.f, 0.f, 0.f, 0.f}, az[kRT] = {0.f, 0.f, 0.f, 0.f}, an[kRT] = {0.f, 0.f, 0.f, 0.f};
        if (t > 0 && active) dot_rows3x(ar, az, an, hcur + rg * kRT, img_s + u, UP, k0, k1);
        __syncthreads();
        if (ks == 1) {
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                red_s[(i) * 256 + (tid & 255)] = ar[i];
                red_s[(4 + i) * 256 + (tid & 255)] = az[i];
                red_s[(8 + i) * 256 + (tid & 255)] = an[i];
            }
        }
        __syncthreads();
        if (ks == 0 && active) {
            float hv[kRT], vr[kRT], vz[kRT], vn[kRT], vh[kRT];
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const float sr = ar[i] + red_s[i * 256 + tid], sz = az[i] + red_s[(4 + i) * 256 + tid];
                vh[i] = an[i] + red_s[(8 + i) * 256 + tid] + bhn;
                vr[i] = 1.0f / (1.0f + expf(-(gir[i] + sr + bhr)));
                vz[i] = 1.0f / (1.0f + expf(-(giz[i] + sz + bhz)));
                vn[i] = tanhf(gin[i] + vr[i] * vh[i]);
                const float hp = hcur[unit * kRowsT + rg * kRT + i];
                hv[i] = (1.0f - vz[i]) * vn[i] + vz[i] * hp;
            }
            bcast_f4_tx(hnext + unit * kRowsT + rg * kRT, make_float4(hv[0], hv[1], hv[2], hv[3]), bar);
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int row = b0 + rg * kRT + i;
                if (row < B) {
                    const long long e = (long long)row * T + t;
                    p.h_seq[e * H + unit] = hv[i];
                    if (t + 1 < T) p.h_prev[(e + 1) * H + unit] = hv[i];      // the shifted copy the W_hh gradient GEMM reads
                    if (t == 0) p.h_prev[e * H + unit] = 0.f;
                    float* gt = p.gates + e * 4 * H + unit;
                    gt[0] = vr[i];
                    gt[H] = vz[i];
                    gt[2 * H] = vn[i];
                    gt[3 * H] = vh[i];
                }
            }
        }
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(H * kRowsT * 4));
        tx_wait(bar, (uint32_t)t & 1u);
    }
    __syncthreads();
    cluster.sync();   // no CTA leaves while a peer could still be sending to it
}

__global__ void __launch_bounds__(kThreads, 1) gru_bwd_kernel(const BiearGruParams p, const float* __restrict__ ws) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int H = p.H, T = p.T, B = p.B, HU = H / kCS;
    const BwdSmemG L(H);
    const int UP = L.UP;
    float* img_s = smem + L.img();
    float* xbuf_s = smem + L.x();
    float* red_s = smem + L.red();
    const uint32_t bar = smem_u32(smem + L.bars());
    const int tid = threadIdx.x;
    const int ks = tid >> 8, rg = (tid >> 6) & 3, u = tid & (kSlots - 1);
    const bool active = u < HU;
    const int unit = rank * HU + u;
    const int b0 = (int)(blockIdx.x / kCS) * kRowsT;
    copy_f4(reinterpret_cast<float4*>(img_s),
            reinterpret_cast<const float4*>(ws + kCS * fwd_img_floats_g(H) + (long long)rank * bwd_img_floats_g(H)),
            (int)(bwd_img_floats_g(H) / 4));
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();
    float carry[kRT] = {0.f, 0.f, 0.f, 0.f};          // W_hh^T dgh of step t+1 for (unit, 4 rows): ks == 0 threads
    float direct[kRT] = {0.f, 0.f, 0.f, 0.f};         // z_{t+1} dL/dh_{t+1}
    const int o0 = ks * (3 * H / 2), o1 = o0 + 3 * H / 2;
    for (int t = T - 1, it = 0; t >= 0; --t, ++it) {
        float* xcur = xbuf_s + (it & 1) * 3 * H * kRowsT;
        if (ks == 0 && active) {
            float d0[kRT], d1[kRT], d2[kRT];
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int row = b0 + rg * kRT + i;
                d0[i] = d1[i] = d2[i] = 0.f;
                float dirn = 0.f;
                if (row < B) {
                    const long long e = (long long)row * T + t;
                    const float dh = __ldg(p.dh_seq + e * H + unit) + carry[i] + direct[i];
                    const float* gt = p.gates + e * 4 * H + unit;
                    const float r = __ldg(gt), z = __ldg(gt + H), n = __ldg(gt + 2 * H), hn = __ldg(gt + 3 * H);
                    const float hp = __ldg(p.h_prev + e * H + unit);
                    const float dn = dh * (1.0f - z), dz = dh * (hp - n);
                    dirn = dh * z;
                    const float dnp = dn * (1.0f - n * n);
                    d0[i] = dnp * hn * r * (1.0f - r);      // dL/d r_pre
                    d1[i] = dz * z * (1.0f - z);            // dL/d z_pre
                    d2[i] = dnp * r;                         // dL/d (W_hn h + b_hn)
                    float* gi = p.dgi + e * 3 * H + unit;
                    float* gh = p.dgh + e * 3 * H + unit;
                    gi[0] = d0[i];
                    gi[H] = d1[i];
                    gi[2 * H] = dnp;
                    gh[0] = d0[i];
                    gh[H] = d1[i];
                    gh[2 * H] = d2[i];
                }
                direct[i] = dirn;
            }
            bcast_f4_tx(xcur + (0 * H + unit) * kRowsT + rg * kRT, make_float4(d0[0], d0[1], d0[2], d0[3]), bar);
            bcast_f4_tx(xcur + (1 * H + unit) * kRowsT + rg * kRT, make_float4(d1[0], d1[1], d1[2], d1[3]), bar);
            bcast_f4_tx(xcur + (2 * H + unit) * kRowsT + rg * kRT, make_float4(d2[0], d2[1], d2[2], d2[3]), bar);
        }
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(3 * H * kRowsT * 4));
        tx_wait(bar, (uint32_t)it & 1u);
        if (t == 0) break;                              // nothing upstream of h_{-1}
        // dL/dh_{t-1} through the recurrent weights: sum over the 3H gate rows of dgh * W_hh[:, unit]
        float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
        if (active) dot_rows_p(acc, xcur + rg * kRT, img_s + u, UP, o0, o1);
        __syncthreads();
        if (ks == 1) {
#pragma unroll
            for (int i = 0; i < kRT; ++i) red_s[i * 256 + (tid & 255)] = acc[i];
        }
        __syncthreads();
        if (ks == 0) {
#pragma unroll
            for (int i = 0; i < kRT; ++i) carry[i] = acc[i] + red_s[i * 256 + tid];
        }
    }
    __syncthreads();
    cluster.sync();
}

static int validate(const BiearGruParams* p, const char* who, bool backward) {
    BIEAR_REQUIRE(p != nullptr, "%s: null parameter block", who);
    BIEAR_REQUIRE(p->B >= 1 && p->T >= 1 && p->H >= 8 && p->H % 4 == 0 && p->H / kCS <= kSlots,
                  "%s: bad geometry B=%d T=%d H=%d (multiple of 4, <= 256)", who, p->B, p->T, p->H);
    BIEAR_REQUIRE(p->gi && p->w_hh && p->b_hh && p->h_seq && p->h_prev && p->gates && p->workspace, "%s: null pointer", who);
    if (backward) BIEAR_REQUIRE(p->dh_seq && p->dgi && p->dgh, "%s: null gradient pointer", who);
    return 0;
}

template <typename Kern>
static int launch(Kern kern, const char* name, int clusters, size_t smem, cudaStream_t st, const BiearGruParams& p) {
    BIEAR_REQUIRE(smem <= 227 * 1024, "%s: H=%d needs %zu B of shared memory", name, p.H, smem);
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> configured;
    int dev = 0;
    int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (e) return e;
    {
        std::lock_guard<std::mutex> lock(mu);
        size_t& have = configured[{dev, reinterpret_cast<const void*>(kern)}];
        if (have < smem) {
            e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
            if (e) return e;
            have = smem;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * kCS));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = check_cuda(cudaLaunchKernelEx(&cfg, kern, p, (const float*)p.workspace), name);
    if (e) return e;
    count_launch();
    return 0;
}

}  // namespace gru
}  // namespace biear

extern "C" int64_t biear_gru_workspace_floats(int H) {
    if (H < 8 || H % 4 || H / biear::kCS > biear::gru::kSlots) return -1;
    return biear::gru::workspace_floats(H);
}

extern "C" int biear_gru_supported(int H) {
    using namespace biear::gru;
    if (H < 8 || H % 4 || H / biear::kCS > kSlots) return 0;
    const size_t limit = 227 * 1024;
    return sizeof(float) * (size_t)FwdSmemG(H).total() <= limit && sizeof(float) * (size_t)BwdSmemG(H).total() <= limit;
}

extern "C" int biear_gru_fwd(const BiearGruParams* p, void* stream) {
    using namespace biear;
    using namespace biear::gru;
    if (int e = validate(p, "biear_gru_fwd", false)) return e;
    cudaStream_t st = as_stream(stream);
    gru_pack_kernel<<<dim3(2 * kCS, 8), 256, 0, st>>>(*p, p->workspace);
    BIEAR_LAUNCH_CHECK("gru_pack_kernel");
    return launch(gru_fwd_kernel, "gru_fwd_kernel", (p->B + kRowsT - 1) / kRowsT, sizeof(float) * (size_t)FwdSmemG(p->H).total(), st, *p);
}

extern "C" int biear_gru_bwd(const BiearGruParams* p, void* stream) {
    using namespace biear;
    using namespace biear::gru;
    if (int e = validate(p, "biear_gru_bwd", true)) return e;
    // (the workspace still holds the images the forward packed: the weights do not change between the two)
    return launch(gru_bwd_kernel, "gru_bwd_kernel", (p->B + kRowsT - 1) / kRowsT, sizeof(float) * (size_t)BwdSmemG(p->H).total(),
                  as_stream(stream), *p);
}
