// One GRU layer's RECURRENCE as a persistent cluster kernel, forward and backward -- the two layers of the back-end's
// ILD / IPD encoders (model_torch.py:828-867: nn.GRU(100 -> 200), nn.GRU(200 -> 100) over the 19 frames), which the
// library runs step by step (19 x (GEMM + cell kernel) forward, 19 x (cell-gradient kernel + GEMM) backward per layer).
//
// The input projection gi = x W_ih^T + b_ih does not depend on the recurrence: the caller computes it for all frames as
// one library GEMM, and likewise the weight gradients / dL/dx from the per-step gate gradients this file writes.  What
// remains is the serial part:
//   forward   gh = W_hh h_{t-1} + b_hh;  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r gh_n),
//             h_t = (1 - z) n + z h_{t-1}                                               (torch.nn.GRU, gate order r, z, n)
//   backward  dgi = [dr', dz', dn'],  dgh = [dr', dz', dn' r],  dL/dh_{t-1} = z dL/dh_t + W_hh^T dgh
// Cluster = 4 CTAs x 512 threads = a tile of 16 rows for all T steps; CTA c owns hidden units [c H/4, (c+1) H/4) and keeps
// its part of W_hh in shared memory for all steps; one exchange between the CTAs per step (st.async + mbarrier transaction
// bytes, seq_dev.cuh); everything else stays in registers / shared memory.  See the two kernels for the thread layouts.
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "seq_dev.cuh"

namespace biear {
namespace gru {

constexpr int kThreads = kSeqThreads;      // 512
constexpr int kRowsT = 16;                 // rows per cluster
static_assert(kRowsT == kR && kRowsT / kRT == 4, "row groups of 4 rows");

__host__ __device__ constexpr int pitch_of(int HU) { return (HU + 3) & ~3; }
// forward: (unit, row group) pairs per CTA = (H / 4) * 4 = H; the contraction is split over as many thread sets as fit
__host__ __device__ constexpr int fwd_ks(int H) { return kThreads / H > 4 ? 4 : kThreads / H; }
// workspace: kCS forward images [k < H][gate < 3][UP] (the backward reads W_hh itself: its slices are contiguous rows)
__host__ __device__ constexpr long long fwd_img_floats_g(int H) { return (long long)H * 3 * pitch_of(H / kCS); }
__host__ __device__ constexpr long long workspace_floats(int H) { return kCS * fwd_img_floats_g(H); }

struct FwdSmemG {   // floats
    int H, UP;
    __host__ __device__ FwdSmemG(int H_) : H(H_), UP(pitch_of(H_ / kCS)) {}
    __host__ __device__ int img() const { return 0; }
    __host__ __device__ int h() const { return H * 3 * UP; }                            // 2 x [H][16]
    __host__ __device__ int red() const { return h() + 2 * H * kRowsT; }                // (KS - 1) x 12 x H
    __host__ __device__ int bars() const { return red() + (fwd_ks(H) - 1) * 12 * H; }  // 2 mbarriers
    __host__ __device__ int total() const { return bars() + 4; }
};
struct BwdSmemG {
    int H, HU;
    __host__ __device__ BwdSmemG(int H_) : H(H_), HU(H_ / kCS) {}
    __host__ __device__ int w() const { return 0; }                                     // [3 HU][H]: this CTA's rows of W_hh
    __host__ __device__ int d() const { return 3 * HU * H; }                            // [3 HU][16]: this CTA's gate gradients
    __host__ __device__ int part() const { return d() + 3 * HU * kRowsT; }              // 2 x [8 slots][4 row groups][HU][4]
    __host__ __device__ int bars() const { return part() + 2 * 8 * HU * kRowsT; }
    __host__ __device__ int total() const { return bars() + 4; }
};

__global__ void __launch_bounds__(256) gru_pack_kernel(const BiearGruParams p, float* __restrict__ ws) {
    const int H = p.H, HU = H / kCS, UP = pitch_of(HU);
    const int c = blockIdx.x;
    float* out = ws + (long long)c * fwd_img_floats_g(H);
    for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < H * 3 * UP; idx += gridDim.y * blockDim.x) {
        const int u = idx % UP, kg = idx / UP, gate = kg % 3, k = kg / 3;          // [k][gate][u]
        out[idx] = u < HU ? p.w_hh[(long long)(gate * H + c * HU + u) * H + k] : 0.f;
    }
}

// Forward.  CTA c owns hidden units [c HU, (c+1) HU): its slice of W_hh ([k][gate][unit]) stays in shared memory, h_t is
// broadcast to the peers once per step (st.async + mbarrier transaction bytes, two mbarriers used alternately).
// thread = (k-split, row group of 4, unit); the k-splits are summed through shared memory in a fixed order.
__global__ void __launch_bounds__(kThreads, 1) gru_fwd_kernel(const BiearGruParams p, const float* __restrict__ ws) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int H = p.H, T = p.T, B = p.B, HU = H / kCS;
    const FwdSmemG L(H);
    const int UP = L.UP, KS = fwd_ks(H);
    float* img_s = smem + L.img();
    float* hbuf_s = smem + L.h();
    float* red_s = smem + L.red();
    const uint32_t bar0 = smem_u32(smem + L.bars());
    const int tid = threadIdx.x;
    const int ks = tid / H, tile = tid - ks * H, rg = tile / HU, u = tile - rg * HU;
    const bool active = ks < KS, owner = ks == 0;
    const int unit = rank * HU + u;                                   // global hidden unit of this thread
    const int b0 = (int)(blockIdx.x / kCS) * kRowsT;
    copy_f4(reinterpret_cast<float4*>(img_s), reinterpret_cast<const float4*>(ws + (long long)rank * fwd_img_floats_g(H)),
            (int)(fwd_img_floats_g(H) / 4));
    for (int i = tid; i < 2 * H * kRowsT; i += kThreads) hbuf_s[i] = 0.f;                  // h_{-1} = 0
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();
    float bhr = 0.f, bhz = 0.f, bhn = 0.f;
    if (owner) {
        bhr = __ldg(p.b_hh + unit);
        bhz = __ldg(p.b_hh + H + unit);
        bhn = __ldg(p.b_hh + 2 * H + unit);
    }
    const int k0 = active ? (H * ks) / KS : 0, k1 = active ? (H * (ks + 1)) / KS : 0;
    for (int t = 0; t < T; ++t) {
        const float* hcur = hbuf_s + (t & 1) * H * kRowsT;
        float* hnext = hbuf_s + ((t + 1) & 1) * H * kRowsT;
        const uint32_t bar = bar0 + 8u * (uint32_t)(t & 1);
        // this step's input projections of (unit, 4 rows): issued before the products, used after them
        float gir[kRT] = {0.f, 0.f, 0.f, 0.f}, giz[kRT] = {0.f, 0.f, 0.f, 0.f}, gin[kRT] = {0.f, 0.f, 0.f, 0.f};
        if (owner) {
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
        if (KS > 1) {
            __syncthreads();                 // (the scratch of the previous step has been read)
            if (active && !owner) {
                float* r = red_s + (ks - 1) * 12 * H + tile;
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    r[i * H] = ar[i];
                    r[(4 + i) * H] = az[i];
                    r[(8 + i) * H] = an[i];
                }
            }
            __syncthreads();
        }
        if (owner) {
            for (int s = 1; s < KS; ++s) {
                const float* r = red_s + (s - 1) * 12 * H + tile;
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    ar[i] += r[i * H];
                    az[i] += r[(4 + i) * H];
                    an[i] += r[(8 + i) * H];
                }
            }
            float hv[kRT], vr[kRT], vz[kRT], vn[kRT], vh[kRT];
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                vh[i] = an[i] + bhn;
                vr[i] = 1.0f / (1.0f + expf(-(gir[i] + ar[i] + bhr)));
                vz[i] = 1.0f / (1.0f + expf(-(giz[i] + az[i] + bhz)));
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
        tx_wait(bar, (uint32_t)(t >> 1) & 1u);
    }
    __syncthreads();
    cluster.sync();   // no CTA leaves while a peer could still be sending to it
}

// Backward.  dL/dh_{t-1} = z dL/dh_t + W_hh^T dgh_t.  CTA c holds ITS rows of W_hh (gate g, units of c: contiguous rows of the
// parameter, all H columns) and the gate gradients of its own units, so the transposed product needs no gathered operand:
// every CTA forms the PARTIAL sums over its 3 HU gate rows for all H units (4 x 4 register tiles, two halves of the gate
// rows), and sends each partial to the CTA that owns the unit (st.async, 8 partial slots per receiver = 4 CTAs x 2 halves,
// double-buffered, two mbarriers used alternately); the owner adds the 8 partials in a fixed order.  Per step and CTA
// that moves 8 HU x 16 floats instead of broadcasting 3 HU x 16 gate gradients to four CTAs, and the products read
// 2 x 128 bits of shared memory per 16 FMAs.
__global__ void __launch_bounds__(kThreads, 1) gru_bwd_kernel(const BiearGruParams p) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int H = p.H, T = p.T, B = p.B, HU = H / kCS, O = 3 * HU;
    const BwdSmemG L(H);
    float* w_s = smem + L.w();
    float* d_s = smem + L.d();
    float* part_s = smem + L.part();
    const uint32_t bar0 = smem_u32(smem + L.bars());
    const int tid = threadIdx.x;
    const int ks = tid / H, tile = tid - ks * H, rg = tile / HU, u = tile - rg * HU;     // owner role: (unit u, rows 4 rg ..)
    const bool owner = ks == 0, worker = ks < 2;
    const int jg = u;                                   // product role: units 4 jg .. 4 jg + 3 of ALL H (H / 4 == HU groups)
    const int unit = rank * HU + u;
    const int b0 = (int)(blockIdx.x / kCS) * kRowsT;
    const int part_buf = 8 * HU * kRowsT;
#pragma unroll 1
    for (int g = 0; g < 3; ++g)
        copy_f4(reinterpret_cast<float4*>(w_s + g * HU * H),
                reinterpret_cast<const float4*>(p.w_hh + (long long)(g * H + rank * HU) * H), HU * H / 4);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();
    float direct[kRT] = {0.f, 0.f, 0.f, 0.f};         // z_{t+1} dL/dh_{t+1}
    float f_dh[kRT], f_r[kRT], f_z[kRT], f_n[kRT], f_hn[kRT], f_hp[kRT];      // the step's saved values, fetched one step ahead
    auto fetch = [&](int t) {
#pragma unroll
        for (int i = 0; i < kRT; ++i) {
            const int row = b0 + rg * kRT + i;
            f_dh[i] = f_r[i] = f_z[i] = f_n[i] = f_hn[i] = f_hp[i] = 0.f;
            if (row < B) {
                const long long e = (long long)row * T + t;
                f_dh[i] = __ldg(p.dh_seq + e * H + unit);
                const float* gt = p.gates + e * 4 * H + unit;
                f_r[i] = __ldg(gt);
                f_z[i] = __ldg(gt + H);
                f_n[i] = __ldg(gt + 2 * H);
                f_hn[i] = __ldg(gt + 3 * H);
                f_hp[i] = __ldg(p.h_prev + e * H + unit);
            }
        }
    };
    if (owner) fetch(T - 1);
    const int o0 = ks == 0 ? 0 : O / 2, o1 = ks == 0 ? O / 2 : O;
    for (int t = T - 1, it = 0; t >= 0; --t, ++it) {
        if (owner) {
            float carry[kRT] = {0.f, 0.f, 0.f, 0.f};   // W_hh^T dgh of step t+1 for (unit, 4 rows)
            if (it > 0) {
                const float* pb = part_s + ((it - 1) & 1) * part_buf + (rg * HU + u) * 4;
#pragma unroll
                for (int slot = 0; slot < 8; ++slot) {
                    const float4 v = *reinterpret_cast<const float4*>(pb + slot * 4 * HU * 4);
                    carry[0] += v.x; carry[1] += v.y; carry[2] += v.z; carry[3] += v.w;
                }
            }
            float d0[kRT], d1[kRT], d2[kRT];
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                const int row = b0 + rg * kRT + i;
                d0[i] = d1[i] = d2[i] = 0.f;
                float dirn = 0.f;
                if (row < B) {
                    const long long e = (long long)row * T + t;
                    const float dh = f_dh[i] + carry[i] + direct[i];
                    const float r = f_r[i], z = f_z[i], n = f_n[i], hn = f_hn[i], hp = f_hp[i];
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
            if (t > 0) {
                store4(d_s + (0 * HU + u) * kRowsT + rg * kRT, d0);
                store4(d_s + (1 * HU + u) * kRowsT + rg * kRT, d1);
                store4(d_s + (2 * HU + u) * kRowsT + rg * kRT, d2);
                fetch(t - 1);
            }
        }
        if (t == 0) break;                              // nothing upstream of h_{-1}
        __syncthreads();
        const uint32_t bar = bar0 + 8u * (uint32_t)(it & 1);
        if (worker) {
            // partial[row][j] = sum over this CTA's gate rows o in [o0, o1) of dgh[o][row] * W_hh[o][j], j = 4 jg .. 4 jg + 3
            float2 lo[4], hi[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) lo[i] = hi[i] = make_float2(0.f, 0.f);
            const float* xs = d_s + rg * kRT;
            const float* wsm = w_s + 4 * jg;
#pragma unroll 4
            for (int o = o0; o < o1; ++o) {
                const float4 x = *reinterpret_cast<const float4*>(xs + o * kRowsT);
                const float4 w = *reinterpret_cast<const float4*>(wsm + o * H);
                const float2 xl = make_float2(x.x, x.y), xh = make_float2(x.z, x.w);
                const float2 w0 = make_float2(w.x, w.x), w1 = make_float2(w.y, w.y), w2 = make_float2(w.z, w.z), w3 = make_float2(w.w, w.w);
                lo[0] = __ffma2_rn(w0, xl, lo[0]); hi[0] = __ffma2_rn(w0, xh, hi[0]);
                lo[1] = __ffma2_rn(w1, xl, lo[1]); hi[1] = __ffma2_rn(w1, xh, hi[1]);
                lo[2] = __ffma2_rn(w2, xl, lo[2]); hi[2] = __ffma2_rn(w2, xh, hi[2]);
                lo[3] = __ffma2_rn(w3, xl, lo[3]); hi[3] = __ffma2_rn(w3, xh, hi[3]);
            }
            const uint32_t slot_base = smem_u32(part_s + (it & 1) * part_buf + ((rank * 2 + ks) * 4 + rg) * HU * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 4 * jg + i, dest = j / HU, uu = j - dest * HU;
                st_async_f4(cluster_addr(slot_base + (uint32_t)uu * 16u, (uint32_t)dest),
                            make_float4(lo[i].x, lo[i].y, hi[i].x, hi[i].y), cluster_addr(bar, (uint32_t)dest));
            }
        }
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(part_buf * 4));
        tx_wait(bar, (uint32_t)(it >> 1) & 1u);
    }
    __syncthreads();
    cluster.sync();
}

static int validate(const BiearGruParams* p, const char* who, bool backward) {
    BIEAR_REQUIRE(p != nullptr, "%s: null parameter block", who);
    BIEAR_REQUIRE(p->B >= 1 && p->T >= 1 && p->H >= 8 && p->H % 4 == 0 && 2 * p->H <= kThreads,
                  "%s: bad geometry B=%d T=%d H=%d (multiple of 4, <= 232)", who, p->B, p->T, p->H);
    if (backward)
        BIEAR_REQUIRE(p->w_hh && p->h_prev && p->gates && p->dh_seq && p->dgi && p->dgh, "%s: null pointer", who);
    else
        BIEAR_REQUIRE(p->gi && p->w_hh && p->b_hh && p->h_seq && p->h_prev && p->gates && p->workspace, "%s: null pointer", who);
    return 0;
}

template <typename Kern, typename... Args>
static int launch(Kern kern, const char* name, int clusters, size_t smem, cudaStream_t st, const BiearGruParams& p, Args... args) {
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
    e = check_cuda(cudaLaunchKernelEx(&cfg, kern, p, args...), name);
    if (e) return e;
    count_launch();
    return 0;
}

}  // namespace gru
}  // namespace biear

extern "C" int64_t biear_gru_workspace_floats(int H) {
    if (!biear_gru_supported(H)) return -1;
    return biear::gru::workspace_floats(H);
}

extern "C" int biear_gru_supported(int H) {
    using namespace biear::gru;
    if (H < 8 || H % 4 || 2 * H > kThreads) return 0;
    const size_t limit = 227 * 1024;
    return sizeof(float) * (size_t)FwdSmemG(H).total() <= limit && sizeof(float) * (size_t)BwdSmemG(H).total() <= limit;
}

extern "C" int biear_gru_fwd(const BiearGruParams* p, void* stream) {
    using namespace biear;
    using namespace biear::gru;
    if (int e = validate(p, "biear_gru_fwd", false)) return e;
    cudaStream_t st = as_stream(stream);
    BIEAR_REQUIRE(biear_gru_supported(p->H), "biear_gru_fwd: H=%d does not fit in shared memory", p->H);
    gru_pack_kernel<<<dim3(kCS, 64), 256, 0, st>>>(*p, p->workspace);   // ~2 strided reads per thread: latency, not bandwidth
    BIEAR_LAUNCH_CHECK("gru_pack_kernel");
    return launch(gru_fwd_kernel, "gru_fwd_kernel", (p->B + kRowsT - 1) / kRowsT, sizeof(float) * (size_t)FwdSmemG(p->H).total(), st, *p,
                  (const float*)p->workspace);
}

extern "C" int biear_gru_bwd(const BiearGruParams* p, void* stream) {
    using namespace biear;
    using namespace biear::gru;
    if (int e = validate(p, "biear_gru_bwd", true)) return e;
    BIEAR_REQUIRE(biear_gru_supported(p->H), "biear_gru_bwd: H=%d does not fit in shared memory", p->H);
    return launch(gru_bwd_kernel, "gru_bwd_kernel", (p->B + kRowsT - 1) / kRowsT, sizeof(float) * (size_t)BwdSmemG(p->H).total(),
                  as_stream(stream), *p);
}
