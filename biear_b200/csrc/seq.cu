// Persistent adaptive-recurrence kernels: the whole 19-frame Q recurrence of the dual front-end, forward and
// backward, each as ONE cluster kernel.
//
// Replaces, for both ears at once, the frame loop of the reference
//   model_torch.py:333-380   per frame: W(Q_t) build + contraction -> Y_t;  Y_ctrl = log1p(clamp(Y,0));
//                            feat = [Y_ctrl, 0.2*Y_ctrl.detach()]; GRU(200->128) one step;
//                            q_out = Linear-LayerNorm-SiLU-Dropout x2 + Linear(128->N); delta = tanh;
//                            Q_{t+1} = clamp(Q0 (1 + dQ delta)) or clamp(Q0 + dQ delta); non-finite fallback
//   model_torch.py:1039-1063 the second W(Q_t) build for the sub-band phase
// (~60 library launches + 2 host syncs per frame there) and everything autograd derives from them.
//
// Forward:  cluster = 4 CTAs = one tile of 16 rows of one controller, resident for all T frames.  Per frame:
//   band stage (CTA c: rows 4c..4c+3; Y/phase/Jacobians -> HBM, log1p(Y) -> every CTA's shared memory)
//   -> GRU cell -> 2 x (Linear, LayerNorm, SiLU, Dropout) -> Linear -> tanh -> Q_{t+1} (-> HBM and -> the shared
//   memory of the CTA that runs the band stage of that row).  Five cluster barriers per frame, no HBM round trip.
// Backward: same cluster/tile ownership, t = T-2 .. 0: closed-form dL/dQ_{t+1} -> tanh/clamp ->
//   transposed Linear / LayerNorm / SiLU / Dropout chain -> GRU cell backward -> dL/dh_{t-1} (registers) and
//   dL/dY_t through the controller (shared memory of the owning CTA).  Weight gradients are NOT formed on the
//   serial chain: the per-sample pre-activation gradients are stored (tile layout) and biear_ctrl_wgrad turns
//   them into dW with split-K GEMMs afterwards.
//
// See seq_dev.cuh for the cluster/thread layout and the weight images.
#include <map>
#include <mutex>
#include <utility>

#include "band_dev.cuh"
#include "seq_dev.cuh"

namespace biear {

// Optional per-phase cycle accounting (diagnostic builds: -DBIEAR_PHASE_PROF, `make prof`): thread 0 of block 0 adds
// the clock64() time between consecutive marks to g_phase_cycles[kernel][phase]; read with biear_debug_phase_cycles.
#ifdef BIEAR_PHASE_PROF
__device__ unsigned long long g_phase_cycles[2][16];
#define PHASE_INIT() long long _ph_last = clock64()
#define PHASE_MARK(kern, i)                                            \
    do {                                                               \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                     \
            const long long _now = clock64();                          \
            g_phase_cycles[kern][i] += (unsigned long long)(_now - _ph_last); \
            _ph_last = _now;                                           \
        }                                                              \
    } while (0)
#else
#define PHASE_INIT() do {} while (0)
#define PHASE_MARK(kern, i) do {} while (0)
#endif

// ==================================================================================================
// weight images
// ==================================================================================================
constexpr int kPackSlices = 16;   // CTAs per image (grid.y)
static_assert(BIEAR_MAX_CTRL == 2, "ctrl_ptr selects between two controllers");
// controller g's tensor (a select, not an indexed read: indexing a kernel-parameter array spills it to local memory)
__device__ __forceinline__ const float* ctrl_ptr(const float* const (&a)[BIEAR_MAX_CTRL], int g) { return g ? a[1] : a[0]; }

// image of CTA rank c of controller g for the forward kernel (one slice of kPackSlices per CTA of the packing grid)
__device__ __forceinline__ void pack_fwd_image(const BiearSeqParams& p, float* __restrict__ out, int g, int c) {
    const int tid0 = blockIdx.y * blockDim.x + threadIdx.x, stride = gridDim.y * blockDim.x;
    const int N = p.N, NU = bands_per_cta(N);
    const float* w_ih = ctrl_ptr(p.w_ih, g);
    const float* w_hh = ctrl_ptr(p.w_hh, g);
    const float* w1 = ctrl_ptr(p.w1, g);
    const float* w2 = ctrl_ptr(p.w2, g);
    const float* w3 = ctrl_ptr(p.w3, g);
    // gate images: consecutive threads walk k (the contiguous axis of the torch layout) for coalesced reads
    for (int idx = tid0; idx < kU * 3 * N; idx += stride) {
        const int k = idx % N, gu = idx / N, gate = gu % 3, u = gu / 3;
        const long long o = gate * kHid + c * kU + u;
        out[fwd_img_wih(N) + (k * 3 + gate) * kU + u] = fmaf(0.2f, w_ih[o * p.Kin + N + k], w_ih[o * p.Kin + k]);
    }
    for (int idx = tid0; idx < kU * 3 * kHid; idx += stride) {
        const int k = idx % kHid, gu = idx / kHid, gate = gu % 3, u = gu / 3;
        const int o = gate * kHid + c * kU + u;
        out[fwd_img_whh(N) + (k * 3 + gate) * kU + u] = w_hh[o * kHid + k];
    }
    for (int idx = tid0; idx < kU * kHid; idx += stride) {
        const int k = idx % kHid, u = idx / kHid;
        out[fwd_img_w1(N) + k * kU + u] = w1[(c * kU + u) * kHid + k];
        out[fwd_img_w2(N) + k * kU + u] = w2[(c * kU + u) * kHid + k];
        const int n = c * NU + u;
        out[fwd_img_w3(N) + k * kU + u] = (u < NU && n < N) ? w3[n * kHid + k] : 0.f;
    }
}

__device__ __forceinline__ void pack_bwd_image(const BiearSeqParams& p, float* __restrict__ out, int g, int c) {
    const int tid0 = blockIdx.y * blockDim.x + threadIdx.x, stride = gridDim.y * blockDim.x;
    const int N = p.N, NU = bands_per_cta(N);
    const float* w_ih = ctrl_ptr(p.w_ih, g);
    const float* w_hh = ctrl_ptr(p.w_hh, g);
    const float* w1 = ctrl_ptr(p.w1, g);
    const float* w2 = ctrl_ptr(p.w2, g);
    const float* w3 = ctrl_ptr(p.w3, g);
    for (int idx = tid0; idx < N * kU; idx += stride)
        out[bwd_img_w3c(N) + idx] = w3[(idx / kU) * kHid + c * kU + (idx % kU)];
    for (int idx = tid0; idx < kHid * kU; idx += stride) {
        const int o = idx / kU, u = idx % kU;
        out[bwd_img_w2c(N) + idx] = w2[o * kHid + c * kU + u];
        out[bwd_img_w1c(N) + idx] = w1[o * kHid + c * kU + u];
    }
    for (int idx = tid0; idx < 3 * kHid * kU; idx += stride) {
        const int o = idx / kU, u = idx % kU;
        out[bwd_img_whhc(N) + idx] = w_hh[o * kHid + c * kU + u];
        const int n = c * NU + u;
        out[bwd_img_wihc(N) + idx] = (u < NU && n < N) ? w_ih[(long long)o * p.Kin + n] : 0.f;
    }
}

// Workspace layout: [G * kCS forward images][G * kCS backward images].
__host__ __device__ inline long long bwd_images_offset(int G, int N) { return (long long)G * kCS * fwd_img_floats(N); }

// Everything the recurrence needs that does not depend on the spectra, in ONE launch (so that a caller can run it on
// a forked stream next to the STFT): both sets of weight images, the zero initial GRU state H[:, 0] and the cleared
// fallback flags.  want: bit 0 forward images + H0 + flags, bit 1 backward images.
__global__ void __launch_bounds__(256) prepare_kernel(const BiearSeqParams p, float* __restrict__ ws, int want) {
    const int n_img = p.G * kCS;
    const int bx = blockIdx.x;
    if (bx < n_img) {
        if (want & 1) pack_fwd_image(p, ws + (long long)bx * fwd_img_floats(p.N), bx / kCS, bx % kCS);
    } else if (bx < 2 * n_img) {
        const int b = bx - n_img;
        if (want & 2) pack_bwd_image(p, ws + bwd_images_offset(p.G, p.N) + (long long)b * bwd_img_floats(p.N), b / kCS, b % kCS);
    } else if (want & 1) {
        const int tid0 = blockIdx.y * blockDim.x + threadIdx.x, stride = gridDim.y * blockDim.x;
        const int tiles = (p.B + kR - 1) / kR;
        const long long per_g = (long long)tiles * kHid * kR / 4;
        for (int g = 0; g < p.G; ++g) {
            float4* h0 = reinterpret_cast<float4*>(p.H + (long long)g * p.T * tiles * kHid * kR);
            for (long long i = tid0; i < per_g; i += stride) h0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int i = tid0; i < (p.T - 1) * p.G + 1; i += stride) p.flags[i] = 0;
        if (p.x_ready)
            for (long long i = tid0; i < (long long)p.E * p.B * p.T + 4; i += stride) p.x_ready[i] = 0;   // + work counter
    }
}

// ==================================================================================================
// forward
// ==================================================================================================
struct FwdSmem {   // offsets in floats
    int N, tile;
    __host__ __device__ FwdSmem(int N_, int F) : N(N_), tile(spec_tile_len(F)) {}
    __host__ __device__ int img() const { return 0; }
    __host__ __device__ int vec() const { return fwd_img_floats(N); }           // per-CTA constants, see V_*
    __host__ __device__ int yc() const { return vec() + 1088; }                 // [128][kR]; aliased by a2
    __host__ __device__ int a1() const { return yc() + kHid * kR; }             // [128][kR], directly behind yc
    __host__ __device__ int h0() const { return a1() + kHid * kR; }             // [128][kR] x 2 (ping-pong)
    __host__ __device__ int stat() const { return h0() + 2 * kHid * kR; }       // 2 x kSeqThreads (LayerNorm partials)
    __host__ __device__ int q() const { return stat() + 2 * kSeqThreads; }      // [128][4]: Q_t of this CTA's 4 rows
    __host__ __device__ int ystage() const { return q() + kHid * kRT; }         // [4][128]
    __host__ __device__ int spec() const { return ystage() + kRT * kHid; }      // [4][tile] float4
    __host__ __device__ int misc() const { return spec() + kRT * tile * 4; }   // 5 mbarriers (8 B each), 16-byte aligned
    __host__ __device__ int total() const { return misc() + 16; }
};
// vec area
constexpr int V_BR = 0, V_BZ = kU, V_BIN = 2 * kU, V_BHN = 3 * kU, V_B1 = 4 * kU, V_B2 = 5 * kU, V_B3 = 6 * kU,
              V_Q0S = 7 * kU, V_DQS = 8 * kU;
constexpr int V_LN1G = 9 * kU, V_LN1B = V_LN1G + kHid, V_LN2G = V_LN1B + kHid, V_LN2B = V_LN2G + kHid, V_FC = V_LN2B + kHid,
              V_Q0 = V_FC + kHid, V_END = V_Q0 + kHid;
static_assert(V_END <= 1088, "vec area");

__device__ __forceinline__ long long tile_base(const BiearSeqParams& p, int g, int t, int tiles, int tile) {
    return ((long long)(g * (p.T - 1) + t)) * tiles + tile;
}
// H is (G, T, tiles, 128, 32): index 0 is the all-zero initial state, h_t sits at step index t + 1.
__device__ __forceinline__ float* h_tile(const BiearSeqParams& p, int g, int t, int tiles, int tile) {
    return p.H + (((long long)g * p.T + (t + 1)) * tiles + tile) * (kHid * kR);
}

// LayerNorm + SiLU + Dropout over the full rows held in buf_s ([feature][32], overwritten in place with the layer
// output).  Every CTA of the cluster does this redundantly (the next layer needs all 128 features everywhere);
// only the warp whose 16 features are the CTA's own slice saves them.
__device__ __forceinline__ void ln_silu_drop_fwd(const BiearSeqParams& p, unsigned long long seed, float* buf_s, float* stat_s,
                                                 const float* __restrict__ gamma_s, const float* __restrict__ beta_s,
                                                 int layer, int t, long long grow0, int rank, float* xh_tile,
                                                 float* d_tile, float* rstd_tile) {
    constexpr int PARTS = kSeqThreads / kR, FPP = kHid / PARTS;     // 32 parts of 4 features
    static_assert(kR == 16 && PARTS == 2 * (kSeqThreads / 32), "lanes l and l^16 of a warp hold two parts of one row");
    const int row = threadIdx.x % kR, part = threadIdx.x / kR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f0 = part * FPP;
    // One pass over the row: sum and sum of squares of x - pivot (pivot = the row's feature 0; the shift keeps
    // E[x^2] - E[x]^2 well conditioned), the two parts of a warp combined by a shuffle, the 16 warps through shared memory.
    const float pivot = buf_s[row];
    float v[FPP];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        v[i] = buf_s[(f0 + i) * kR + row] - pivot;
        s += v[i];
        ss = fmaf(v[i], v[i], ss);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    ss += __shfl_xor_sync(0xffffffffu, ss, 16);
    float2* stat2 = reinterpret_cast<float2*>(stat_s);               // [row][17] float2: conflict-free both ways
    if (lane < kR) stat2[row * 17 + warp] = make_float2(s, ss);
    __syncthreads();
    float S = 0.f, SS = 0.f;
#pragma unroll
    for (int w = 0; w < kSeqThreads / 32; ++w) {
        const float2 q = stat2[row * 17 + w];
        S += q.x;
        SS += q.y;
    }
    const float mean = S * (1.0f / kHid);
    const float var = fmaxf(fmaf(-mean, mean, SS * (1.0f / kHid)), 0.0f) * kHid;   // sum of squared deviations
#pragma unroll
    for (int i = 0; i < FPP; ++i) v[i] -= mean;
    const float rstd = rsqrtf(var * (1.0f / kHid) + kLnEps);
    const bool mine = f0 / kU == rank;
    if (rank == 0 && part == 0) rstd_tile[layer * kR + row] = rstd;
#pragma unroll
    for (int i4 = 0; i4 < FPP / 4; ++i4) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.training) sc = dropout_scale4(seed, t, layer, grow0 + row, (f0 >> 2) + i4);
        const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i4 * 4 + j, f = f0 + i;
            const float xh = v[i] * rstd;
            const float y = fmaf(xh, gamma_s[f], beta_s[f]);
            const float o = (y * sigmoid_fast(y)) * scv[j];
            buf_s[f * kR + row] = o;
            if (mine) {
                xh_tile[f * kR + row] = xh;
                d_tile[f * kR + row] = o;
            }
        }
    }
    __syncthreads();
}

// Streaming hand-over of the spectra (BiearSeqParams.x_ready): wait until the concurrently running STFT has published
// frame t of this CTA's rows.  One thread per row polls (acquire), the block barrier hands the ordering to everybody.
// The spin is bounded: a broken protocol must end in a launch failure, never in a hung GPU.
__device__ __forceinline__ void wait_spectra(const BiearSeqParams& p, long long grow0, int b0, int t) {
    if (!p.x_ready) return;
    if (threadIdx.x < kRT && b0 + (int)threadIdx.x < p.B) {
        const int32_t* flag = p.x_ready + (grow0 + threadIdx.x) * p.T + t;
        int v = 0;
        for (unsigned spin = 0; ; ++spin) {
            asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v != 0) break;
            if (spin > (1u << 24)) __trap();
            __nanosleep(64);
        }
    }
    __syncthreads();
}

// Spectra of this CTA's 4 rows for frame t -> shared {1, abs, re, im} tiles (zeros for padding rows / bins), in two
// steps so that the HBM latency hides behind the controller phases: prefetch_spectra() issues 8-byte cp.async copies
// of the raw complex bins straight into the {re, im} half of their tile slots; finish_spectra() (same thread -> same
// slots) waits for them and fills in {1, abs}.
__device__ __forceinline__ void prefetch_spectra(const BiearSeqParams& p, float4* spec_s, int tile, long long grow0,
                                                 int b0, int t) {
#pragma unroll
    for (int i = 0; i < kRT; ++i) {
        const float2* src = reinterpret_cast<const float2*>(p.X) + ((grow0 + i) * p.T + t) * p.F;
        const bool row_ok = b0 + i < p.B;
        for (int k = threadIdx.x; k < tile; k += kSeqThreads) {
            float4* slot = spec_s + i * tile + k;
            if (row_ok && k < p.F) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&slot->z);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + k) : "memory");
            } else {
                *slot = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void finish_spectra(const BiearSeqParams& p, float4* spec_s, int tile, int b0) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int i = 0; i < kRT; ++i) {
        if (b0 + i >= p.B) continue;
        for (int k = threadIdx.x; k < p.F; k += kSeqThreads) {
            float4* slot = spec_s + i * tile + k;
            *slot = spec_entry(make_float2(slot->z, slot->w));
        }
    }
}

// The STRICT pass: ONE cluster walks all tiles frame by frame, one chain of 16 rows, with the state going through global
// memory and hardware cluster barriers between the phases -- which makes the reference's batch-global non-finite
// fallback (model_torch.py:378-380) exact.  It exits at once unless the fast pass (seq_fwd2_kernel below) recorded a
// non-finite Q (or p.force_strict).  Same arithmetic per row as the fast pass.
__global__ void __launch_bounds__(kSeqThreads, 1) seq_fwd_strict_kernel(const BiearSeqParams p, const float* __restrict__ img) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int N = p.N, T = p.T, S = p.T - 1;
    const int tiles = (p.B + kR - 1) / kR;
    const int n_tiles = p.G * tiles;
    const int NU = bands_per_cta(N);
    const FwdSmem L(N, p.F);
    float* img_s = smem + L.img();
    float* vec_s = smem + L.vec();
    float* yc_s = smem + L.yc();
    float* a2_s = yc_s;
    float* hbuf_s = smem + L.h0();
    float* a1_s = smem + L.a1();
    float* stat_s = smem + L.stat();
    // Scratch of the k-split reductions lives in activation buffers that are idle at that point of the frame (this is
    // what lets N = 128 fit in 227 KB): the GRU's 16 accumulators park in yc|a1 (both dead once every local dot
    // product has passed the reduction's first barrier; peers write them only after #3 / the next #2), Linear 1 in
    // yc (a2, its alias, is written after #3), Linear 2 and 3 in a1 (next written after the next frame's #2).
    float* red_gru_s = yc_s;
    float* red_l1_s = yc_s;
    float* red_l23_s = a1_s;
    static_assert(2 * 16 * 128 <= 2 * kHid * kR && (kKS - 1) * kRT * 128 <= kHid * kR, "reduction scratch fits its hosts");
    float* q_s = smem + L.q();
    float* ystage_s = smem + L.ystage();
    float4* spec_s = reinterpret_cast<float4*>(smem + L.spec());
    const int tid = threadIdx.x, lane = tid & 31;
    const int ks = tid >> 7, slot = tid & 127, rg = slot / kU, u = slot % kU;
    const int ug = rank * kU + u;
    const int nu_c = max(0, min(NU, N - rank * NU));
    const int quads = (N + 3) >> 2;
    int* any_flag = p.flags + S * p.G;
    const unsigned long long seed = p.seed_ptr ? *p.seed_ptr : p.seed;
    if (!p.force_strict && *reinterpret_cast<volatile int*>(any_flag) == 0) return;   // uniform over the grid
    cluster.sync();                                       // everyone has read the old any-flag
    for (int i = rank * kSeqThreads + tid; i < S * p.G; i += kCS * kSeqThreads) p.flags[i] = 0;
    __threadfence();
    cluster.sync();
    const int tl_begin = 0, tl_end = n_tiles;
    int cur_g = -1;
    constexpr int hsel = 0;        // h_{t-1} is reloaded from global memory every frame: the two buffers keep their roles

    for (int t = 0; t < T; ++t) {
        for (int tl = tl_begin; tl < tl_end; ++tl) {
            const int g = tl / tiles, tile = tl % tiles;
            const int b0 = tile * kR;
            const long long grow0 = (long long)g * p.B + b0 + rank * kRT;   // global row of this CTA's first band-stage row
            const int bb0 = b0 + rank * kRT;                                // its clip index
            if (g != cur_g) {   // (re)load this CTA's weight image and constants
                __syncthreads();
                copy_f4(reinterpret_cast<float4*>(img_s),
                        reinterpret_cast<const float4*>(img + (long long)(g * kCS + rank) * fwd_img_floats(N)),
                        fwd_img_floats(N) / 4);
                for (int i = tid; i < kHid; i += kSeqThreads) {
                    vec_s[V_LN1G + i] = ctrl_ptr(p.ln1_g, g)[i];
                    vec_s[V_LN1B + i] = ctrl_ptr(p.ln1_b, g)[i];
                    vec_s[V_LN2G + i] = ctrl_ptr(p.ln2_g, g)[i];
                    vec_s[V_LN2B + i] = ctrl_ptr(p.ln2_b, g)[i];
                    vec_s[V_FC + i] = i < N ? p.fc[i] : 1.0f;
                    vec_s[V_Q0 + i] = i < N ? p.q0[i] : 1.0f;
                }
                if (tid < kU) {
                    const float* b_ih = ctrl_ptr(p.b_ih, g);
                    const float* b_hh = ctrl_ptr(p.b_hh, g);
                    const int o = rank * kU + tid;
                    vec_s[V_BR + tid] = b_ih[o] + b_hh[o];
                    vec_s[V_BZ + tid] = b_ih[kHid + o] + b_hh[kHid + o];
                    vec_s[V_BIN + tid] = b_ih[2 * kHid + o];
                    vec_s[V_BHN + tid] = b_hh[2 * kHid + o];
                    vec_s[V_B1 + tid] = ctrl_ptr(p.b1, g)[o];
                    vec_s[V_B2 + tid] = ctrl_ptr(p.b2, g)[o];
                    const int n = rank * NU + tid;
                    const bool own = tid < NU && n < N;
                    vec_s[V_B3 + tid] = own ? ctrl_ptr(p.b3, g)[n] : 0.f;
                    vec_s[V_Q0S + tid] = own ? p.q0[n] : 1.f;
                    vec_s[V_DQS + tid] = own ? p.dq[n] : 0.f;
                }
                cur_g = g;
                __syncthreads();
            }
            // ---- state of this (tile, frame) ---------------------------------------------------------------
            bool fallback_prev = false;                       // Q_t was replaced by Q0 and h_{t-1} dropped
            if (t > 0) fallback_prev = __ldcg(p.flags + (t - 1) * p.G + g) != 0;
            const bool use_q0 = (t == 0) || fallback_prev;
            const bool h_zero = use_q0;
            float* hcur_s = hbuf_s + hsel * kHid * kR;
            float* hnext_s = hbuf_s + (hsel ^ 1) * kHid * kR;
            if (use_q0) {
                for (int idx = tid; idx < kRT * N; idx += kSeqThreads) {
                    const int i = idx / N, n = idx - i * N;
                    q_s[n * kRT + i] = vec_s[V_Q0 + n];
                    if (bb0 + i < p.B) p.Q[((grow0 + i) * T + t) * N + n] = vec_s[V_Q0 + n];
                }
            } else {
                for (int idx = tid; idx < kRT * N; idx += kSeqThreads) {
                    const int i = idx / N, n = idx - i * N;
                    q_s[n * kRT + i] = (bb0 + i < p.B) ? __ldcg(p.Q + ((grow0 + i) * T + t) * N + n) : vec_s[V_Q0 + n];
                }
                const float4* hsrc = reinterpret_cast<const float4*>(h_tile(p, g, t - 1, tiles, tile));
                for (int i = tid; i < kHid * kR / 4; i += kSeqThreads) reinterpret_cast<float4*>(hcur_s)[i] = __ldcg(hsrc + i);
            }
            if (fallback_prev && rank == 0) {   // h_{t-1} was dropped: it must not feed dW_hh either
                float4* hdst = reinterpret_cast<float4*>(h_tile(p, g, t - 1, tiles, tile));
                for (int i = tid; i < kHid * kR / 4; i += kSeqThreads) hdst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // spectra of this CTA's rows for frame t (the STFT has long finished when this pass runs; the wait is a no-op
            // then, but the hand-over protocol allows a caller to stream them here too)
            wait_spectra(p, grow0, bb0, t);
            prefetch_spectra(p, spec_s, L.tile, grow0, bb0, t);
            finish_spectra(p, spec_s, L.tile, bb0);
            __syncthreads();

            // ---- band stage of frame t for this CTA's 4 rows (model_torch.py:340-346, 1050-1060) ---------------
            // The 4 x quads (row, 4-band quad) pairs are dealt round-robin to the 16 warps, widest quads first and
            // rows rotating, so every warp gets the same mix.  Lane l of a warp OWNS band (l & 3) of the warp's
            // (l >> 2)-th pair: it computes that band's window parameters once, lends them to the 8 lanes that walk
            // the window, gets the 8 reduced sums back and does the (expensive: 3 divisions, atan2, log1p) epilogue
            // for it -- once per warp with up to 32 bands in flight instead of once per quad with 4.
            bool own_store = false;
            long long own_e = 0;
            float oY = 0.f, oJ = 0.f, oP = 0.f, oK = 0.f;
            {
                const int warp = tid >> 5;
                constexpr int kWarps = kSeqThreads / 32;
                const int n_pairs = kRT * quads;
                const int p_own = warp + kWarps * (lane >> 2);
                const int row_own = ((p_own & 3) + (p_own >> 4)) & 3;
                const int n_own = ((quads - 1 - (p_own >> 2)) << 2) + (lane & 3);
                const bool own = p_own < n_pairs && n_own < N;
                const float fc = own ? vec_s[V_FC + n_own] : 1.0f;
                const float q = own ? q_s[n_own * kRT + row_own] : 1.0f;
                const BandParams bp_own = band_params(fc, q, p.df, p.cutoff, p.F, own);
                BandSums keep = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int m = 0, pr = warp; pr < n_pairs; ++m, pr += kWarps) {
                    const int row = ((pr & 3) + (pr >> 4)) & 3;
                    const int src = (m << 2) + (lane >> 3);          // lane owning the band this lane helps with
                    BandParams bp;
                    bp.bw = 0.f;
                    bp.a = __shfl_sync(0xffffffffu, bp_own.a, src);
                    bp.b = __shfl_sync(0xffffffffu, bp_own.b, src);
                    bp.kc = __shfl_sync(0xffffffffu, bp_own.kc, src);
                    bp.k_lo = __shfl_sync(0xffffffffu, bp_own.k_lo, src);
                    bp.k_hi = __shfl_sync(0xffffffffu, bp_own.k_hi, src);
                    const BandSums sums = band_accumulate(spec_s + row * L.tile, p.F, bp, lane);
                    const int from = (lane & 3) << 3;                // any lane of the group that holds my band's sums
                    const bool mine = (lane >> 2) == m;
                    float v;
                    v = __shfl_sync(0xffffffffu, sums.S, from);   if (mine) keep.S = v;
                    v = __shfl_sync(0xffffffffu, sums.Y, from);   if (mine) keep.Y = v;
                    v = __shfl_sync(0xffffffffu, sums.Zr, from);  if (mine) keep.Zr = v;
                    v = __shfl_sync(0xffffffffu, sums.Zi, from);  if (mine) keep.Zi = v;
                    v = __shfl_sync(0xffffffffu, sums.m2, from);  if (mine) keep.m2 = v;
                    v = __shfl_sync(0xffffffffu, sums.a2, from);  if (mine) keep.a2 = v;
                    v = __shfl_sync(0xffffffffu, sums.z2r, from); if (mine) keep.z2r = v;
                    v = __shfl_sync(0xffffffffu, sums.z2i, from); if (mine) keep.z2i = v;
                }
                if (own) {
                    const BandResult r = band_finish(keep);
                    ystage_s[row_own * kHid + n_own] = log1pf(fmaxf(r.Y, 0.0f));
                    own_store = bb0 + row_own < p.B;
                    own_e = ((grow0 + row_own) * T + t) * N + n_own;
                    const float qe = q + 1e-8f;
                    const float kappa = -fc / (qe * qe * bp_own.bw);
                    oY = r.Y;
                    oJ = kappa * (r.a2 - r.Yraw * r.m2);
                    if (p.phase) {
                        oP = atan2f(r.Zi, r.Zr);
                        const float mag2 = r.Zr * r.Zr + r.Zi * r.Zi;
                        oK = mag2 > 0.0f ? kappa * (r.Zr * r.z2i - r.Zi * r.z2r) / mag2 : 0.0f;
                    }
                }
            }
            // Results of this lane's band.  They go to HBM AFTER the cluster barrier below has been signalled: the
            // barrier's release waits for every earlier global store of the thread, and nobody in the cluster needs these.
            auto store_band_outputs = [&]() {
                if (!own_store) return;
                p.Y[own_e] = oY;
                p.dYdQ[own_e] = oJ;
                if (p.logY) p.logY[own_e] = fminf(fmaxf(logf(oY + 1e-8f), -12.0f), 12.0f);
                if (p.phase) {
                    p.phase[own_e] = oP;
                    p.dPdQ[own_e] = oK;
                }
            };
            if (t == T - 1) {
                // The reference runs the controller once more and discards the result (model_torch.py:361-380).
                store_band_outputs();
                __syncthreads();
                continue;
            }
            __syncthreads();
            const long long tb = tile_base(p, g, t, tiles, tile);
            if (tid < N) {   // features of my 4 rows -> every CTA of the cluster (+ saved for dW_ih)
                const float4 v = make_float4(ystage_s[tid], ystage_s[kHid + tid], ystage_s[2 * kHid + tid],
                                             ystage_s[3 * kHid + tid]);
#pragma unroll
                for (int dst = 0; dst < kCS; ++dst)
                    *reinterpret_cast<float4*>(cluster.map_shared_rank(yc_s, dst) + tid * kR + rank * kRT) = v;
            }
            cluster.barrier_arrive();   // #1: my part of yc is delivered ...
            if (tid < N)
                *reinterpret_cast<float4*>(p.yc + tb * N * kR + tid * kR + rank * kRT) =
                    make_float4(ystage_s[tid], ystage_s[kHid + tid], ystage_s[2 * kHid + tid], ystage_s[3 * kHid + tid]);
            store_band_outputs();
            cluster.barrier_wait();     // ... and complete everywhere

            // ---- GRU cell (torch gate order r, z, n; n = tanh(i_n + r * (W_hn h + b_hn))) ----------------------
            {
                float ar[kRT] = {0.f, 0.f, 0.f, 0.f}, az[kRT] = {0.f, 0.f, 0.f, 0.f};
                float ain[kRT] = {0.f, 0.f, 0.f, 0.f}, ahn[kRT] = {0.f, 0.f, 0.f, 0.f};
                int k0, k1;
                k_range(N, ks, k0, k1);
                dot_rows3(ar, az, ain, yc_s + rg * kRT, img_s + fwd_img_wih(N) + u, k0, k1);
                k_range(kHid, ks, k0, k1);
                if (!h_zero) dot_rows3(ar, az, ahn, hcur_s + rg * kRT, img_s + fwd_img_whh(N) + u, k0, k1);
                float acc[16];
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    acc[i] = ar[i];
                    acc[4 + i] = az[i];
                    acc[8 + i] = ain[i];
                    acc[12 + i] = ahn[i];
                }
                reduce_ks<16>(acc, red_gru_s, ks, slot);
                float hv[kRT], vr[kRT], vz[kRT], vn[kRT], vh[kRT];
                if (ks == 0) {
                    const float br = vec_s[V_BR + u], bz = vec_s[V_BZ + u], bin = vec_s[V_BIN + u], bhn = vec_s[V_BHN + u];
#pragma unroll
                    for (int i = 0; i < kRT; ++i) {
                        vr[i] = sigmoid_fast(acc[i] + br);
                        vz[i] = sigmoid_fast(acc[4 + i] + bz);
                        vh[i] = acc[12 + i] + bhn;
                        vn[i] = tanhf(acc[8 + i] + bin + vr[i] * vh[i]);
                        const float hp = h_zero ? 0.0f : hcur_s[ug * kR + rg * kRT + i];
                        hv[i] = (1.0f - vz[i]) * vn[i] + vz[i] * hp;
                    }
                    broadcast_rows(cluster, hnext_s, ug, rg * kRT, hv);
                }
                cluster.barrier_arrive();   // #2 signalled before the saves go out (a later release covers them)
                if (ks == 0) {
                    store4(h_tile(p, g, t, tiles, tile) + ug * kR + rg * kRT, hv);
                    float* gt = p.gates + tb * 4 * kHid * kR + ug * kR + rg * kRT;
                    store4(gt, vr);
                    store4(gt + kHid * kR, vz);
                    store4(gt + 2 * kHid * kR, vn);
                    store4(gt + 3 * kHid * kR, vh);
                }
            }
            cluster.barrier_wait();   // #2: h_t complete everywhere

            // ---- Linear 1 -> LayerNorm -> SiLU -> Dropout ------------------------------------------------------
            {
                float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
                int k0, k1;
                k_range(kHid, ks, k0, k1);
                dot_rows(acc, hnext_s + rg * kRT, img_s + fwd_img_w1(N) + u, k0, k1);
                reduce_ks1<kRT>(acc, red_l1_s, ks, slot);
                if (ks == 0) {
                    const float bb = vec_s[V_B1 + u];
#pragma unroll
                    for (int i = 0; i < kRT; ++i) acc[i] += bb;
                    broadcast_rows(cluster, a1_s, ug, rg * kRT, acc);
                }
            }
            cluster.sync();   // #3
            ln_silu_drop_fwd(p, seed, a1_s, stat_s, vec_s + V_LN1G, vec_s + V_LN1B, 0, t, (long long)g * p.B + b0, rank,
                             p.xh1 + tb * kHid * kR, p.d1 + tb * kHid * kR, p.rstd + tb * 2 * kR);
            // ---- Linear 2 -> LayerNorm -> SiLU -> Dropout ------------------------------------------------------
            {
                float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
                int k0, k1;
                k_range(kHid, ks, k0, k1);
                dot_rows(acc, a1_s + rg * kRT, img_s + fwd_img_w2(N) + u, k0, k1);
                reduce_ks1<kRT>(acc, red_l23_s, ks, slot);
                if (ks == 0) {
                    const float bb = vec_s[V_B2 + u];
#pragma unroll
                    for (int i = 0; i < kRT; ++i) acc[i] += bb;
                    // a2 aliases yc: every CTA's yc reads (its GRU products) ended before it sent h_t, i.e. before #2
                    broadcast_rows(cluster, a2_s, ug, rg * kRT, acc);   // a2 aliases yc: all yc reads ended before #2
                }
            }
            cluster.sync();   // #4
            ln_silu_drop_fwd(p, seed, a2_s, stat_s, vec_s + V_LN2G, vec_s + V_LN2B, 1, t, (long long)g * p.B + b0, rank,
                             p.xh2 + tb * kHid * kR, p.d2 + tb * kHid * kR, p.rstd + tb * 2 * kR);
            // ---- Linear 3 -> tanh -> Q_{t+1} (model_torch.py:367-380) -------------------------------------------
            {
                float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
                const bool mine = u < nu_c;
                int k0, k1;
                k_range(kHid, ks, k0, k1);
                if (mine) dot_rows(acc, a2_s + rg * kRT, img_s + fwd_img_w3(N) + u, k0, k1);
                reduce_ks1<kRT>(acc, red_l23_s, ks, slot);
                const bool fin = ks == 0 && mine;
                const int n = rank * NU + u;
                float qv[kRT] = {0.f, 0.f, 0.f, 0.f}, dv[kRT] = {0.f, 0.f, 0.f, 0.f};
                if (fin) {
                    const float bb = vec_s[V_B3 + u], q0 = vec_s[V_Q0S + u], dq = vec_s[V_DQS + u];
                    bool bad = false;
#pragma unroll
                    for (int i = 0; i < kRT; ++i) {
                        dv[i] = tanhf(acc[i] + bb);
                        const float qu = p.relative ? q0 * (1.0f + dq * dv[i]) : fmaf(dq, dv[i], q0);
                        qv[i] = fminf(fmaxf(qu, p.q_min), p.q_max);
                        bad = bad || (b0 + rg * kRT + i < p.B && !finite_f(qu));
                    }
                    // Q_{t+1} of rows 4rg..4rg+3 goes to the CTA that runs their band stage
                    store4(cluster.map_shared_rank(q_s, rg) + n * kRT, qv);
                    if (bad) {   // NaN / Inf: the reference falls back for the whole batch of this ear
                        atomicOr(p.flags + t * p.G + g, 1);
                        atomicOr(any_flag, 1);
                    }
                }
                auto store_q = [&]() {
                    if (!fin) return;
#pragma unroll
                    for (int i = 0; i < kRT; ++i) {
                        const int b = b0 + rg * kRT + i;
                        if (b < p.B) {
                            const long long e = (((long long)g * p.B + b) * T + (t + 1)) * N + n;
                            p.Q[e] = qv[i];
                            p.delta[e] = dv[i];
                        }
                    }
                };
                // Q / flags / H are read back from global memory by other CTAs in the next frame: they precede the barrier
                store_q();
                __threadfence();
                cluster.barrier_arrive();   // #5
            }
            cluster.barrier_wait();   // #5: Q_{t+1} delivered; a2 / yc free for the next frame
        }
    }
}

// ==================================================================================================
// forward, fast pass: TWO independent chains per CTA
// ==================================================================================================
// The tile of 16 rows is split into two half-tiles of 8 rows; warps 0-7 of every CTA of the cluster carry half-tile 0
// through the recurrence, warps 8-15 half-tile 1, each with its own activation buffers, named block barrier and hand-over
// mbarriers, sharing only the (read-only) weight image.  The two chains drift apart on their own, so the FMA-bound band
// stage of one half overlaps the latency-bound controller phases of the other: every chain keeps the serial latency of
// its phases, but the throughput-bound parts of a phase (the band loop, the shared-memory traffic of the dot products)
// now serve 8 rows instead of 16.  Same arithmetic per row as the single-chain kernel (which remains as the strict pass).
constexpr int kCh = 2;                         // chains per CTA
constexpr int kCT = kSeqThreads / kCh;         // 256 threads per chain
constexpr int kRC = kR / kCh;                  // 8 rows of the tile per chain
constexpr int kRB = kRC / kCS;                 // 2 band-stage rows per CTA and chain
constexpr int kRGC = kRC / kRT;                // 2 row groups of 4 rows in the GEMM phases
constexpr int kSlotsC = kRGC * kU;             // 64 (row group, unit) slots per k-split
constexpr int kWarpsC = kCT / 32;              // 8 warps per chain
static_assert(kKS * kSlotsC == kCT && kRB * kCS == kRC, "chain thread layout");

struct Fwd2Smem {   // offsets in floats
    int N, tile;
    __host__ __device__ Fwd2Smem(int N_, int F) : N(N_), tile(spec_tile_len(F)) {}
    __host__ __device__ int vec() const { return fwd_img_floats(N); }           // behind the (shared) weight image
    __host__ __device__ int chain0() const { return vec() + 1088; }
    // per chain:
    __host__ __device__ int c_yc() const { return 0; }                           // [128][kRC]; aliased by a2
    __host__ __device__ int c_a1() const { return kHid * kRC; }                  // directly behind yc
    __host__ __device__ int c_h() const { return 2 * kHid * kRC; }               // x 2 (ping-pong)
    __host__ __device__ int c_stat() const { return 4 * kHid * kRC; }            // 2 * kCT
    __host__ __device__ int c_q() const { return c_stat() + 2 * kCT; }           // [128][kRB]
    __host__ __device__ int c_ystage() const { return c_q() + kHid * kRB; }      // [kRB][128]
    __host__ __device__ int c_spec() const { return c_ystage() + kRB * kHid; }   // [kRB][tile] float4
    __host__ __device__ int c_bars() const { return c_spec() + kRB * tile * 4; } // 5 mbarriers
    __host__ __device__ int c_total() const { return c_bars() + 16; }
    __host__ __device__ int total() const { return chain0() + kCh * c_total(); }
};

__device__ __forceinline__ void chain_sync(int chain) {      // block barrier of one chain (named barrier 1 + chain)
    asm volatile("bar.sync %0, %1;" ::"r"(chain + 1), "n"(kCT) : "memory");
}
// Band-stage token between the two chains (named barriers 3 and 4, one chain arrives, the other waits): the band stages of
// the two half-tiles strictly alternate -- chain 0, chain 1, chain 0, ... -- so that each one has the FMA pipe to itself
// while the other chain is in its (latency-bound) controller phases.  Left alone the two chains stay in lockstep: they do
// the same work, and a phase offset between them neither grows nor shrinks.
__device__ __forceinline__ void token_pass(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kSeqThreads) : "memory"); }
__device__ __forceinline__ void token_take(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kSeqThreads) : "memory"); }
constexpr int kTokenToChain0 = 3, kTokenToChain1 = 4;

// k-split sums of a chain (see reduce_ks / reduce_ks1): 64 slots, the chain's named barrier
template <int NACC>
__device__ __forceinline__ void reduce_ks_c(float* acc, float* red_s, int ks, int slot, int chain) {
    chain_sync(chain);
    if (ks >= 2) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) red_s[((ks - 2) * NACC + i) * kSlotsC + slot] = acc[i];
    }
    chain_sync(chain);
    if (ks < 2) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] += red_s[(ks * NACC + i) * kSlotsC + slot];
    }
    chain_sync(chain);
    if (ks == 1) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) red_s[i * kSlotsC + slot] = acc[i];
    }
    chain_sync(chain);
    if (ks == 0) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] += red_s[i * kSlotsC + slot];
    }
}

template <int NACC>
__device__ __forceinline__ void reduce_ks1_c(float* acc, float* red_s, int ks, int slot, int chain) {
    chain_sync(chain);
    if (ks > 0) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) red_s[((ks - 1) * NACC + i) * kSlotsC + slot] = acc[i];
    }
    chain_sync(chain);
    if (ks == 0) {
#pragma unroll
        for (int k = 0; k < kKS - 1; ++k)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] += red_s[(k * NACC + i) * kSlotsC + slot];
    }
}

// LayerNorm + SiLU + Dropout over the full rows of one chain in buf_s ([feature][kRC], in place), redundantly in every CTA;
// the threads whose features are the CTA's own slice save them (tile layout, row offset row_off inside the 16-row tile).
// `sc`: the keep-mask scales of this thread's 4 features (drawn by the caller while it waits for hand-over #1).
__device__ __forceinline__ void ln_silu_drop_fwd_c(const float4 sc, float* buf_s, float* stat_s,
                                                   const float* __restrict__ gamma_s, const float* __restrict__ beta_s,
                                                   int layer, int rank, int chain, int tid,
                                                   float* xh_tile, float* d_tile, float* rstd_tile, int row_off) {
    constexpr int PARTS = kCT / kRC, FPP = kHid / PARTS;             // 32 parts of 4 features
    static_assert(kRC == 8 && FPP == 4, "lanes l, l^8, l^16, l^24 of a warp hold four parts of one row");
    const int row = tid % kRC, part = tid / kRC;
    const int lane = tid & 31, warp = tid >> 5;
    const int f0 = part * FPP;
    const float pivot = buf_s[row];
    float v[FPP];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        v[i] = buf_s[(f0 + i) * kRC + row] - pivot;
        s += v[i];
        ss = fmaf(v[i], v[i], ss);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    ss += __shfl_xor_sync(0xffffffffu, ss, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    ss += __shfl_xor_sync(0xffffffffu, ss, 16);
    float2* stat2 = reinterpret_cast<float2*>(stat_s);               // [row][kWarpsC + 1] float2
    if (lane < kRC) stat2[row * (kWarpsC + 1) + warp] = make_float2(s, ss);
    chain_sync(chain);
    float S = 0.f, SS = 0.f;
#pragma unroll
    for (int w = 0; w < kWarpsC; ++w) {
        const float2 q = stat2[row * (kWarpsC + 1) + w];
        S += q.x;
        SS += q.y;
    }
    const float mean = S * (1.0f / kHid);
    const float var = fmaxf(fmaf(-mean, mean, SS * (1.0f / kHid)), 0.0f) * kHid;
#pragma unroll
    for (int i = 0; i < FPP; ++i) v[i] -= mean;
    const float rstd = rsqrtf(var * (1.0f / kHid) + kLnEps);
    const bool mine = f0 / kU == rank;
    if (rank == 0 && part == 0) rstd_tile[layer * kR + row_off + row] = rstd;
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        const int f = f0 + i;
        const float xh = v[i] * rstd;
        const float y = fmaf(xh, gamma_s[f], beta_s[f]);
        const float o = (y * sigmoid_fast(y)) * scv[i];
        buf_s[f * kRC + row] = o;
        if (mine) {
            xh_tile[f * kR + row_off + row] = xh;
            d_tile[f * kR + row_off + row] = o;
        }
    }
    chain_sync(chain);
}

__global__ void __launch_bounds__(kSeqThreads, 1) seq_fwd2_kernel(const BiearSeqParams p, const float* __restrict__ img) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int N = p.N, T = p.T, S = p.T - 1;
    const int tiles = (p.B + kR - 1) / kR;
    const int NU = bands_per_cta(N);
    const Fwd2Smem L(N, p.F);
    float* img_s = smem;
    float* vec_s = smem + L.vec();
    const int chain = (int)threadIdx.x / kCT, tid = (int)threadIdx.x % kCT, lane = tid & 31, warp = tid >> 5;
    float* cs = smem + L.chain0() + chain * L.c_total();
    float* yc_s = cs + L.c_yc();
    float* a2_s = yc_s;
    float* a1_s = cs + L.c_a1();
    float* hbuf_s = cs + L.c_h();
    float* stat_s = cs + L.c_stat();
    float* q_s = cs + L.c_q();
    float* ystage_s = cs + L.c_ystage();
    float4* spec_s = reinterpret_cast<float4*>(cs + L.c_spec());
    // reduction scratch in activation buffers that are idle at that point of the frame (as in seq_fwd_strict_kernel)
    float* red_gru_s = yc_s;      // yc | a1: 2 x 16 x 64 floats
    float* red_l1_s = yc_s;
    float* red_l23_s = a1_s;
    static_assert(2 * 16 * kSlotsC <= 2 * kHid * kRC && (kKS - 1) * kRT * kSlotsC <= kHid * kRC, "reduction scratch fits");
    const int ks = tid >> 6, slot = tid & (kSlotsC - 1), rg = slot / kU, u = slot % kU;
    const int ug = rank * kU + u;
    const int nu_c = max(0, min(NU, N - rank * NU));
    const int quads = (N + 3) >> 2;
    int* any_flag = p.flags + S * p.G;
    const unsigned long long seed = p.seed_ptr ? *p.seed_ptr : p.seed;
    const int tl = (int)(blockIdx.x / kCS);
    const int g = tl / tiles, tile = tl % tiles;
    const int row_off = chain * kRC;                               // this chain's rows inside the 16-row tile
    const int b0 = tile * kR + row_off;                            // clip index of the chain's first row
    const long long crow0 = (long long)g * p.B + b0;               // global row of the chain's first row
    const long long grow0 = crow0 + rank * kRB;                    // ... of this CTA's first band-stage row
    const int bb0 = b0 + rank * kRB;

    const uint32_t bars = smem_u32(cs + L.c_bars());
    auto bar_of = [&](int i) { return bars + 8u * (uint32_t)i; };  // i = 0..4: yc, h, a1, a2, Q
    const uint32_t tx_bytes[5] = {(uint32_t)(N * kRC * 4), (uint32_t)(kHid * kRC * 4), (uint32_t)(kHid * kRC * 4),
                                  (uint32_t)(kHid * kRC * 4), (uint32_t)(N * kRB * 4)};
    auto arm = [&](int i) {
        if (tid == 0) mbar_arrive_expect_tx(bar_of(i), tx_bytes[i]);
    };

    // ---- set-up by the whole CTA: weight image, constants, barriers -----------------------------------------------
    copy_f4(reinterpret_cast<float4*>(img_s),
            reinterpret_cast<const float4*>(img + (long long)(g * kCS + rank) * fwd_img_floats(N)), fwd_img_floats(N) / 4);
    for (int i = threadIdx.x; i < kHid; i += kSeqThreads) {
        vec_s[V_LN1G + i] = ctrl_ptr(p.ln1_g, g)[i];
        vec_s[V_LN1B + i] = ctrl_ptr(p.ln1_b, g)[i];
        vec_s[V_LN2G + i] = ctrl_ptr(p.ln2_g, g)[i];
        vec_s[V_LN2B + i] = ctrl_ptr(p.ln2_b, g)[i];
        vec_s[V_FC + i] = i < N ? p.fc[i] : 1.0f;
        vec_s[V_Q0 + i] = i < N ? p.q0[i] : 1.0f;
    }
    if (threadIdx.x < kU) {
        const int i = threadIdx.x;
        const float* b_ih = ctrl_ptr(p.b_ih, g);
        const float* b_hh = ctrl_ptr(p.b_hh, g);
        const int o = rank * kU + i;
        vec_s[V_BR + i] = b_ih[o] + b_hh[o];
        vec_s[V_BZ + i] = b_ih[kHid + o] + b_hh[kHid + o];
        vec_s[V_BIN + i] = b_ih[2 * kHid + o];
        vec_s[V_BHN + i] = b_hh[2 * kHid + o];
        vec_s[V_B1 + i] = ctrl_ptr(p.b1, g)[o];
        vec_s[V_B2 + i] = ctrl_ptr(p.b2, g)[o];
        const int n = rank * NU + i;
        const bool own = i < NU && n < N;
        vec_s[V_B3 + i] = own ? ctrl_ptr(p.b3, g)[n] : 0.f;
        vec_s[V_Q0S + i] = own ? p.q0[n] : 1.f;
        vec_s[V_DQS + i] = own ? p.dq[n] : 0.f;
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) mbar_init(bar_of(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();                                           // every CTA's barriers exist before anybody signals them
    PHASE_INIT();
    int spec_t = -1, hsel = 0;

    // spectra of this chain's kRB rows for frame t -> {1, abs, re, im} tiles (cp.async now, conversion later)
    auto prefetch = [&](int t) {
#pragma unroll
        for (int i = 0; i < kRB; ++i) {
            const float2* src = reinterpret_cast<const float2*>(p.X) + ((grow0 + i) * p.T + t) * p.F;
            const bool row_ok = bb0 + i < p.B;
            for (int k = tid; k < L.tile; k += kCT) {
                float4* slot4 = spec_s + i * L.tile + k;
                if (row_ok && k < p.F) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(&slot4->z);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + k) : "memory");
                } else {
                    *slot4 = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto finish = [&]() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int i = 0; i < kRB; ++i) {
            if (bb0 + i >= p.B) continue;
            for (int k = tid; k < p.F; k += kCT) {
                float4* slot4 = spec_s + i * L.tile + k;
                *slot4 = spec_entry(make_float2(slot4->z, slot4->w));
            }
        }
    };
    // streamed hand-over of the spectra (x_ready): flag of (row, frame), polled by one thread per band-stage row
    auto flag_of = [&](int t) { return p.x_ready + (grow0 + tid) * p.T + t; };
    auto poll = [&](int t, int v) {
        for (unsigned spin = 0; v == 0; ++spin) {
            asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(flag_of(t)) : "memory");
            if (v != 0) break;
            if (spin > (1u << 24)) __trap();
            __nanosleep(64);
        }
    };

    for (int t = 0; t < T; ++t) {
        float* hcur_s = hbuf_s + hsel * kHid * kRC;
        float* hnext_s = hbuf_s + (hsel ^ 1) * kHid * kRC;
        const bool h_zero = t == 0;
        if (t == 0) {
            for (int idx = tid; idx < kRB * N; idx += kCT) {
                const int i = idx / N, n = idx - i * N;
                q_s[n * kRB + i] = vec_s[V_Q0 + n];
                if (bb0 + i < p.B) p.Q[((grow0 + i) * T + t) * N + n] = vec_s[V_Q0 + n];
            }
        }
        PHASE_MARK(0, 0);    // loop head
        if (spec_t != t) {   // first frame: fetch and convert now (also orders the q_s fill)
            if (p.x_ready && tid < kRB && bb0 + tid < p.B) poll(t, 0);
            chain_sync(chain);
            prefetch(t);
            finish();
            chain_sync(chain);
        }
        PHASE_MARK(0, 1);    // spectra ready
        int next_ready = 1;
        if (p.x_ready && t + 1 < T && tid < kRB && bb0 + tid < p.B)
            asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(next_ready) : "l"(flag_of(t + 1)) : "memory");

        // ---- band stage of frame t for this chain's kRB rows of this CTA ------------------------------------------
        // The kRB x quads (row, quad) pairs are dealt to the chain's 8 warps widest quads first, in SNAKE order (round m goes
        // warp 0..7 for even m, 7..0 for odd m): with the window widths growing monotonically with the band index a plain
        // round-robin gives the first warps 11 % more bins than the average, the snake 4 %.  Lane l OWNS band (l & 3) of
        // the warp's (l >> 2)-th pair (parameters once, epilogue once per warp).
        // (token: chain 0 goes first in every frame, chain 1 follows it, chain 0's next frame follows chain 1)
        if (chain == 0) {
            if (t > 0) token_take(kTokenToChain0);
        } else {
            token_take(kTokenToChain1);
        }
        bool own_store = false;
        long long own_e = 0;
        float oY = 0.f, oJ = 0.f, oP = 0.f, oK = 0.f;
        {
            const int n_pairs = kRB * quads;
            const int m_own = lane >> 2;
            const int p_own = kWarpsC * m_own + ((m_own & 1) ? kWarpsC - 1 - warp : warp);
            const int row_own = p_own & (kRB - 1);
            const int n_own = ((quads - 1 - (p_own / kRB)) << 2) + (lane & 3);
            const bool own = p_own < n_pairs && n_own < N;
            const float fc = own ? vec_s[V_FC + n_own] : 1.0f;
            const float q = own ? q_s[n_own * kRB + row_own] : 1.0f;
            const BandParams bp_own = band_params(fc, q, p.df, p.cutoff, p.F, own);
            BandSums keep = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int m = 0; m * kWarpsC < n_pairs; ++m) {
                const int pr = kWarpsC * m + ((m & 1) ? kWarpsC - 1 - warp : warp);
                if (pr >= n_pairs) break;                          // (warp-uniform)
                const int row = pr & (kRB - 1);
                const int src = (m << 2) + (lane >> 3);          // lane owning the band this lane helps with
                BandParams bp;
                bp.bw = 0.f;
                bp.a = __shfl_sync(0xffffffffu, bp_own.a, src);
                bp.b = __shfl_sync(0xffffffffu, bp_own.b, src);
                bp.kc = __shfl_sync(0xffffffffu, bp_own.kc, src);
                bp.k_lo = __shfl_sync(0xffffffffu, bp_own.k_lo, src);
                bp.k_hi = __shfl_sync(0xffffffffu, bp_own.k_hi, src);
                const BandSums sums = band_accumulate(spec_s + row * L.tile, p.F, bp, lane);
                const int from = (lane & 3) << 3;                // any lane of the group that holds my band's sums
                const bool mine = (lane >> 2) == m;
                float v;
                v = __shfl_sync(0xffffffffu, sums.S, from);   if (mine) keep.S = v;
                v = __shfl_sync(0xffffffffu, sums.Y, from);   if (mine) keep.Y = v;
                v = __shfl_sync(0xffffffffu, sums.Zr, from);  if (mine) keep.Zr = v;
                v = __shfl_sync(0xffffffffu, sums.Zi, from);  if (mine) keep.Zi = v;
                v = __shfl_sync(0xffffffffu, sums.m2, from);  if (mine) keep.m2 = v;
                v = __shfl_sync(0xffffffffu, sums.a2, from);  if (mine) keep.a2 = v;
                v = __shfl_sync(0xffffffffu, sums.z2r, from); if (mine) keep.z2r = v;
                v = __shfl_sync(0xffffffffu, sums.z2i, from); if (mine) keep.z2i = v;
            }
            if (own) {
                const BandResult r = band_finish(keep);
                ystage_s[row_own * kHid + n_own] = log1pf(fmaxf(r.Y, 0.0f));
                own_store = bb0 + row_own < p.B;
                own_e = ((grow0 + row_own) * T + t) * N + n_own;
                const float qe = q + 1e-8f;
                const float kappa = -fc / (qe * qe * bp_own.bw);
                oY = r.Y;
                oJ = kappa * (r.a2 - r.Yraw * r.m2);
                if (p.phase) {
                    oP = atan2f(r.Zi, r.Zr);
                    const float mag2 = r.Zr * r.Zr + r.Zi * r.Zi;
                    oK = mag2 > 0.0f ? kappa * (r.Zr * r.z2i - r.Zi * r.z2r) / mag2 : 0.0f;
                }
            }
        }
        auto store_band_outputs = [&]() {
            if (!own_store) return;
            p.Y[own_e] = oY;
            p.dYdQ[own_e] = oJ;
            if (p.logY) p.logY[own_e] = fminf(fmaxf(logf(oY + 1e-8f), -12.0f), 12.0f);
            if (p.phase) {
                p.phase[own_e] = oP;
                p.dPdQ[own_e] = oK;
            }
        };
        PHASE_MARK(0, 2);    // band stage (this warp)
        if (chain == 0) token_pass(kTokenToChain1);
        else if (t + 1 < T) token_pass(kTokenToChain0);
        if (t == T - 1) {
            // The reference runs the controller once more and discards the result (model_torch.py:361-380).
            store_band_outputs();
            continue;
        }
        if (p.x_ready && tid < kRB && bb0 + tid < p.B) poll(t + 1, next_ready);   // (ordered for the chain by the barrier below)
        chain_sync(chain);
        prefetch(t + 1);     // the tile is free again: the next frame's spectra travel behind the controller phases
        spec_t = t + 1;
        const long long tb = tile_base(p, g, t, tiles, tile);
        const uint32_t par = (uint32_t)t & 1u;
        if (tid < N) {   // features of my kRB rows -> every CTA of the cluster (+ saved for dW_ih)
            const float2 v = make_float2(ystage_s[tid], ystage_s[kHid + tid]);
            const uint32_t a = smem_u32(yc_s + tid * kRC + rank * kRB);
#pragma unroll
            for (uint32_t dst = 0; dst < (uint32_t)kCS; ++dst)
                st_async_f2(cluster_addr(a, dst), v, cluster_addr(bar_of(0), dst));
            *reinterpret_cast<float2*>(p.yc + tb * N * kR + tid * kR + row_off + rank * kRB) = v;
        }
        arm(0);
        store_band_outputs();
        // the two dropout masks of this frame (Philox, ~120 instructions) are drawn here, while hand-over #1 is in flight,
        // instead of inside the LayerNorm phases of the serial chain
        float4 drop1 = make_float4(1.f, 1.f, 1.f, 1.f), drop2 = drop1;
        if (p.training) {
            drop1 = dropout_scale4(seed, t, 0, crow0 + (tid % kRC), tid / kRC);
            drop2 = dropout_scale4(seed, t, 1, crow0 + (tid % kRC), tid / kRC);
        }
        tx_wait(bar_of(0), par);
        PHASE_MARK(0, 3);    // band barrier + push + hand-over #1

        // ---- GRU cell --------------------------------------------------------------------------------------------
        {
            float ar[kRT] = {0.f, 0.f, 0.f, 0.f}, az[kRT] = {0.f, 0.f, 0.f, 0.f};
            float ain[kRT] = {0.f, 0.f, 0.f, 0.f}, ahn[kRT] = {0.f, 0.f, 0.f, 0.f};
            int k0, k1;
            k_range(N, ks, k0, k1);
            dot_rows3<kRC>(ar, az, ain, yc_s + rg * kRT, img_s + fwd_img_wih(N) + u, k0, k1);
            k_range(kHid, ks, k0, k1);
            if (!h_zero) dot_rows3<kRC>(ar, az, ahn, hcur_s + rg * kRT, img_s + fwd_img_whh(N) + u, k0, k1);
            float acc[16];
#pragma unroll
            for (int i = 0; i < kRT; ++i) {
                acc[i] = ar[i];
                acc[4 + i] = az[i];
                acc[8 + i] = ain[i];
                acc[12 + i] = ahn[i];
            }
            reduce_ks_c<16>(acc, red_gru_s, ks, slot, chain);
            if (ks == 0) {
                float hv[kRT], vr[kRT], vz[kRT], vn[kRT], vh[kRT];
                const float br = vec_s[V_BR + u], bz = vec_s[V_BZ + u], bin = vec_s[V_BIN + u], bhn = vec_s[V_BHN + u];
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    vr[i] = sigmoid_fast(acc[i] + br);
                    vz[i] = sigmoid_fast(acc[4 + i] + bz);
                    vh[i] = acc[12 + i] + bhn;
                    vn[i] = tanhf(acc[8 + i] + bin + vr[i] * vh[i]);
                    const float hp = h_zero ? 0.0f : hcur_s[ug * kRC + rg * kRT + i];
                    hv[i] = (1.0f - vz[i]) * vn[i] + vz[i] * hp;
                }
                bcast_f4_tx(hnext_s + ug * kRC + rg * kRT, make_float4(hv[0], hv[1], hv[2], hv[3]), bar_of(1));
                store4(h_tile(p, g, t, tiles, tile) + ug * kR + row_off + rg * kRT, hv);
                float* gt = p.gates + tb * 4 * kHid * kR + ug * kR + row_off + rg * kRT;
                store4(gt, vr);
                store4(gt + kHid * kR, vz);
                store4(gt + 2 * kHid * kR, vn);
                store4(gt + 3 * kHid * kR, vh);
            }
            arm(1);
        }
        tx_wait(bar_of(1), par);
        PHASE_MARK(0, 4);    // GRU

        // ---- Linear 1 -> LayerNorm -> SiLU -> Dropout --------------------------------------------------------------
        {
            float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
            int k0, k1;
            k_range(kHid, ks, k0, k1);
            dot_rows<kRC>(acc, hnext_s + rg * kRT, img_s + fwd_img_w1(N) + u, k0, k1);
            reduce_ks1_c<kRT>(acc, red_l1_s, ks, slot, chain);
            if (ks == 0) {
                const float bb = vec_s[V_B1 + u];
                bcast_f4_tx(a1_s + ug * kRC + rg * kRT, make_float4(acc[0] + bb, acc[1] + bb, acc[2] + bb, acc[3] + bb), bar_of(2));
            }
            arm(2);
        }
        tx_wait(bar_of(2), par);
        PHASE_MARK(0, 5);    // Linear 1
        ln_silu_drop_fwd_c(drop1, a1_s, stat_s, vec_s + V_LN1G, vec_s + V_LN1B, 0, rank, chain, tid,
                           p.xh1 + tb * kHid * kR, p.d1 + tb * kHid * kR, p.rstd + tb * 2 * kR, row_off);
        PHASE_MARK(0, 6);    // LayerNorm 1

        // ---- Linear 2 -> LayerNorm -> SiLU -> Dropout --------------------------------------------------------------
        {
            float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
            int k0, k1;
            k_range(kHid, ks, k0, k1);
            dot_rows<kRC>(acc, a1_s + rg * kRT, img_s + fwd_img_w2(N) + u, k0, k1);
            reduce_ks1_c<kRT>(acc, red_l23_s, ks, slot, chain);
            if (ks == 0) {
                const float bb = vec_s[V_B2 + u];
                // a2 aliases yc: every CTA's yc reads (its GRU products) ended before it sent h_t
                bcast_f4_tx(a2_s + ug * kRC + rg * kRT, make_float4(acc[0] + bb, acc[1] + bb, acc[2] + bb, acc[3] + bb), bar_of(3));
            }
            arm(3);
        }
        tx_wait(bar_of(3), par);
        PHASE_MARK(0, 7);    // Linear 2
        ln_silu_drop_fwd_c(drop2, a2_s, stat_s, vec_s + V_LN2G, vec_s + V_LN2B, 1, rank, chain, tid,
                           p.xh2 + tb * kHid * kR, p.d2 + tb * kHid * kR, p.rstd + tb * 2 * kR, row_off);
        PHASE_MARK(0, 8);    // LayerNorm 2

        // ---- Linear 3 -> tanh -> Q_{t+1} (model_torch.py:367-380) ---------------------------------------------------
        {
            float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
            const bool mine = u < nu_c;
            int k0, k1;
            k_range(kHid, ks, k0, k1);
            if (mine) dot_rows<kRC>(acc, a2_s + rg * kRT, img_s + fwd_img_w3(N) + u, k0, k1);
            reduce_ks1_c<kRT>(acc, red_l23_s, ks, slot, chain);
            const bool fin = ks == 0 && mine;
            const int n = rank * NU + u;
            float qv[kRT] = {0.f, 0.f, 0.f, 0.f}, dv[kRT] = {0.f, 0.f, 0.f, 0.f};
            if (fin) {
                const float bb = vec_s[V_B3 + u], q0 = vec_s[V_Q0S + u], dq = vec_s[V_DQS + u];
                bool bad = false;
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    dv[i] = tanhf(acc[i] + bb);
                    const float qu = p.relative ? q0 * (1.0f + dq * dv[i]) : fmaf(dq, dv[i], q0);
                    qv[i] = fminf(fmaxf(qu, p.q_min), p.q_max);
                    bad = bad || (b0 + rg * kRT + i < p.B && !finite_f(qu));
                }
                // Q_{t+1} of chain rows 4 rg + {0,1} goes to CTA 2 rg, of rows 4 rg + {2,3} to CTA 2 rg + 1 (their band stage)
                const uint32_t a = smem_u32(q_s + n * kRB);
                st_async_f2(cluster_addr(a, 2u * rg), make_float2(qv[0], qv[1]), cluster_addr(bar_of(4), 2u * rg));
                st_async_f2(cluster_addr(a, 2u * rg + 1u), make_float2(qv[2], qv[3]), cluster_addr(bar_of(4), 2u * rg + 1u));
                if (bad) {   // NaN / Inf: the reference falls back for the whole batch of this ear -> strict replay pass
                    atomicOr(p.flags + t * p.G + g, 1);
                    atomicOr(any_flag, 1);
                }
            }
            finish();        // the next frame's spectrum tile (its cp.async data has long arrived)
            arm(4);
            if (fin) {
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    const int b = b0 + rg * kRT + i;
                    if (b < p.B) {
                        const long long e = (((long long)g * p.B + b) * T + (t + 1)) * N + n;
                        p.Q[e] = qv[i];
                        p.delta[e] = dv[i];
                    }
                }
            }
        }
        tx_wait(bar_of(4), par);   // Q_{t+1} of my band-stage rows has landed ...
        chain_sync(chain);         // ... and every thread of the chain has converted its slots of the next spectrum tile
        PHASE_MARK(0, 9);    // Linear 3 + Q
        hsel ^= 1;
    }
    __syncthreads();
    cluster.sync();   // no CTA leaves (and frees its shared memory) while a peer could still be sending to it
}

// ==================================================================================================
// backward
// ==================================================================================================
struct BwdSmem {   // offsets in floats
    int N;
    __host__ __device__ BwdSmem(int N_) : N(N_) {}
    __host__ __device__ int img() const { return 0; }
    __host__ __device__ int vec() const { return bwd_img_floats(N); }              // LN params, q0, dq
    __host__ __device__ int bufa() const { return vec() + 1024; }                  // [128][32]; first quarter of gate
    __host__ __device__ int gate() const { return bufa(); }                        // [4*128][32]: drp, dzp, dnp, dhn
    __host__ __device__ int dpre() const { return gate() + 4 * kHid * kR; }        // [128][32]; aliased by bufb
    __host__ __device__ int red() const { return dpre() + kHid * kR; }             // (kKS - 1) x 12 x 128
    __host__ __device__ int stat() const { return red() + (kKS - 1) * 12 * 128; }  // 2 x kSeqThreads
    __host__ __device__ int pre() const { return stat() + 2 * kSeqThreads; }       // [4][kR][kU]: ext, jac, fac, jac of the right ear
    __host__ __device__ int bars() const { return pre() + 4 * kR * kU; }           // 4 mbarriers (8 B each)
    __host__ __device__ int total() const { return bars() + 16; }
};
constexpr int VB_LN1G = 0, VB_LN1B = 128, VB_LN2G = 256, VB_LN2B = 384, VB_Q0 = 512, VB_DQ = 640;   // < 1024

// The LayerNorm inputs saved by the forward (this thread's 4 features of its row) and the row's rstd: loaded by the
// caller BEFORE the preceding GEMM phase so that the global-memory latency hides behind it.
struct LnSaved {
    float xh[kHid / (kSeqThreads / kR)];
    float rstd;
};
__device__ __forceinline__ LnSaved load_ln_saved(const float* __restrict__ xh_tile, const float* __restrict__ rstd_tile, int layer) {
    constexpr int PARTS = kSeqThreads / kR, FPP = kHid / PARTS;
    const int row = threadIdx.x % kR, f0 = (threadIdx.x / kR) * FPP;
    LnSaved v;
#pragma unroll
    for (int i = 0; i < FPP; ++i) v.xh[i] = __ldg(xh_tile + (f0 + i) * kR + row);
    v.rstd = __ldg(rstd_tile + layer * kR + row);
    return v;
}

// Backward of Dropout -> SiLU -> LayerNorm on the full rows in buf_s (dL/d(layer output) on entry, dL/d(pre-LayerNorm
// activation) on exit).  Redundant in every CTA; the warp owning the CTA's feature slice writes the saved gradients.
__device__ __forceinline__ void ln_silu_drop_bwd(const BiearSeqParams& p, unsigned long long seed, float* buf_s, float* stat_s,
                                                 const float* __restrict__ gamma_s, const float* __restrict__ beta_s,
                                                 int layer, int t, long long grow0, int rank,
                                                 const float (&xh)[kHid / (kSeqThreads / kR)], float rstd,
                                                 float (&dv_out)[kHid / (kSeqThreads / kR)],
                                                 float (&da_out)[kHid / (kSeqThreads / kR)]) {
    // dv_out / da_out: dL/d(LN output) and dL/d(LN input) of this thread's (row, 4 features); the caller stores the CTA's
    // own slice to HBM (store_ln_grads) AFTER it has signalled the next cluster barrier, whose release would otherwise
    // wait for those stores.
    constexpr int PARTS = kSeqThreads / kR, FPP = kHid / PARTS;
    const int row = threadIdx.x % kR, part = threadIdx.x / kR;
    const int f0 = part * FPP;
    float dxh[FPP];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < FPP / 4; ++i4) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.training) sc = dropout_scale4(seed, t, layer, grow0 + row, (f0 >> 2) + i4);
        const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i4 * 4 + j, f = f0 + i;
            const float gm = gamma_s[f];
            const float y = fmaf(xh[i], gm, beta_s[f]);
            const float sg = sigmoid_fast(y);
            const float d = buf_s[f * kR + row] * scv[j];
            const float dv = d * sg * (1.0f + y * (1.0f - sg));
            dv_out[i] = dv;
            dxh[i] = dv * gm;
            s1 += dxh[i];
            s2 = fmaf(dxh[i], xh[i], s2);
        }
    }
    s1 += __shfl_xor_sync(0xffffffffu, s1, 16);                     // lanes l and l^16 hold two parts of one row
    s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
    float2* stat2 = reinterpret_cast<float2*>(stat_s);               // [row][17] float2
    if ((threadIdx.x & 31) < kR) stat2[row * 17 + (threadIdx.x >> 5)] = make_float2(s1, s2);
    __syncthreads();
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int w = 0; w < kSeqThreads / 32; ++w) {
        const float2 q = stat2[row * 17 + w];
        m1 += q.x;
        m2 += q.y;
    }
    m1 *= (1.0f / kHid);
    m2 *= (1.0f / kHid);
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        const int f = f0 + i;
        const float da = rstd * (dxh[i] - m1 - xh[i] * m2);
        buf_s[f * kR + row] = da;
        da_out[i] = da;
    }
    __syncthreads();
}

__device__ __forceinline__ void store_ln_grads(int rank, const float (&dv)[kHid / (kSeqThreads / kR)],
                                               const float (&da)[kHid / (kSeqThreads / kR)], float* gv_tile, float* ga_tile) {
    constexpr int PARTS = kSeqThreads / kR, FPP = kHid / PARTS;
    const int row = threadIdx.x % kR, f0 = (threadIdx.x / kR) * FPP;
    if (f0 / kU != rank) return;
#pragma unroll
    for (int i = 0; i < FPP; ++i) {
        gv_tile[(f0 + i) * kR + row] = dv[i];
        ga_tile[(f0 + i) * kR + row] = da[i];
    }
}

// SINGLE: the single-controller front-end (model_torch.py:695-776; forward in seq_single.cu): one controller, both ears'
// band outputs feed it and its one Q drives both ears' band stages, so dL/dQ sums the two ears' closed-form terms and the
// transposed GRU input product has two column slices (current features of the left and of the right ear; the memory inputs
// are detached).  Those two slices do not fit in shared memory next to the rest: they are read from L2 (wihc_lr).
template <bool SINGLE>
__global__ void __launch_bounds__(kSeqThreads, 1) seq_bwd_kernel(const BiearSeqParams p, const float* __restrict__ img,
                                                                 const float* __restrict__ wihc_lr) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int N = p.N, T = p.T, S = p.T - 1;
    const int tiles = (p.B + kR - 1) / kR;
    const int NU = bands_per_cta(N);
    const BwdSmem L(N);
    float* img_s = smem + L.img();
    float* vec_s = smem + L.vec();
    float* bufa_s = smem + L.bufa();
    float* gate_s = smem + L.gate();
    float* dpre_s = smem + L.dpre();
    float* bufb_s = dpre_s;
    float* red_s = smem + L.red();
    float* stat_s = smem + L.stat();
    float* pre_s = smem + L.pre();
    const int tid = threadIdx.x;
    const int ks = tid >> 7, slot = tid & 127, rg = slot / kU, u = slot % kU;
    const int ug = rank * kU + u;
    const int nu_c = max(0, min(NU, N - rank * NU));
    const int tl = blockIdx.x / kCS;
    const int g = tl / tiles, tile = tl % tiles;
    const int b0 = tile * kR;
    const long long grow0 = (long long)g * p.B + b0 + rank * kRT;
    const int bb0 = b0 + rank * kRT;
    const unsigned long long seed = p.seed_ptr ? *p.seed_ptr : p.seed;

    if (SINGLE)
        copy_f4(reinterpret_cast<float4*>(img_s),
                reinterpret_cast<const float4*>(img + (long long)rank * single_bwd_res_floats(N)), single_bwd_res_floats(N) / 4);
    else
        copy_f4(reinterpret_cast<float4*>(img_s),
                reinterpret_cast<const float4*>(img + (long long)(g * kCS + rank) * bwd_img_floats(N)), bwd_img_floats(N) / 4);
    const float* wl_g = SINGLE ? wihc_lr + (long long)rank * 2 * 3 * kHid * kU : nullptr;   // W_ih[:, n] (left ear's features)
    const float* wr_g = SINGLE ? wl_g + 3 * kHid * kU : nullptr;                             // W_ih[:, 2N + n] (right ear's)
    for (int i = tid; i < kHid; i += kSeqThreads) {
        vec_s[VB_LN1G + i] = ctrl_ptr(p.ln1_g, g)[i];
        vec_s[VB_LN1B + i] = ctrl_ptr(p.ln1_b, g)[i];
        vec_s[VB_LN2G + i] = ctrl_ptr(p.ln2_g, g)[i];
        vec_s[VB_LN2B + i] = ctrl_ptr(p.ln2_b, g)[i];
        vec_s[VB_Q0 + i] = i < N ? p.q0[i] : 1.0f;
        vec_s[VB_DQ + i] = i < N ? p.dq[i] : 0.0f;
    }
    float dh_carry[kRT] = {0.f, 0.f, 0.f, 0.f};     // dL/dh_t arriving from step t+1 (owned by the ks == 0 threads)
    static_assert(kRT * kHid <= kSeqThreads, "one (row, band) element of the dL/dpre assembly per thread");
    // Part of dL/dpre_{t} that does not depend on the recurrence, for the element (row i, band n) this thread assembles:
    //   ext = gY dY/dQ + gphase dphase/dQ + gQ,  jac = dY/dQ,  fac = d clamp/dQ * dQ/ddelta * d tanh  (all at frame t+1)
    // (SINGLE: index e of gY / gP / gLogY is the EAR, and the row-major (E*B,T,N) tensors hold ear e at row offset e*B)
    constexpr int NE = SINGLE ? 2 : 1;
    const float* gY_e[NE];
    const float* gP_e[NE];
    const float* gLX_e[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        gY_e[e] = SINGLE ? (e ? p.gY[1] : p.gY[0]) : ctrl_ptr(p.gY, g);
        gP_e[e] = SINGLE ? (e ? p.gP[1] : p.gP[0]) : ctrl_ptr(p.gP, g);
        gLX_e[e] = SINGLE ? (e ? p.gLogY[1] : p.gLogY[0]) : ctrl_ptr(p.gLogY, g);
    }
    const float* gQ_g = ctrl_ptr(p.gQ, g);
    struct Pre { float ext, jac, fac, jac2; };
    struct PreRaw { float gy[NE], jac[NE], gp[NE], dp[NE], y[NE], glx[NE], gq, delta; bool live; };
    // issue_pre: only the global loads (so that they are in flight during whatever comes next);
    // finish_pre: the arithmetic, called when the values are needed.
    static_assert(kR * kU == kSeqThreads, "one (tile row, band of the CTA's slice) element per thread");
    const int pre_r = tid / kU, pre_u = tid % kU;                 // element this thread fetches: row pre_r, band rank*NU + pre_u
    auto issue_pre = [&](int t) {
        PreRaw r;
#pragma unroll
        for (int e = 0; e < NE; ++e) r.gy[e] = r.jac[e] = r.gp[e] = r.dp[e] = r.y[e] = r.glx[e] = 0.f;
        r.gq = r.delta = 0.f;
        r.live = false;
        if (t < 0 || pre_u >= nu_c || b0 + pre_r >= p.B) return r;
        const long long el = (((long long)(b0 + pre_r) * T) + (t + 1)) * N + rank * NU + pre_u;   // within one ear's / controller's tensors
        r.live = true;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const long long ee = el + (long long)(SINGLE ? e : g) * p.B * T * N;                    // within the (E*B,T,N) ones
            r.jac[e] = __ldg(p.dYdQ + ee);
            if (gY_e[e]) r.gy[e] = __ldg(gY_e[e] + el);
            if (gLX_e[e]) {
                r.y[e] = __ldg(p.Y + ee);
                r.glx[e] = __ldg(gLX_e[e] + el);
            }
            if (gP_e[e]) {
                r.gp[e] = __ldg(gP_e[e] + el);
                r.dp[e] = __ldg(p.dPdQ + ee);
            }
        }
        if (gQ_g) r.gq = __ldg(gQ_g + el);
        r.delta = __ldg(p.delta + el + (long long)g * p.B * T * N);
        return r;
    };
    auto finish_pre = [&](const PreRaw& w) {
        Pre r = {0.f, 0.f, 0.f, 0.f};
        if (!w.live) return r;
        const int n = rank * NU + pre_u;
        float ext = w.gq;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            float gy = w.gy[e];
            if (gLX_e[e]) {   // d clamp(log(Y + 1e-8), +-12) / dY, as autograd derives it (model_torch.py:1080-1083)
                const float ye = w.y[e] + 1e-8f;
                const float lx = logf(ye);
                if (lx >= -12.0f && lx <= 12.0f) gy += w.glx[e] / ye;
            }
            ext += fmaf(w.gp[e], w.dp[e], gy * w.jac[e]);
        }
        r.jac = w.jac[0];
        r.jac2 = w.jac[NE - 1];
        r.ext = ext;
        const float q0 = vec_s[VB_Q0 + n], dq = vec_s[VB_DQ + n];
        const float qu = p.relative ? q0 * (1.0f + dq * w.delta) : fmaf(dq, w.delta, q0);
        const float scale = p.relative ? q0 * dq : dq;
        r.fac = (qu >= p.q_min && qu <= p.q_max) ? scale * (1.0f - w.delta * w.delta) : 0.f;
        return r;
    };
    auto publish_pre = [&](const Pre& pf) {
        pre_s[pre_r * kU + pre_u] = pf.ext;
        pre_s[kR * kU + pre_r * kU + pre_u] = pf.jac;
        pre_s[2 * kR * kU + pre_r * kU + pre_u] = pf.fac;
        if (SINGLE) pre_s[3 * kR * kU + pre_r * kU + pre_u] = pf.jac2;
    };
    // Hand-overs inside the cluster: st.async + one mbarrier per phase type (seq_dev.cuh): 0 = dL/dpre, 1 = Linear 3^T
    // output (bufa), 2 = Linear 2^T output (bufb), 3 = the four gate-gradient blocks.
    const uint32_t bars = smem_u32(smem + L.bars());
    auto bar_of = [&](int i) { return bars + 8u * (uint32_t)i; };
    const uint32_t tx_bytes[4] = {(uint32_t)(N * kR * 4), (uint32_t)(kHid * kR * 4), (uint32_t)(kHid * kR * 4),
                                  (uint32_t)(4 * kHid * kR * 4)};
    auto arm = [&](int i) {
        if (tid == 0) mbar_arrive_expect_tx(bar_of(i), tx_bytes[i]);
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mbar_init(bar_of(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();                                  // vec_s ready (fetch_pre reads q0 / dq from it); every CTA's barriers exist
    // dL/dpre of step t for (band n of this CTA's slice, rows 4rg..4rg+3), from the pre-fetched recurrence-independent
    // parts in pre_s and dL/dY_{t+1} through the controller (dyc, zero for the last step): broadcast to every CTA of the
    // cluster (the next Linear^T contracts over all bands) and saved for dW3.  Returns the values for the deferred store.
    // (SINGLE: dyc2 = dL/dY_{t+1} of the right ear through the controller, paired with the right ear's dY/dQ)
    auto push_dpre = [&](int t, const float dyc[kRT], const float dyc2[kRT], float (&dp)[kRT]) {
        const bool flagged_t = p.flags[t * p.G + g] != 0;           // Q_{t+1} was replaced by Q0: no gradient through it
#pragma unroll
        for (int i = 0; i < kRT; ++i) {
            const int r = rg * kRT + i;
            float v = pre_s[r * kU + u] + dyc[i] * pre_s[kR * kU + r * kU + u];
            if (SINGLE) v = fmaf(dyc2[i], pre_s[3 * kR * kU + r * kU + u], v);
            dp[i] = flagged_t ? 0.f : v * pre_s[2 * kR * kU + r * kU + u];
        }
        broadcast_rows_tx(dpre_s, rank * NU + u, rg * kRT, dp, bar_of(0));
    };
    {   // prologue: dL/dpre of the last step (nothing arrives through a later controller step)
        publish_pre(finish_pre(issue_pre(S - 1)));
        __syncthreads();
        if (ks == 0 && u < nu_c) {
            const float zero[kRT] = {0.f, 0.f, 0.f, 0.f};
            float dp[kRT];
            push_dpre(S - 1, zero, zero, dp);
            store4(p.G_pre + tile_base(p, g, S - 1, tiles, tile) * N * kR + (rank * NU + u) * kR + rg * kRT, dp);
        }
        arm(0);
        tx_wait(bar_of(0), 0);       // hand-over 0 completes once in the prologue, then once per step with t > 0
    }

    PHASE_INIT();
    for (int t = S - 1; t >= 0; --t) {
        PHASE_MARK(1, 0);
        const bool flagged = p.flags[t * p.G + g] != 0;                    // Q_{t+1} was replaced by Q0, h_t dropped
        const bool h_reset = (t == 0) || (p.flags[(t - 1) * p.G + g] != 0);
        const long long tb = tile_base(p, g, t, tiles, tile);
        PHASE_MARK(1, 1);    // (dL/dpre of this step was assembled and pushed at the end of the previous one)

        // ---- Linear 3 ^T --------------------------------------------------------------------------------------
        const LnSaved ln2 = load_ln_saved(p.xh2 + tb * kHid * kR, p.rstd + tb * 2 * kR, 1);   // used after #2
        {
            float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
            int k0, k1;
            k_range(N, ks, k0, k1);
            dot_rows(acc, dpre_s + rg * kRT, img_s + bwd_img_w3c(N) + u, k0, k1);
            reduce_ks1<kRT>(acc, red_s, ks, slot);
            if (ks == 0) broadcast_rows_tx(bufa_s, ug, rg * kRT, acc, bar_of(1));
        }
        const uint32_t par = (uint32_t)(S - 1 - t) & 1u;      // hand-overs 1..3 complete exactly once per step
        arm(1);
        tx_wait(bar_of(1), par);   // #2
        PHASE_MARK(1, 2);    // Linear 3 ^T
        constexpr int kFPP = kHid / (kSeqThreads / kR);
        float dv2[kFPP], da2[kFPP];
        ln_silu_drop_bwd(p, seed, bufa_s, stat_s, vec_s + VB_LN2G, vec_s + VB_LN2B, 1, t, (long long)g * p.B + b0, rank,
                         ln2.xh, ln2.rstd, dv2, da2);
        const LnSaved ln1 = load_ln_saved(p.xh1 + tb * kHid * kR, p.rstd + tb * 2 * kR, 0);   // used after #3
        PHASE_MARK(1, 3);    // LayerNorm 2 backward
        // ---- Linear 2 ^T --------------------------------------------------------------------------------------
        {
            float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
            int k0, k1;
            k_range(kHid, ks, k0, k1);
            dot_rows(acc, bufa_s + rg * kRT, img_s + bwd_img_w2c(N) + u, k0, k1);
            reduce_ks1<kRT>(acc, red_s, ks, slot);
            // bufb aliases dpre: every CTA's dpre reads (its Linear 3^T products) ended before it sent its part of bufa
            if (ks == 0) broadcast_rows_tx(bufb_s, ug, rg * kRT, acc, bar_of(2));
        }
        arm(2);
        store_ln_grads(rank, dv2, da2, p.G_v2 + tb * kHid * kR, p.G_a2 + tb * kHid * kR);
        tx_wait(bar_of(2), par);   // #3
        PHASE_MARK(1, 4);    // Linear 2 ^T
        float dv1[kFPP], da1[kFPP];
        ln_silu_drop_bwd(p, seed, bufb_s, stat_s, vec_s + VB_LN1G, vec_s + VB_LN1B, 0, t, (long long)g * p.B + b0, rank,
                         ln1.xh, ln1.rstd, dv1, da1);
        PHASE_MARK(1, 5);    // LayerNorm 1 backward
        // ---- Linear 1 ^T, GRU cell backward ------------------------------------------------------------------------
        float dh_direct[kRT] = {0.f, 0.f, 0.f, 0.f};
        {
            // the saved gates / previous state of (unit ug, rows 4rg..4rg+3): fetched before the product, used after it
            float4 r4, z4, n4, h4, p4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ks == 0) {
                const float* gt = p.gates + tb * 4 * kHid * kR + ug * kR + rg * kRT;
                r4 = __ldg(reinterpret_cast<const float4*>(gt));
                z4 = __ldg(reinterpret_cast<const float4*>(gt + kHid * kR));
                n4 = __ldg(reinterpret_cast<const float4*>(gt + 2 * kHid * kR));
                h4 = __ldg(reinterpret_cast<const float4*>(gt + 3 * kHid * kR));
                if (!h_reset)
                    p4 = __ldg(reinterpret_cast<const float4*>(h_tile(p, g, t - 1, tiles, tile) + ug * kR + rg * kRT));
            }
            float acc[kRT] = {0.f, 0.f, 0.f, 0.f};
            int k0, k1;
            k_range(kHid, ks, k0, k1);
            dot_rows(acc, bufb_s + rg * kRT, img_s + bwd_img_w1c(N) + u, k0, k1);
            reduce_ks1<kRT>(acc, red_s, ks, slot);
            float v0[kRT], v1[kRT], v2[kRT], v3[kRT];
            if (ks == 0) {
                const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
                const float nn[4] = {n4.x, n4.y, n4.z, n4.w}, hn[4] = {h4.x, h4.y, h4.z, h4.w};
                const float hp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    float dh = acc[i];
                    if (t < S - 1 && !flagged) dh += dh_carry[i];
                    const float dn = dh * (1.0f - zz[i]);
                    const float dz = dh * (hp[i] - nn[i]);
                    dh_direct[i] = dh * zz[i];
                    const float dnp = dn * (1.0f - nn[i] * nn[i]);
                    v0[i] = dnp * hn[i] * rr[i] * (1.0f - rr[i]);      // dL/d r_pre
                    v1[i] = dz * zz[i] * (1.0f - zz[i]);                // dL/d z_pre
                    v2[i] = dnp;                                         // dL/d (i_n pre-activation)
                    v3[i] = dnp * rr[i];                                 // dL/d (W_hn h + b_hn)
                }
                // gate_s[0:128] aliases bufa: every CTA's bufa reads (its Linear 2^T products) ended before it sent bufb
                broadcast_rows_tx(gate_s, 0 * kHid + ug, rg * kRT, v0, bar_of(3));
                broadcast_rows_tx(gate_s, 1 * kHid + ug, rg * kRT, v1, bar_of(3));
                broadcast_rows_tx(gate_s, 2 * kHid + ug, rg * kRT, v2, bar_of(3));
                broadcast_rows_tx(gate_s, 3 * kHid + ug, rg * kRT, v3, bar_of(3));
            }
            arm(3);
            store_ln_grads(rank, dv1, da1, p.G_v1 + tb * kHid * kR, p.G_a1 + tb * kHid * kR);
            if (ks == 0) {
                float* gg = p.GG + tb * 4 * kHid * kR + ug * kR + rg * kRT;
                store4(gg, v0);
                store4(gg + kHid * kR, v1);
                store4(gg + 2 * kHid * kR, v2);
                store4(gg + 3 * kHid * kR, v3);
            }
        }
        tx_wait(bar_of(3), par);   // #4
        PHASE_MARK(1, 6);    // Linear 1 ^T + GRU cell backward
        const PreRaw pf_raw = issue_pre(t - 1);    // next step's recurrence-independent inputs: in flight during the products below
        PHASE_MARK(1, 8);    // last phase: issue the next step's loads
        // ---- dL/dh_{t-1} = z * dh + W_hh^T [drp, dzp, dhn];  dL/dY_t = W_ih[:, :N]^T [drp, dzp, dnp] * d log1p --------
        {
            constexpr int NA = SINGLE ? 3 : 2;         // dh, dL/dc (left | only ear)[, dL/dc right ear]
            float acc[NA * kRT];
#pragma unroll
            for (int i = 0; i < NA * kRT; ++i) acc[i] = 0.f;
            const float* x = gate_s + rg * kRT;
            const float* whhc = img_s + bwd_img_whhc(N) + u;
            // rows o of W_hh: [0,128) r gate, [128,256) z gate, [256,384) n gate (pairs with dhn = gate block 3, i.e.
            // gate_s rows o + 128); rows o of W_ih pair with gate_s rows o (drp, dzp, dnp)
            int o0, o1;
            k_range(3 * kHid, ks, o0, o1);
            if (SINGLE) {
                // the two W_ih column slices come from L2 (coalesced over u, 16 loads in flight per thread)
                const float* wl = wl_g + u;
                const float* wr = wr_g + u;
                float2 hl = make_float2(0.f, 0.f), hh = make_float2(0.f, 0.f), ll = hl, lh = hl, rl = hl, rh = hl;
#ifdef BIEAR_SKIP_DOTS
                o1 = o0;
#endif
#pragma unroll 8
                for (int o = o0; o < o1; ++o) {
                    const float wlv = __ldg(wl + o * kU), wrv = __ldg(wr + o * kU), whv = whhc[o * kU];
                    const float4 xg = *reinterpret_cast<const float4*>(x + o * kR);
                    const float4 xh = o < 2 * kHid ? xg : *reinterpret_cast<const float4*>(x + (o + kHid) * kR);
                    const float2 pl = make_float2(wlv, wlv), pr = make_float2(wrv, wrv), ph = make_float2(whv, whv);
                    ll = __ffma2_rn(pl, make_float2(xg.x, xg.y), ll); lh = __ffma2_rn(pl, make_float2(xg.z, xg.w), lh);
                    rl = __ffma2_rn(pr, make_float2(xg.x, xg.y), rl); rh = __ffma2_rn(pr, make_float2(xg.z, xg.w), rh);
                    hl = __ffma2_rn(ph, make_float2(xh.x, xh.y), hl); hh = __ffma2_rn(ph, make_float2(xh.z, xh.w), hh);
                }
                acc[0] = hl.x; acc[1] = hl.y; acc[2] = hh.x; acc[3] = hh.y;
                acc[4] = ll.x; acc[5] = ll.y; acc[6] = lh.x; acc[7] = lh.y;
                acc[NA * kRT - 4] = rl.x; acc[NA * kRT - 3] = rl.y; acc[NA * kRT - 2] = rh.x; acc[NA * kRT - 1] = rh.y;
            } else {
                const float* wihc = img_s + bwd_img_wihc(N) + u;
                // o < 256: both products read gate_s row o (one load); o >= 256: W_hh pairs with row o + 128, W_ih with row o.
                // (Bands beyond the CTA's slice have zero W_ih columns in the image: computing them is harmless.)
                dot_rows_pair(acc, acc + kRT, x, whhc, wihc, min(o0, 2 * kHid), min(o1, 2 * kHid));
                dot_rows(acc, x + kHid * kR, whhc, max(o0, 2 * kHid), max(o1, 2 * kHid));
                dot_rows(acc + kRT, x, wihc, max(o0, 2 * kHid), max(o1, 2 * kHid));
            }
            PHASE_MARK(1, 9);    // last phase: the K = 384 transposed products
            const bool mine = u < nu_c;
            const int n = rank * NU + u;
            float yv[kRT] = {0.f, 0.f, 0.f, 0.f};       // Y_t of this thread's 4 rows (for d log1p): fetched before the products
            float yv2[kRT] = {0.f, 0.f, 0.f, 0.f};      // (SINGLE: the right ear's)
            if (ks == 0 && mine) {
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    const int b = b0 + rg * kRT + i;
                    if (b < p.B) {
                        yv[i] = __ldg(p.Y + (((long long)g * p.B + b) * T + t) * N + n);
                        if (SINGLE) yv2[i] = __ldg(p.Y + (((long long)p.B + b) * T + t) * N + n);
                    }
                }
            }
            // the recurrence-independent parts of the NEXT step's dL/dpre (step t-1) have arrived: publish them CTA-wide
            if (t > 0) publish_pre(finish_pre(pf_raw));
            PHASE_MARK(1, 10);   // last phase: Y_t loads + finish_pre (waits for the loads issued above)
            reduce_ks1<NA * kRT>(acc, red_s, ks, slot);   // (its barriers also order the pre_s writes before the reads below)
            PHASE_MARK(1, 11);   // last phase: k-split reduction
            float dp[kRT] = {0.f, 0.f, 0.f, 0.f};
            const bool fin = ks == 0 && mine && t > 0;
            if (ks == 0) {
#pragma unroll
                for (int i = 0; i < kRT; ++i) dh_carry[i] = acc[i] + dh_direct[i];
                if (fin) {   // dL/dY_t through the controller (d log1p), straight into dL/dpre of step t-1: no exchange needed
                    float dy[kRT], dy2[kRT];
#pragma unroll
                    for (int i = 0; i < kRT; ++i) {
                        dy[i] = yv[i] >= 0.0f ? acc[kRT + i] / (1.0f + yv[i]) : 0.0f;
                        dy2[i] = SINGLE ? (yv2[i] >= 0.0f ? acc[(NA - 1) * kRT + i] / (1.0f + yv2[i]) : 0.0f) : 0.0f;
                    }
                    push_dpre(t - 1, dy, dy2, dp);
                }
            }
            PHASE_MARK(1, 12);   // last phase: dh carry + dL/dpre assembly + push
            if (t > 0) arm(0);          // #5: dL/dpre of step t-1 is on its way everywhere
            if (fin) store4(p.G_pre + tile_base(p, g, t - 1, tiles, tile) * N * kR + n * kR + rg * kRT, dp);
        }
        if (t > 0) tx_wait(bar_of(0), (uint32_t)(S - t) & 1u);   // (prologue = completion 0, step t = completion S - t)
        PHASE_MARK(1, 7);    // W_hh^T / W_ih^T products
    }
    cluster.sync();   // no CTA leaves (and frees its shared memory) while a peer could still be sending to it
}

// ==================================================================================================
// host side
// ==================================================================================================
static int validate_seq(const BiearSeqParams* p, const char* who, bool backward) {
    BIEAR_REQUIRE(p != nullptr, "%s: null parameter block", who);
    BIEAR_REQUIRE(p->G >= 1 && p->B >= 1 && p->T >= 1 && p->N >= 1 && p->N <= kHid && p->F >= 2,
                  "%s: bad geometry G=%d B=%d T=%d N=%d F=%d", who, p->G, p->B, p->T, p->N, p->F);
    const bool single = p->G == 1 && p->E == 2;
    BIEAR_REQUIRE((p->E == p->G && p->Kin == 2 * p->N) || (single && p->Kin == 4 * p->N),
                  "%s: fused are the dual front-end (one controller per ear, E == G, Kin = 2N) and the single-controller one "
                  "(G = 1, E = 2, Kin = 4N); got E=%d G=%d Kin=%d N=%d", who, p->E, p->G, p->Kin, p->N);
    BIEAR_REQUIRE(!single || !p->x_ready, "%s: the single-controller kernel does not take streamed spectra (x_ready)", who);
    BIEAR_REQUIRE(p->G <= BIEAR_MAX_CTRL, "%s: at most %d controllers per call, got %d", who, BIEAR_MAX_CTRL, p->G);
    BIEAR_REQUIRE(p->fc && p->q0 && p->dq, "%s: null constant pointer", who);
    for (int g = 0; g < p->G; ++g)
        BIEAR_REQUIRE(p->w_ih[g] && p->w_hh[g] && p->b_ih[g] && p->b_hh[g] && p->w1[g] && p->b1[g] && p->ln1_g[g] &&
                          p->ln1_b[g] && p->w2[g] && p->b2[g] && p->ln2_g[g] && p->ln2_b[g] && p->w3[g] && p->b3[g],
                      "%s: null weight pointer of controller %d", who, g);
    BIEAR_REQUIRE(p->Y && p->Q && p->delta && p->dYdQ && p->flags && p->workspace, "%s: null output pointer", who);
    BIEAR_REQUIRE(!p->phase == !p->dPdQ, "%s: phase and dPdQ must be given together", who);
    if (p->T > 1)
        BIEAR_REQUIRE(p->H && p->gates && p->xh1 && p->d1 && p->xh2 && p->d2 && p->rstd && p->yc,
                      "%s: null saved-state pointer", who);
    if (backward) {
        BIEAR_REQUIRE(p->GG && p->G_a1 && p->G_v1 && p->G_a2 && p->G_v2 && p->G_pre, "%s: null backward buffer", who);
        for (int g = 0; g < p->G; ++g)
            BIEAR_REQUIRE(!p->gP[g] || p->dPdQ, "%s: gphase given but the forward saved no dphase/dQ", who);
    } else {
        BIEAR_REQUIRE(p->X, "%s: null spectra", who);
    }
    return 0;
}

template <typename Kern, typename... Extra>
static int launch_cluster(Kern kern, const char* name, int clusters, size_t smem, cudaStream_t st,
                          const BiearSeqParams& p, const float* img, Extra... extra) {
    // the opt-in is per device and sticky; remember the largest size set so that steady-state launches (and launches
    // recorded into a CUDA graph) make no attribute call
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
    cfg.blockDim = dim3(kSeqThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = check_cuda(cudaLaunchKernelEx(&cfg, kern, p, img, extra...), name);
    if (e) return e;
    count_launch();
    return 0;
}

}  // namespace biear

extern "C" int biear_adaptive_tile_rows(void) { return biear::kR; }

extern "C" int biear_adaptive_supported(int N, int F) {
    using namespace biear;
    if (N < 1 || N > kHid || F < 2) return 0;
    const size_t limit = 227 * 1024;
    return sizeof(float) * (size_t)FwdSmem(N, F).total() <= limit && sizeof(float) * (size_t)Fwd2Smem(N, F).total() <= limit &&
           sizeof(float) * (size_t)BwdSmem(N).total() <= limit;
}

extern "C" int biear_single_supported(int N, int F) {
    using namespace biear;
    if (N < 1 || N > kHid || F < 2) return 0;
    const size_t limit = 227 * 1024;
    return single_fwd_smem_bytes(N, F) <= limit && sizeof(float) * (size_t)BwdSmem(N).total() <= limit;
}

extern "C" int64_t biear_single_workspace_floats(int N) {
    using namespace biear;
    if (N < 1 || N > kHid) return 0;
    return (int64_t)single_workspace_floats(N);
}

extern "C" int64_t biear_adaptive_workspace_floats(int G, int N) {
    using namespace biear;
    if (G < 1 || N < 1 || N > kHid) return 0;
    return (int64_t)G * kCS * (fwd_img_floats(N) + bwd_img_floats(N));
}

namespace biear {
static int launch_prepare(const BiearSeqParams* p, int want, cudaStream_t st) {
    prepare_kernel<<<dim3(2 * p->G * kCS + 1, kPackSlices), 256, 0, st>>>(*p, p->workspace, want);
    BIEAR_LAUNCH_CHECK("prepare_kernel");
    return 0;
}
}  // namespace biear

extern "C" int biear_adaptive_prepare(const BiearSeqParams* p, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(p != nullptr, "biear_adaptive_prepare: null parameter block");
    const bool single = p->G == 1 && p->E == 2 && p->Kin == 4 * p->N;
    BIEAR_REQUIRE(p->G >= 1 && p->G <= BIEAR_MAX_CTRL && p->B >= 1 && p->T >= 1 && p->N >= 1 && p->N <= kHid &&
                      (single || p->Kin == 2 * p->N),
                  "biear_adaptive_prepare: bad geometry G=%d B=%d T=%d N=%d Kin=%d", p->G, p->B, p->T, p->N, p->Kin);
    for (int g = 0; g < p->G; ++g)
        BIEAR_REQUIRE(p->w_ih[g] && p->w_hh[g] && p->w1[g] && p->w2[g] && p->w3[g],
                      "biear_adaptive_prepare: null weight pointer of controller %d", g);
    BIEAR_REQUIRE(p->workspace && p->H && p->flags, "biear_adaptive_prepare: null workspace / H / flags");
    if (single) return launch_single_prepare(p, 3, as_stream(stream));
    return launch_prepare(p, 3, as_stream(stream));
}

extern "C" int biear_adaptive_fwd(const BiearSeqParams* p, void* stream) {
    using namespace biear;
    if (int e = validate_seq(p, "biear_adaptive_fwd", false)) return e;
    cudaStream_t st = as_stream(stream);
    if (p->G == 1 && p->E == 2) {   // single controller (seq_single.cu)
        if (!p->prepared)
            if (int e = launch_single_prepare(p, 1, st)) return e;
        return launch_single_fwd(p, st);
    }
    const size_t smem = sizeof(float) * (size_t)FwdSmem(p->N, p->F).total();
    BIEAR_REQUIRE(smem <= 227 * 1024, "biear_adaptive_fwd: N=%d F=%d needs %zu B of shared memory", p->N, p->F, smem);
    const int tiles = (p->B + kR - 1) / kR;
    if (!p->prepared)
        if (int e = launch_prepare(p, 1, st)) return e;
    if (!p->force_strict) {   // fast pass: two chains of 8 rows per cluster
        const size_t smem2 = sizeof(float) * (size_t)Fwd2Smem(p->N, p->F).total();
        BIEAR_REQUIRE(smem2 <= 227 * 1024, "biear_adaptive_fwd: N=%d F=%d needs %zu B of shared memory", p->N, p->F, smem2);
        if (int e = launch_cluster(seq_fwd2_kernel, "seq_fwd2_kernel", p->G * tiles, smem2, st, *p, p->workspace)) return e;
    }
    // replay with batch-global fallback semantics; returns immediately unless a non-finite Q was recorded
    return launch_cluster(seq_fwd_strict_kernel, "seq_fwd_strict_kernel", 1, smem, st, *p, p->workspace);
}

extern "C" int biear_adaptive_bwd(const BiearSeqParams* p, void* stream) {
    using namespace biear;
    if (int e = validate_seq(p, "biear_adaptive_bwd", true)) return e;
    if (p->T < 2) return 0;
    cudaStream_t st = as_stream(stream);
    const size_t smem = sizeof(float) * (size_t)BwdSmem(p->N).total();
    BIEAR_REQUIRE(smem <= 227 * 1024, "biear_adaptive_bwd: N=%d needs %zu B of shared memory", p->N, smem);
    const int tiles = (p->B + kR - 1) / kR;
    if (p->G == 1 && p->E == 2) {   // single controller
        if (!p->prepared)
            if (int e = launch_single_prepare(p, 2, st)) return e;
        return launch_cluster(seq_bwd_kernel<true>, "seq_bwd_kernel<single>", tiles, smem, st, *p,
                              (const float*)(p->workspace + single_off_bres(p->N)),
                              (const float*)(p->workspace + single_off_bstr(p->N)));
    }
    if (!p->prepared)
        if (int e = launch_prepare(p, 2, st)) return e;
    return launch_cluster(seq_bwd_kernel<false>, "seq_bwd_kernel", p->G * tiles, smem, st, *p,
                          (const float*)(p->workspace + bwd_images_offset(p->G, p->N)), (const float*)nullptr);
}

// Diagnostics: how many clusters of the persistent kernels can be resident at once on the current device.
extern "C" int biear_adaptive_occupancy(int N, int F, int* fwd_clusters, int* bwd_clusters) {
    using namespace biear;
    BIEAR_REQUIRE(N >= 1 && N <= kHid && F >= 2 && fwd_clusters && bwd_clusters, "biear_adaptive_occupancy: bad arguments");
    const size_t smem_f = sizeof(float) * (size_t)Fwd2Smem(N, F).total();
    const size_t smem_b = sizeof(float) * (size_t)BwdSmem(N).total();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    for (int pass = 0; pass < 2; ++pass) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kCS * 64);
        cfg.blockDim = dim3(kSeqThreads);
        cfg.dynamicSmemBytes = pass ? smem_b : smem_f;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e;
        if (pass == 0) {
            e = cudaFuncSetAttribute(seq_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, seq_fwd2_kernel, &cfg);
        } else {
            e = cudaFuncSetAttribute(seq_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, seq_bwd_kernel<false>, &cfg);
        }
        if (int rc = check_cuda(e, "cudaOccupancyMaxActiveClusters")) return rc;
        *(pass ? bwd_clusters : fwd_clusters) = n;
    }
    return 0;
}

// Diagnostic builds only (-DBIEAR_PHASE_PROF): copy out and clear the per-phase cycle counters (2 x 16); returns
// BIEAR_EINVAL in normal builds.
extern "C" int biear_debug_phase_cycles(unsigned long long* out_host) {
#ifdef BIEAR_PHASE_PROF
    using namespace biear;
    unsigned long long zero[2][16] = {};
    unsigned long long zero_b[8] = {};
    if (int e = check_cuda(cudaMemcpyFromSymbol(out_host, g_phase_cycles, sizeof(zero)), "cudaMemcpyFromSymbol")) return e;
    if (int e = check_cuda(cudaMemcpyFromSymbol(out_host + 32, g_band_prof, sizeof(zero_b)), "cudaMemcpyFromSymbol")) return e;
    if (int e = check_cuda(cudaMemcpyToSymbol(g_band_prof, zero_b, sizeof(zero_b)), "cudaMemcpyToSymbol")) return e;
    return check_cuda(cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)), "cudaMemcpyToSymbol");
#else
    (void)out_host;
    return biear::fail_invalid("biear_debug_phase_cycles: library built without BIEAR_PHASE_PROF");
#endif
}

// The same for the single-controller forward kernel (16 counters), see csrc/seq_single.cu.
extern "C" int biear_debug_phase_cycles_single(unsigned long long* out_host) { return biear::debug_phase_cycles_single(out_host); }
