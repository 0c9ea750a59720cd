// Persistent forward recurrence of the SINGLE-controller binaural front-end
//   model_torch.py:695-776   BinauralAdaptiveGammatoneFB_SingleController.forward: one Q for both ears per clip,
//                            controller input [YLc, YLmem, YRc, YRmem] (4N wide), GRU(4N -> 128) + the same MLP,
//                            carried memory Y_mem <- 0.8 Y_mem + 0.2 Y_ctrl.detach(), batch-global non-finite fallback
//   model_torch.py:1039-1063 the sub-band phase from the same W(Q_t)
// as ONE cluster kernel, like seq_fwd2_kernel does for the dual front-end (seq.cu) -- but with a different thread layout
// and with the GRU input weights STREAMED, because 384 x 4N fp32 (154 KB per CTA of a 4-CTA cluster at N = 100) do not
// fit next to the rest of the controller in the 227 KB of an SM:
//
//   * cluster = 4 CTAs x 512 threads = 8 clips (half a 16-row tile of the saved-state tensors) for all T frames;
//     CTA c runs the band stage of clips 2c, 2c+1 for BOTH ears (4 items per frame, the same band work per CTA as the
//     dual kernel) with the clip's one Q vector;
//   * W_hh, Linear 1-3 stay resident in shared memory (96 KB); this CTA's slice of W_ih (its 32 hidden units x 3 gates x
//     4N inputs) cycles through a 4-stage ring of 15 KB bulk-async copies (cp.async.bulk + mbarrier transaction bytes)
//     issued by one thread: the weights do not depend on the data, so the ring runs ahead across frame boundaries and the
//     GRU product never waits for L2 in steady state;
//   * GEMM phases: thread = (row group of 4 clips, unit, k-split) with the EIGHT K-SPLITS IN ADJACENT LANES: the partial
//     sums are combined with three shuffle steps -- no shared-memory reduction, no block barrier inside a phase.  Bank
//     conflicts are avoided by a skew in the weight images (column (u + 4 (k & 7)) & 31) and by keeping the activations
//     as [row group][feature][4 clips];
//   * hand-overs between the CTAs: st.async + mbarrier transaction bytes (seq_dev.cuh), five per frame;
//   * the carried memory lives in every CTA's input buffer next to the current features and is advanced there.
//
// STRICT = true is the replay pass with the reference's batch-global fallback semantics (one cluster walks all half-tiles
// frame by frame, state through global memory); it exits at once unless the fast pass recorded a non-finite Q.
#include <map>
#include <mutex>
#include <utility>

#include "band_dev.cuh"
#include "seq_dev.cuh"

namespace biear {
namespace sc {

// Optional per-phase cycle accounting (diagnostic builds: -DBIEAR_PHASE_PROF, `make prof`), see seq.cu.
#ifdef BIEAR_PHASE_PROF
__device__ unsigned long long g_phase_cycles1[16];
#define PHASE1_INIT() long long _ph_last = clock64()
#define PHASE1_MARK(i)                                                 \
    do {                                                               \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                     \
            const long long _now = clock64();                          \
            g_phase_cycles1[i] += (unsigned long long)(_now - _ph_last); \
            _ph_last = _now;                                           \
        }                                                              \
    } while (0)
#else
#define PHASE1_INIT() do {} while (0)
#define PHASE1_MARK(i) do {} while (0)
#endif

constexpr int kRows = 8;                 // clips per cluster
constexpr int kRG = kRows / kRT;         // 2 row groups of 4 clips
constexpr int kKL = 8;                   // k-splits = adjacent lanes
constexpr int kItems = 4;                // band-stage items per CTA: 2 clips x 2 ears
constexpr int kClipsB = 2;               // band-stage clips per CTA
constexpr int kStages = 4;               // W_ih ring
constexpr int kWarps = kSeqThreads / 32;
static_assert(kRG * kU * kKL == kSeqThreads, "thread layout of the GEMM phases");
static_assert(kClipsB * kCS == kRows, "band-stage ownership");

// vec area (per-CTA constants)
constexpr int V_BR = 0, V_BZ = kU, V_BIN = 2 * kU, V_BHN = 3 * kU, V_B1 = 4 * kU, V_B2 = 5 * kU, V_B3 = 6 * kU,
              V_Q0S = 7 * kU, V_DQS = 8 * kU;
constexpr int V_LN1G = 9 * kU, V_LN1B = V_LN1G + kHid, V_LN2G = V_LN1B + kHid, V_LN2B = V_LN2G + kHid, V_FC = V_LN2B + kHid,
              V_Q0 = V_FC + kHid, V_END = V_Q0 + kHid;
static_assert(V_END <= 1088, "vec area");

struct Smem {   // offsets in floats
    int N, Kp, CK, tile;
    __host__ __device__ Smem(int N_, int F) : N(N_), Kp(single_kp(N_)), CK(single_chunk_rows(N_)), tile(spec_tile_len(F)) {}
    __host__ __device__ int img() const { return 0; }                           // resident: whh3, w1, w2, w3 (skewed)
    __host__ __device__ int vec() const { return single_fwd_res_floats(); }
    __host__ __device__ int in() const { return vec() + 1088; }                 // [kRG][Kp][4]: mL | mR | cL | cR (+ zero pad)
    __host__ __device__ int h() const { return in() + Kp * kRows; }             // 2 x [kRG][128][4]
    __host__ __device__ int a1() const { return h() + 2 * kHid * kRows; }
    __host__ __device__ int a2() const { return a1() + kHid * kRows; }
    __host__ __device__ int stat() const { return a2() + kHid * kRows; }        // [kRG][4][8] float2
    __host__ __device__ int q() const { return stat() + 2 * kRG * kRT * 8; }    // [128][2]
    __host__ __device__ int ystage() const { return q() + kHid * kClipsB; }     // [kItems][128]
    __host__ __device__ int spec() const { return ystage() + kItems * kHid; }   // [kItems][tile] float4
    __host__ __device__ int ring() const { return spec() + kItems * tile * 4; } // kStages x CK x 96
    __host__ __device__ int bars() const { return ring() + kStages * CK * 3 * kU; }   // 5 + 2 kStages mbarriers
    __host__ __device__ int total() const { return bars() + 32; }
};
static_assert(5 + 2 * kStages <= 16, "mbarrier area");

__device__ __forceinline__ int skew_col(int u, int k) { return (u + 4 * (k & 7)) & 31; }

// ---- weight images -------------------------------------------------------------------------------------------------------
// forward, resident part of CTA rank c:  whh3 [k<128][gate<3][32], w1 | w2 | w3 [k<128][32], columns skewed
// forward, streamed part:                wih3 [k<Kp][gate<3][32] with the input order mL | mR | cL | cR (torch: cL mL cR mR):
//                                        the memory half first -- it does not depend on the current frame's band stage
// backward, resident part:               w3c [n<N][32] | w2c | w1c [o<128][32] | whhc [o<384][32]   (as seq_dev.cuh)
// backward, streamed part:               wihcL [o<384][32] | wihcR [o<384][32]: columns n and 2N + n of W_ih (the memory
//                                        inputs are detached: no gradient flows through their columns)
__device__ __forceinline__ int torch_col(int k, int N) {   // my input order -> torch column of weight_ih
    const int part = k / N, n = k - part * N;          // kernel order: mL, mR, cL, cR
    const int tpart = part == 0 ? 1 : (part == 1 ? 3 : (part == 2 ? 0 : 2));
    return tpart * N + n;
}

__global__ void __launch_bounds__(256) prepare_single_kernel(const BiearSeqParams p, float* __restrict__ ws, int want) {
    const int N = p.N, Kp = single_kp(N), NU = bands_per_cta(N);
    const int tid0 = blockIdx.y * blockDim.x + threadIdx.x, stride = gridDim.y * blockDim.x;
    const int bx = blockIdx.x;
    const float *w_ih = p.w_ih[0], *w_hh = p.w_hh[0], *w1 = p.w1[0], *w2 = p.w2[0], *w3 = p.w3[0];
    if (bx < kCS) {
        if (!(want & 1)) return;
        const int c = bx;
        float* out = ws + single_off_fres() + (long long)c * single_fwd_res_floats();
        for (int idx = tid0; idx < kU * 3 * kHid; idx += stride) {
            const int k = idx % kHid, gu = idx / kHid, gate = gu % 3, u = gu / 3;
            out[(k * 3 + gate) * kU + skew_col(u, k)] = w_hh[(gate * kHid + c * kU + u) * kHid + k];
        }
        float* o1 = out + kHid * 3 * kU;
        for (int idx = tid0; idx < kU * kHid; idx += stride) {
            const int k = idx % kHid, u = idx / kHid, col = skew_col(u, k);
            o1[k * kU + col] = w1[(c * kU + u) * kHid + k];
            o1[kHid * kU + k * kU + col] = w2[(c * kU + u) * kHid + k];
            const int n = c * NU + u;
            o1[2 * kHid * kU + k * kU + col] = (u < NU && n < N) ? w3[n * kHid + k] : 0.f;
        }
    } else if (bx < 2 * kCS) {
        if (!(want & 1)) return;
        const int c = bx - kCS;
        float* out = ws + single_off_fstr(N) + (long long)c * Kp * 3 * kU;
        for (int idx = tid0; idx < kU * 3 * Kp; idx += stride) {
            const int k = idx % Kp, gu = idx / Kp, gate = gu % 3, u = gu / 3;
            out[(k * 3 + gate) * kU + skew_col(u, k)] =
                k < 4 * N ? w_ih[(long long)(gate * kHid + c * kU + u) * p.Kin + torch_col(k, N)] : 0.f;
        }
    } else if (bx < 3 * kCS) {
        if (!(want & 2)) return;
        const int c = bx - 2 * kCS;
        float* out = ws + single_off_bres(N) + (long long)c * single_bwd_res_floats(N);
        for (int idx = tid0; idx < N * kU; idx += stride)
            out[bwd_img_w3c(N) + idx] = w3[(idx / kU) * kHid + c * kU + (idx % kU)];
        for (int idx = tid0; idx < kHid * kU; idx += stride) {
            const int o = idx / kU, u = idx % kU;
            out[bwd_img_w2c(N) + idx] = w2[o * kHid + c * kU + u];
            out[bwd_img_w1c(N) + idx] = w1[o * kHid + c * kU + u];
        }
        for (int idx = tid0; idx < 3 * kHid * kU; idx += stride)
            out[bwd_img_whhc(N) + idx] = w_hh[(idx / kU) * kHid + c * kU + (idx % kU)];
    } else if (bx < 4 * kCS) {
        if (!(want & 2)) return;
        const int c = bx - 3 * kCS;
        float* out = ws + single_off_bstr(N) + (long long)c * 2 * 3 * kHid * kU;
        for (int idx = tid0; idx < 3 * kHid * kU; idx += stride) {
            const int o = idx / kU, u = idx % kU, n = c * NU + u;
            const bool own = u < NU && n < N;
            out[idx] = own ? w_ih[(long long)o * p.Kin + n] : 0.f;
            out[3 * kHid * kU + idx] = own ? w_ih[(long long)o * p.Kin + 2 * N + n] : 0.f;
        }
    } else if (want & 1) {
        const int tiles = (p.B + kR - 1) / kR;
        float4* h0 = reinterpret_cast<float4*>(p.H);
        for (long long i = tid0; i < (long long)tiles * kHid * kR / 4; i += stride) h0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid0; i < (p.T - 1) + 1; i += stride) p.flags[i] = 0;
    }
}

// ---- small device helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sum8(float v) {   // over the 8 adjacent lanes of a k-split group; every lane gets the sum
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// acc[g][i] += sum_{k = ks, ks+8, .. < K} x[k][i] * w[(k * NG + g) * 32]      (x: this row group's [K][4]; w offset to the
// thread's skewed column)
template <int NG>
__device__ __forceinline__ void dot8(float2 (&lo)[NG], float2 (&hi)[NG], const float* __restrict__ x, const float* __restrict__ w,
                                     int K, int ks) {
#pragma unroll 4
    for (int k = ks; k < K; k += kKL) {
        const float4 xv = *reinterpret_cast<const float4*>(x + k * 4);
        const float2 xl = make_float2(xv.x, xv.y), xh = make_float2(xv.z, xv.w);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float wv = w[(k * NG + g) * kU];
            const float2 w2 = make_float2(wv, wv);
            lo[g] = __ffma2_rn(w2, xl, lo[g]);
            hi[g] = __ffma2_rn(w2, xh, hi[g]);
        }
    }
}

__device__ __forceinline__ void bar_all() { __syncthreads(); }

// LayerNorm + SiLU + Dropout over the 8 full rows in buf_s ([rg][feature][4], in place), redundantly in every CTA; the
// warps whose features are the CTA's own slice save them (tile layout, rows row_off .. row_off + 7 of the 16-row tile).
// thread = (rg, fh = 16-feature block, fl, i): features fh*16 + 2 fl, +1 of clip rg*4 + i
__device__ __forceinline__ void ln_silu_drop(const BiearSeqParams& p, unsigned long long seed, float* buf_s, float* stat_s,
                                             const float* __restrict__ gamma_s, const float* __restrict__ beta_s, int layer,
                                             int t, long long grow0, int rank, float* xh_tile, float* d_tile,
                                             float* rstd_tile, int row_off) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp >> 3, fh = warp & 7, fl = lane >> 2, i = lane & 3;
    const int f0 = fh * 16 + fl * 2;
    float* row = buf_s + rg * kHid * kRT + i;
    const float pivot = row[0];
    float v0 = row[f0 * kRT] - pivot, v1 = row[(f0 + 1) * kRT] - pivot;
    float s = v0 + v1, ss = fmaf(v1, v1, v0 * v0);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    float2* stat2 = reinterpret_cast<float2*>(stat_s);               // [rg][i][fh]
    if (fl == 0) stat2[(rg * kRT + i) * 8 + fh] = make_float2(s, ss);
    __syncthreads();
    float S = 0.f, SS = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const float2 q = stat2[(rg * kRT + i) * 8 + w];
        S += q.x;
        SS += q.y;
    }
    const float mean = S * (1.0f / kHid);
    const float var = fmaxf(fmaf(-mean, mean, SS * (1.0f / kHid)), 0.0f);
    const float rstd = rsqrtf(var + kLnEps);
    v0 -= mean;
    v1 -= mean;
    const bool mine = (fh >> 1) == rank;
    const int r8 = rg * kRT + i;
    if (rank == 0 && fh == 0 && fl == 0) rstd_tile[layer * kR + row_off + r8] = rstd;
    float sc0 = 1.f, sc1 = 1.f;
    if (p.training) {
        const float4 sc = dropout_scale4(seed, t, layer, grow0 + r8, f0 >> 2);
        sc0 = (f0 & 2) ? sc.z : sc.x;
        sc1 = (f0 & 2) ? sc.w : sc.y;
    }
    const float xh0 = v0 * rstd, xh1 = v1 * rstd;
    const float y0 = fmaf(xh0, gamma_s[f0], beta_s[f0]), y1 = fmaf(xh1, gamma_s[f0 + 1], beta_s[f0 + 1]);
    const float o0 = (y0 * sigmoid_fast(y0)) * sc0, o1 = (y1 * sigmoid_fast(y1)) * sc1;
    row[f0 * kRT] = o0;
    row[(f0 + 1) * kRT] = o1;
    if (mine) {
        xh_tile[f0 * kR + row_off + r8] = xh0;
        xh_tile[(f0 + 1) * kR + row_off + r8] = xh1;
        d_tile[f0 * kR + row_off + r8] = o0;
        d_tile[(f0 + 1) * kR + row_off + r8] = o1;
    }
    __syncthreads();
}

template <bool STRICT>
__global__ void __launch_bounds__(kSeqThreads, 1) seq1_fwd_kernel(const BiearSeqParams p, const float* __restrict__ ws) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int N = p.N, T = p.T, S = p.T - 1, B = p.B;
    const int tiles = (B + kR - 1) / kR;
    const int n_half = 2 * tiles;                        // 8-clip halves of the 16-row tiles
    const int NU = bands_per_cta(N);
    const Smem L(N, p.F);
    const int Kp = L.Kp, CK = L.CK, n_chunks = Kp / CK;
    float* img_s = smem + L.img();
    float* vec_s = smem + L.vec();
    float* in_s = smem + L.in();
    float* hbuf_s = smem + L.h();
    float* a1_s = smem + L.a1();
    float* a2_s = smem + L.a2();
    float* stat_s = smem + L.stat();
    float* q_s = smem + L.q();
    float* ystage_s = smem + L.ystage();
    float4* spec_s = reinterpret_cast<float4*>(smem + L.spec());
    float* ring_s = smem + L.ring();
    const float* whh_s = img_s;
    const float* w1_s = img_s + kHid * 3 * kU;
    const float* w2_s = w1_s + kHid * kU;
    const float* w3_s = w2_s + kHid * kU;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // GEMM-phase coordinates
    const int ks = lane & 7, rg = warp >> 3, u = (warp & 7) * 4 + (lane >> 3);
    const int ucol = (u + 4 * ks) & 31;
    const int ug = rank * kU + u;
    const int nu_c = max(0, min(NU, N - rank * NU));
    const int quads = (N + 3) >> 2;
    int* any_flag = p.flags + S;
    const unsigned long long seed = p.seed_ptr ? *p.seed_ptr : p.seed;
    if (STRICT) {
        if (!p.force_strict && *reinterpret_cast<volatile int*>(any_flag) == 0) return;   // uniform over the grid
    }

    const uint32_t bars = smem_u32(smem + L.bars());
    auto bar_of = [&](int i) { return bars + 8u * (uint32_t)i; };          // 0..4: c, h, a1, a2, Q
    auto full_of = [&](int s) { return bars + 8u * (uint32_t)(5 + s); };
    auto empty_of = [&](int s) { return bars + 8u * (uint32_t)(5 + kStages + s); };
    const uint32_t tx_bytes[5] = {(uint32_t)(2 * N * kRows * 4), (uint32_t)(kHid * kRows * 4), (uint32_t)(kHid * kRows * 4),
                                  (uint32_t)(kHid * kRows * 4), (uint32_t)(N * kClipsB * 4)};
    auto arm = [&](int i) {
        if (tid == 0) mbar_arrive_expect_tx(bar_of(i), tx_bytes[i]);
    };
    const uint32_t chunk_bytes = (uint32_t)(CK * 3 * kU * 4);
    const float* wih_g = ws + single_off_fstr(N) + (long long)rank * Kp * 3 * kU;
    auto issue_chunk = [&](unsigned gc) {                 // thread 0 only: chunk number gc (running over frames) -> its slot
        const int s = (int)(gc % kStages);
        mbar_arrive_expect_tx(full_of(s), chunk_bytes);
        bulk_g2s(smem_u32(ring_s + s * CK * 3 * kU), wih_g + (long long)(gc % (unsigned)n_chunks) * CK * 3 * kU, chunk_bytes,
                 full_of(s));
    };

    // ---- set-up ------------------------------------------------------------------------------------------------------
    copy_f4(reinterpret_cast<float4*>(img_s),
            reinterpret_cast<const float4*>(ws + single_off_fres() + (long long)rank * single_fwd_res_floats()),
            single_fwd_res_floats() / 4);
    for (int i = tid; i < kHid; i += kSeqThreads) {
        vec_s[V_LN1G + i] = p.ln1_g[0][i];
        vec_s[V_LN1B + i] = p.ln1_b[0][i];
        vec_s[V_LN2G + i] = p.ln2_g[0][i];
        vec_s[V_LN2B + i] = p.ln2_b[0][i];
        vec_s[V_FC + i] = i < N ? p.fc[i] : 1.0f;
        vec_s[V_Q0 + i] = i < N ? p.q0[i] : 1.0f;
    }
    if (tid < kU) {
        const float* b_ih = p.b_ih[0];
        const float* b_hh = p.b_hh[0];
        const int o = rank * kU + tid;
        vec_s[V_BR + tid] = b_ih[o] + b_hh[o];
        vec_s[V_BZ + tid] = b_ih[kHid + o] + b_hh[kHid + o];
        vec_s[V_BIN + tid] = b_ih[2 * kHid + o];
        vec_s[V_BHN + tid] = b_hh[2 * kHid + o];
        vec_s[V_B1 + tid] = p.b1[0][o];
        vec_s[V_B2 + tid] = p.b2[0][o];
        const int n = rank * NU + tid;
        const bool own = tid < NU && n < N;
        vec_s[V_B3 + tid] = own ? p.b3[0][n] : 0.f;
        vec_s[V_Q0S + tid] = own ? p.q0[n] : 1.f;
        vec_s[V_DQS + tid] = own ? p.dq[n] : 0.f;
    }
    for (int i = tid; i < Kp * kRows; i += kSeqThreads) in_s[i] = 0.f;     // memory = 0, padding rows = 0
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) mbar_init(bar_of(i), 1);
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_of(s), 1);
            mbar_init(empty_of(s), kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        fence_proxy_async();
        for (unsigned gc = 0; gc < (unsigned)kStages; ++gc) issue_chunk(gc);
    }
    cluster.sync();                                           // every CTA's barriers exist before anybody signals them
    if (STRICT) {                                             // the replay decides the flags anew
        for (int i = rank * kSeqThreads + tid; i < S; i += kCS * kSeqThreads) p.flags[i] = 0;
        __threadfence();
        cluster.sync();
    }
    unsigned gchunk = 0;                                      // W_ih chunks consumed so far
    uint32_t iter = 0;                                        // controller iterations completed (hand-over parity)
    int hsel = 0;
    int spec_key = -1;                                        // (frame * n_half + half) whose spectra sit in spec_s

    PHASE1_INIT();
    const int half_begin = STRICT ? 0 : (int)(blockIdx.x / kCS), half_end = STRICT ? n_half : half_begin + 1;
    for (int t = 0; t < T; ++t) {
        for (int hf = half_begin; hf < half_end; ++hf) {
            const int tile = hf >> 1, row_off = (hf & 1) * kRows;
            const int b0 = hf * kRows;                        // first clip of the cluster's 8
            const int bb0 = b0 + rank * kClipsB;              // first clip of this CTA's band stage
            auto item_row = [&](int item) { return (long long)(item >> 1) * B + bb0 + (item & 1); };   // row of (E*B, ...) tensors
            auto prefetch = [&](int tf) {
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    const float2* src = reinterpret_cast<const float2*>(p.X) + (item_row(i) * T + tf) * p.F;
                    const bool row_ok = bb0 + (i & 1) < B;
                    for (int k = tid; k < L.tile; k += kSeqThreads) {
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
                for (int i = 0; i < kItems; ++i) {
                    if (bb0 + (i & 1) >= B) continue;
                    for (int k = tid; k < p.F; k += kSeqThreads) {
                        float4* slot4 = spec_s + i * L.tile + k;
                        *slot4 = spec_entry(make_float2(slot4->z, slot4->w));
                    }
                }
            };

            PHASE1_MARK(0);   // loop tail / head
            // ---- state of this (frame, half-tile) ----------------------------------------------------------------
            bool h_zero = t == 0;
            float* hcur_s = hbuf_s + hsel * kHid * kRows;
            float* hnext_s = hbuf_s + (hsel ^ 1) * kHid * kRows;
            if (!STRICT) {
                if (t == 0) {
                    for (int idx = tid; idx < kClipsB * N; idx += kSeqThreads) {
                        const int j = idx / N, n = idx - j * N;
                        q_s[n * kClipsB + j] = vec_s[V_Q0 + n];
                        if (bb0 + j < B) p.Q[((long long)(bb0 + j) * T + t) * N + n] = vec_s[V_Q0 + n];
                    }
                }
            } else {
                const bool fallback_prev = t > 0 && __ldcg(p.flags + (t - 1)) != 0;
                const bool use_q0 = t == 0 || fallback_prev;
                h_zero = use_q0;
                for (int idx = tid; idx < kClipsB * N; idx += kSeqThreads) {
                    const int j = idx / N, n = idx - j * N;
                    const long long e = ((long long)(bb0 + j) * T + t) * N + n;
                    float qv = vec_s[V_Q0 + n];
                    if (bb0 + j < B) {
                        if (use_q0) p.Q[e] = qv;
                        else qv = __ldcg(p.Q + e);
                    }
                    q_s[n * kClipsB + j] = qv;
                }
                // h_{t-1} and the carried memory come back from what the previous frame saved
                const float* hsrc = p.H + (((long long)t) * tiles + tile) * (kHid * kR);      // step index t = h_{t-1}
                for (int idx = tid; idx < kHid * kRG; idx += kSeqThreads) {
                    const int f = idx >> 1, g2 = idx & 1;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!h_zero) v = __ldcg(reinterpret_cast<const float4*>(hsrc + f * kR + row_off + g2 * kRT));
                    *reinterpret_cast<float4*>(hcur_s + (g2 * kHid + f) * kRT) = v;
                }
                if (fallback_prev && rank == 0) {   // h_{t-1} was dropped: it must not feed dW_hh either
                    float* hdst = p.H + (((long long)t) * tiles + tile) * (kHid * kR);
                    for (int idx = tid; idx < kHid * kRG; idx += kSeqThreads)
                        *reinterpret_cast<float4*>(hdst + (idx >> 1) * kR + row_off + (idx & 1) * kRT) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const float* insrc = t > 0 ? p.yc + (((long long)(t - 1)) * tiles + tile) * (4 * N * kR) : nullptr;
                for (int idx = tid; idx < 2 * N * kRG; idx += kSeqThreads) {
                    const int k2 = idx >> 1, g2 = idx & 1;              // k2 < 2N: L bands then R bands
                    const int ear = k2 / N, n = k2 - ear * N;
                    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t > 0) {
                        const float4 c = __ldcg(reinterpret_cast<const float4*>(insrc + (ear * 2 * N + n) * kR + row_off + g2 * kRT));
                        const float4 mo = __ldcg(reinterpret_cast<const float4*>(insrc + (ear * 2 * N + N + n) * kR + row_off + g2 * kRT));
                        m = make_float4(__fadd_rn(__fmul_rn(0.8f, mo.x), __fmul_rn(0.2f, c.x)), __fadd_rn(__fmul_rn(0.8f, mo.y), __fmul_rn(0.2f, c.y)),
                                        __fadd_rn(__fmul_rn(0.8f, mo.z), __fmul_rn(0.2f, c.z)), __fadd_rn(__fmul_rn(0.8f, mo.w), __fmul_rn(0.2f, c.w)));
                    }
                    *reinterpret_cast<float4*>(in_s + (g2 * Kp + k2) * kRT) = m;
                }
            }
            const int want_key = t * n_half + hf;
            if (spec_key != want_key) {     // first frame / strict pass: fetch and convert now (also orders the q_s fill)
                __syncthreads();
                prefetch(t);
                finish();
                spec_key = want_key;
            }
            __syncthreads();
            PHASE1_MARK(1);   // state + spectra ready

            // ---- band stage of frame t for this CTA's 4 items (model_torch.py:729-737, 1050-1060) ---------------------
            // The kItems x quads (item, 4-band quad) pairs are dealt to the 16 warps widest quads first in snake order (round
            // m goes warp 0..15 for even m, 15..0 for odd m: evens out the bins per warp), items rotating; lane l OWNS band
            // (l & 3) of the warp's (l >> 2)-th pair (parameters once, epilogue once).
            bool own_store = false;
            long long own_e = 0;
            float oY = 0.f, oJ = 0.f, oP = 0.f, oK = 0.f;
            {
                const int n_pairs = kItems * quads;
                const int m_own = lane >> 2;
                const int p_own = kWarps * m_own + ((m_own & 1) ? kWarps - 1 - warp : warp);
                const int item_own = ((p_own & 3) + (p_own >> 4)) & 3;
                const int n_own = ((quads - 1 - (p_own >> 2)) << 2) + (lane & 3);
                const bool own = p_own < n_pairs && n_own < N;
                const float fc = own ? vec_s[V_FC + n_own] : 1.0f;
                const float q = own ? q_s[n_own * kClipsB + (item_own & 1)] : 1.0f;
                const BandParams bp_own = band_params(fc, q, p.df, p.cutoff, p.F, own);
                BandSums keep = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int m = 0; m * kWarps < n_pairs; ++m) {
                    const int pr = kWarps * m + ((m & 1) ? kWarps - 1 - warp : warp);
                    if (pr >= n_pairs) break;                        // (warp-uniform)
                    const int item = ((pr & 3) + (pr >> 4)) & 3;
                    const int src = (m << 2) + (lane >> 3);          // lane owning the band this lane helps with
                    BandParams bp;
                    bp.bw = 0.f;
                    bp.a = __shfl_sync(0xffffffffu, bp_own.a, src);
                    bp.b = __shfl_sync(0xffffffffu, bp_own.b, src);
                    bp.kc = __shfl_sync(0xffffffffu, bp_own.kc, src);
                    bp.k_lo = __shfl_sync(0xffffffffu, bp_own.k_lo, src);
                    bp.k_hi = __shfl_sync(0xffffffffu, bp_own.k_hi, src);
                    const BandSums sums = band_accumulate(spec_s + item * L.tile, p.F, bp, lane);
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
                    ystage_s[item_own * kHid + n_own] = log1pf(fmaxf(r.Y, 0.0f));
                    own_store = bb0 + (item_own & 1) < B;
                    own_e = (item_row(item_own) * T + t) * N + n_own;
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
            if (own_store) {
                p.Y[own_e] = oY;
                p.dYdQ[own_e] = oJ;
                if (p.logY) p.logY[own_e] = fminf(fmaxf(logf(oY + 1e-8f), -12.0f), 12.0f);
                if (p.phase) {
                    p.phase[own_e] = oP;
                    p.dPdQ[own_e] = oK;
                }
            }
            if (t == T - 1) {
                // The reference runs the controller once more and discards the result (model_torch.py:750-771).
                __syncthreads();
                continue;
            }
            __syncthreads();
            PHASE1_MARK(2);   // band stage
            if (!STRICT) {           // the tile is free again: the next frame's spectra travel behind the controller phases
                prefetch(t + 1);
                spec_key = (t + 1) * n_half + hf;
            }
            const long long tb = (long long)t * tiles + tile;
            const uint32_t par = iter & 1u;
            if (tid < 2 * N) {   // current features of my 2 clips, both ears -> every CTA of the cluster
                const int ear = tid / N, n = tid - ear * N;
                const float2 v = make_float2(ystage_s[(ear * 2) * kHid + n], ystage_s[(ear * 2 + 1) * kHid + n]);
                const uint32_t a = smem_u32(in_s + ((rank >> 1) * Kp + 2 * N + ear * N + n) * kRT + (rank & 1) * kClipsB);
#pragma unroll
                for (uint32_t dst = 0; dst < (uint32_t)kCS; ++dst)
                    st_async_f2(cluster_addr(a, dst), v, cluster_addr(bar_of(0), dst));
            }
            arm(0);
            // ---- GRU cell, part A: everything that does not depend on this frame's band stage -- the memory half of the
            // input product (the first n_pre chunks of the W_ih ring) and the recurrent product -- runs while hand-over #1
            // (the current features of the cluster's 8 clips) is in flight.
            float2 lo[4], hi[4];       // r, z, i_n, h_n
#pragma unroll
            for (int g = 0; g < 4; ++g) lo[g] = hi[g] = make_float2(0.f, 0.f);
            const float* x_rg = in_s + rg * Kp * kRT;
            auto ring_chunks = [&](int c_begin, int c_end) {
                const int kper = CK / kKL;
                for (int c = c_begin; c < c_end; ++c, ++gchunk) {
                    const int s = (int)(gchunk % kStages);
                    if (tid == 0 && gchunk >= 1) {   // refill the slot of the previous chunk (everybody has long left it)
                        const unsigned gp = gchunk - 1;
                        mbar_wait(empty_of((int)(gp % kStages)), (gp / kStages) & 1u);
                        issue_chunk(gp + kStages);
                    }
                    mbar_wait(full_of(s), (gchunk / kStages) & 1u);
                    const float* wst = ring_s + s * CK * 3 * kU + ucol;
                    const float* xc = x_rg + c * CK * kRT;
#pragma unroll 5
                    for (int j = 0; j < kper; ++j) {
                        const int kk = j * kKL + ks;
                        const float4 xv = *reinterpret_cast<const float4*>(xc + kk * kRT);
                        const float2 xl = make_float2(xv.x, xv.y), xh = make_float2(xv.z, xv.w);
                        const float w0 = wst[(kk * 3) * kU], w1v = wst[(kk * 3 + 1) * kU], w2v = wst[(kk * 3 + 2) * kU];
                        const float2 p0 = make_float2(w0, w0), p1 = make_float2(w1v, w1v), p2 = make_float2(w2v, w2v);
                        lo[0] = __ffma2_rn(p0, xl, lo[0]); hi[0] = __ffma2_rn(p0, xh, hi[0]);
                        lo[1] = __ffma2_rn(p1, xl, lo[1]); hi[1] = __ffma2_rn(p1, xh, hi[1]);
                        lo[2] = __ffma2_rn(p2, xl, lo[2]); hi[2] = __ffma2_rn(p2, xh, hi[2]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_of(s));
                }
            };
            const int n_pre = (2 * N) % CK == 0 ? (2 * N) / CK : 0;     // chunks that hold memory inputs only
            ring_chunks(0, n_pre);
            if (!h_zero) {
                const float* hx = hcur_s + rg * kHid * kRT;
                const float* w = whh_s + ucol;
#pragma unroll 4
                for (int k = ks; k < kHid; k += kKL) {
                    const float4 xv = *reinterpret_cast<const float4*>(hx + k * kRT);
                    const float2 xl = make_float2(xv.x, xv.y), xh = make_float2(xv.z, xv.w);
                    const float w0 = w[(k * 3) * kU], w1v = w[(k * 3 + 1) * kU], w2v = w[(k * 3 + 2) * kU];
                    const float2 p0 = make_float2(w0, w0), p1 = make_float2(w1v, w1v), p2 = make_float2(w2v, w2v);
                    lo[0] = __ffma2_rn(p0, xl, lo[0]); hi[0] = __ffma2_rn(p0, xh, hi[0]);
                    lo[1] = __ffma2_rn(p1, xl, lo[1]); hi[1] = __ffma2_rn(p1, xh, hi[1]);
                    lo[3] = __ffma2_rn(p2, xl, lo[3]); hi[3] = __ffma2_rn(p2, xh, hi[3]);
                }
            }
            tx_wait(bar_of(0), par);
            PHASE1_MARK(3);   // prefetch issue + push + GRU part A + hand-over #1
            // saved controller input (torch column order cL | mL | cR | mR, tile layout): ranks 0 / 1 save cL / cR here,
            // ranks 2 / 3 save mL / mR where the memory is advanced
            float* in_tile = p.yc + tb * (4 * N * kR);
            if (rank < 2 && tid < N * kRG) {
                const int n = tid >> 1, g2 = tid & 1;
                const float4 v = *reinterpret_cast<const float4*>(in_s + (g2 * Kp + 2 * N + rank * N + n) * kRT);
                *reinterpret_cast<float4*>(in_tile + (rank * 2 * N + n) * kR + row_off + g2 * kRT) = v;
            }

            // ---- GRU cell, part B: the current-feature half of the input product, then the cell (torch gate order r, z, n;
            // n = tanh(i_n + r * (W_hn h + b_hn))) ------------------------------------------------------------------------
            {
                ring_chunks(n_pre, n_chunks);
                float acc[4][kRT];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    acc[g][0] = sum8(lo[g].x); acc[g][1] = sum8(lo[g].y);
                    acc[g][2] = sum8(hi[g].x); acc[g][3] = sum8(hi[g].y);
                }
                // every lane of the k-split group holds the sums: all compute the cell, lanes 0-3 deliver to CTA 0-3,
                // lanes 4-7 write the saved state
                float hv[kRT], vr[kRT], vz[kRT], vn[kRT], vh[kRT];
                const float br = vec_s[V_BR + u], bz = vec_s[V_BZ + u], bin = vec_s[V_BIN + u], bhn = vec_s[V_BHN + u];
                const float4 hp4 = h_zero ? make_float4(0.f, 0.f, 0.f, 0.f)
                                          : *reinterpret_cast<const float4*>(hcur_s + (rg * kHid + ug) * kRT);
                const float hp[kRT] = {hp4.x, hp4.y, hp4.z, hp4.w};
#pragma unroll
                for (int i = 0; i < kRT; ++i) {
                    vr[i] = sigmoid_fast(acc[0][i] + br);
                    vz[i] = sigmoid_fast(acc[1][i] + bz);
                    vh[i] = acc[3][i] + bhn;
                    vn[i] = tanhf(acc[2][i] + bin + vr[i] * vh[i]);
                    hv[i] = (1.0f - vz[i]) * vn[i] + vz[i] * hp[i];
                }
                const float4 h4 = make_float4(hv[0], hv[1], hv[2], hv[3]);
                if (ks < kCS) {
                    const uint32_t a = smem_u32(hnext_s + (rg * kHid + ug) * kRT);
                    st_async_f4(cluster_addr(a, (uint32_t)ks), h4, cluster_addr(bar_of(1), (uint32_t)ks));
                } else {
                    const long long off = (long long)ug * kR + row_off + rg * kRT;
                    float* gt = p.gates + tb * 4 * kHid * kR + off;
                    if (ks == 4) {
                        store4(p.H + (((long long)(t + 1)) * tiles + tile) * (kHid * kR) + off, hv);
                        store4(gt + 3 * kHid * kR, vh);
                    } else if (ks == 5) {
                        store4(gt, vr);
                    } else if (ks == 6) {
                        store4(gt + kHid * kR, vz);
                    } else {
                        store4(gt + 2 * kHid * kR, vn);
                    }
                }
                arm(1);
            }
            tx_wait(bar_of(1), par);
            PHASE1_MARK(4);   // GRU + #2
            // ---- carried memory: Y_mem <- 0.8 Y_mem + 0.2 Y_ctrl.detach() (model_torch.py:770-771); every warp of every
            // CTA has finished its GRU products (its h values are part of hand-over 1), so the buffer may change now.
            // Ranks 2 / 3 save the OLD memory (this step's controller input) on the way.
            for (int idx = tid; idx < 2 * N * kRG; idx += kSeqThreads) {
                const int k2 = idx >> 1, g2 = idx & 1;
                float4* mp = reinterpret_cast<float4*>(in_s + (g2 * Kp + k2) * kRT);
                const float4 c = *reinterpret_cast<const float4*>(in_s + (g2 * Kp + 2 * N + k2) * kRT);
                const float4 m = *mp;
                const int ear = k2 >= N ? 1 : 0;
                if (rank == 2 + ear)
                    *reinterpret_cast<float4*>(in_tile + (ear * 2 * N + N + (k2 - ear * N)) * kR + row_off + g2 * kRT) = m;
                if (!STRICT)
                    *mp = make_float4(__fadd_rn(__fmul_rn(0.8f, m.x), __fmul_rn(0.2f, c.x)), __fadd_rn(__fmul_rn(0.8f, m.y), __fmul_rn(0.2f, c.y)),
                                      __fadd_rn(__fmul_rn(0.8f, m.z), __fmul_rn(0.2f, c.z)), __fadd_rn(__fmul_rn(0.8f, m.w), __fmul_rn(0.2f, c.w)));
            }

            // ---- Linear 1 -> LayerNorm -> SiLU -> Dropout ------------------------------------------------------------------
            {
                float2 lo[1] = {make_float2(0.f, 0.f)}, hi[1] = {make_float2(0.f, 0.f)};
                dot8<1>(lo, hi, hnext_s + rg * kHid * kRT, w1_s + ucol, kHid, ks);
                const float bb = vec_s[V_B1 + u];
                const float4 v = make_float4(sum8(lo[0].x) + bb, sum8(lo[0].y) + bb, sum8(hi[0].x) + bb, sum8(hi[0].y) + bb);
                if (ks < kCS) {
                    const uint32_t a = smem_u32(a1_s + (rg * kHid + ug) * kRT);
                    st_async_f4(cluster_addr(a, (uint32_t)ks), v, cluster_addr(bar_of(2), (uint32_t)ks));
                }
                arm(2);
            }
            tx_wait(bar_of(2), par);
            PHASE1_MARK(5);   // memory update + Linear 1 + #3
            ln_silu_drop(p, seed, a1_s, stat_s, vec_s + V_LN1G, vec_s + V_LN1B, 0, t, (long long)b0, rank,
                         p.xh1 + tb * kHid * kR, p.d1 + tb * kHid * kR, p.rstd + tb * 2 * kR, row_off);
            PHASE1_MARK(6);   // LayerNorm 1
            // ---- Linear 2 -> LayerNorm -> SiLU -> Dropout ------------------------------------------------------------------
            {
                float2 lo[1] = {make_float2(0.f, 0.f)}, hi[1] = {make_float2(0.f, 0.f)};
                dot8<1>(lo, hi, a1_s + rg * kHid * kRT, w2_s + ucol, kHid, ks);
                const float bb = vec_s[V_B2 + u];
                const float4 v = make_float4(sum8(lo[0].x) + bb, sum8(lo[0].y) + bb, sum8(hi[0].x) + bb, sum8(hi[0].y) + bb);
                if (ks < kCS) {
                    const uint32_t a = smem_u32(a2_s + (rg * kHid + ug) * kRT);
                    st_async_f4(cluster_addr(a, (uint32_t)ks), v, cluster_addr(bar_of(3), (uint32_t)ks));
                }
                arm(3);
            }
            tx_wait(bar_of(3), par);
            PHASE1_MARK(7);   // Linear 2 + #4
            ln_silu_drop(p, seed, a2_s, stat_s, vec_s + V_LN2G, vec_s + V_LN2B, 1, t, (long long)b0, rank,
                         p.xh2 + tb * kHid * kR, p.d2 + tb * kHid * kR, p.rstd + tb * 2 * kR, row_off);
            PHASE1_MARK(8);   // LayerNorm 2
            // ---- Linear 3 -> tanh -> Q_{t+1} (model_torch.py:757-768) ---------------------------------------------------------
            {
                float2 lo[1] = {make_float2(0.f, 0.f)}, hi[1] = {make_float2(0.f, 0.f)};
                const bool mine = u < nu_c;                  // (uniform over the 8 lanes of a k-split group)
                if (mine) dot8<1>(lo, hi, a2_s + rg * kHid * kRT, w3_s + ucol, kHid, ks);
                const float acc[kRT] = {sum8(lo[0].x), sum8(lo[0].y), sum8(hi[0].x), sum8(hi[0].y)};
                const int n = rank * NU + u;
                if (mine) {
                    const float bb = vec_s[V_B3 + u], q0 = vec_s[V_Q0S + u], dq = vec_s[V_DQS + u];
                    float qv[kRT], dv[kRT];
                    bool bad = false;
#pragma unroll
                    for (int i = 0; i < kRT; ++i) {
                        dv[i] = tanhf(acc[i] + bb);
                        const float qu = p.relative ? q0 * (1.0f + dq * dv[i]) : fmaf(dq, dv[i], q0);
                        qv[i] = fminf(fmaxf(qu, p.q_min), p.q_max);
                        bad = bad || (b0 + rg * kRT + i < B && !finite_f(qu));
                    }
                    // Q_{t+1} of clips 4 rg + {0,1} goes to CTA 2 rg, of clips 4 rg + {2,3} to CTA 2 rg + 1 (their band stage)
                    const uint32_t a = smem_u32(q_s + n * kClipsB);
                    if (ks == 0)
                        st_async_f2(cluster_addr(a, 2u * rg), make_float2(qv[0], qv[1]), cluster_addr(bar_of(4), 2u * rg));
                    else if (ks == 1)
                        st_async_f2(cluster_addr(a, 2u * rg + 1u), make_float2(qv[2], qv[3]), cluster_addr(bar_of(4), 2u * rg + 1u));
                    else if (ks == 2 && bad) {   // NaN / Inf: the reference falls back for the whole batch -> strict replay pass
                        atomicOr(p.flags + t, 1);
                        atomicOr(any_flag, 1);
                    } else if (ks >= 4) {
                        const int i = ks - 4, b = b0 + rg * kRT + i;
                        if (b < B) {
                            const long long e = ((long long)b * T + (t + 1)) * N + n;
                            p.Q[e] = qv[i];
                            p.delta[e] = dv[i];
                        }
                    }
                }
                if (!STRICT) finish();       // the next frame's spectrum tile (its cp.async data has long arrived)
                arm(4);
            }
            tx_wait(bar_of(4), par);   // Q_{t+1} of my band-stage clips has landed ...
            PHASE1_MARK(9);   // Linear 3 + Q + #5
            ++iter;
            if (STRICT) {
                __threadfence();       // Q / H / saved inputs / flags are read back from global memory in the next frame
                cluster.sync();
            } else {
                __syncthreads();       // ... and every thread has converted its slots of the next spectrum tile
                hsel ^= 1;
            }
        }
    }
    if (tid == 0) {   // drain the W_ih chunks that were prefetched but never consumed: no bulk copy may outlive the CTA
        const unsigned issued_end = gchunk >= 1 ? gchunk - 1 + kStages : kStages;
        for (unsigned gc = gchunk; gc < issued_end; ++gc) mbar_wait(full_of((int)(gc % kStages)), (gc / kStages) & 1u);
    }
    __syncthreads();
    cluster.sync();   // no CTA leaves (and frees its shared memory) while a peer could still be sending to it
}

template <typename Kern>
static int launch(Kern kern, const char* name, int clusters, size_t smem, cudaStream_t st, const BiearSeqParams& p) {
    // the shared-memory opt-in is per device and sticky: remember the largest size set, so that steady-state launches (and
    // launches recorded into a CUDA graph) make no attribute call
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
    e = check_cuda(cudaLaunchKernelEx(&cfg, kern, p, (const float*)p.workspace), name);
    if (e) return e;
    count_launch();
    return 0;
}

}  // namespace sc

size_t single_fwd_smem_bytes(int N, int F) { return sizeof(float) * (size_t)sc::Smem(N, F).total(); }

int launch_single_prepare(const BiearSeqParams* p, int want, cudaStream_t st) {
    sc::prepare_single_kernel<<<dim3(4 * kCS + 1, 16), 256, 0, st>>>(*p, p->workspace, want);
    BIEAR_LAUNCH_CHECK("prepare_single_kernel");
    return 0;
}

int debug_phase_cycles_single(unsigned long long* out_host) {
#ifdef BIEAR_PHASE_PROF
    unsigned long long zero[16] = {};
    if (int e = check_cuda(cudaMemcpyFromSymbol(out_host, sc::g_phase_cycles1, sizeof(zero)), "cudaMemcpyFromSymbol")) return e;
    return check_cuda(cudaMemcpyToSymbol(sc::g_phase_cycles1, zero, sizeof(zero)), "cudaMemcpyToSymbol");
#else
    (void)out_host;
    return fail_invalid("biear_debug_phase_cycles_single: library built without BIEAR_PHASE_PROF");
#endif
}

int launch_single_fwd(const BiearSeqParams* p, cudaStream_t st) {
    const size_t smem = single_fwd_smem_bytes(p->N, p->F);
    BIEAR_REQUIRE(smem <= 227 * 1024, "biear_adaptive_fwd (single controller): N=%d F=%d needs %zu B of shared memory", p->N, p->F, smem);
    const int halves = 2 * ((p->B + kR - 1) / kR);
    if (!p->force_strict)
        if (int e = sc::launch(sc::seq1_fwd_kernel<false>, "seq1_fwd_kernel", halves, smem, st, *p)) return e;
    return sc::launch(sc::seq1_fwd_kernel<true>, "seq1_fwd_strict_kernel", 1, smem, st, *p);
}

}  // namespace biear
