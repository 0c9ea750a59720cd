/*
 * biear_b200.h -- C ABI of libbiear_b200.so: the B200 (sm_100a) implementation of BiEAR's
 * active-mode binaural front-end hot path.
 *
 * The reference (anonymous-speech-researcher/BiEAR) is pure Python/PyTorch and has no FFI of its
 * own; the "binding" a maintainer adds is a ctypes stub inside model_torch.py.
 * Each entry point below names the reference code it replaces (paths relative to the reference
 * root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller allocates and
 *     owns all buffers; contiguous row-major fp32 unless a stride argument says otherwise;
 *     complex values are interleaved (re, im) float pairs.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Functions only
 *     enqueue work; they never synchronise (the *_host variants do, once, at the end).
 *   - return value: 0 on success, otherwise a cudaError_t (> 0) or BIEAR_EINVAL (-1);
 *     biear_last_error() returns a thread-local description of the last failure.
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef BIEAR_B200_H
#define BIEAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BIEAR_ABI_VERSION 20
#define BIEAR_EINVAL (-1)

/* ABI version (== BIEAR_ABI_VERSION of the library that was built). */
int biear_abi_version(void);

/* Description of the last error raised on the calling thread ("" if none). */
const char* biear_last_error(void);

/* Number of kernels this library has launched since load / since the last reset (all threads). */
int64_t biear_launch_count(void);
void biear_reset_launch_count(void);

/* One-time per-device setup (uploads the FFT twiddle table).  Synchronous; call it once per device
 * before the first biear_stft_fwd and outside any CUDA-graph capture.  Idempotent. */
int biear_init(void);

/*
 * Framing + Hann window + zero-padded real FFT for every (row, frame).
 * Replaces model_torch.py:289-312 (_frame_1s) and :334-335 (frame * win_fn; torch.fft.rfft(n=n_fft)).
 *   wav     (rows, nsamp) with row stride wav_row_stride floats; only the first `fs` samples of a
 *           row are used, shorter rows are zero padded (reference: pad/truncate to fs).
 *   win_fn  (win) window samples (the module's hann_window buffer).
 *   X       (rows, T, n_fft/2+1, 2) out.
 * Supported: n_fft == 1024; any fs, T >= 1, win >= 1, hop >= 1.
 */
/*
 * 16-bit PCM waveform -> float32 in [-1, 1): out[i] = in[i] * scale (scale = 1/32768 is the reference harness's own
 * convention for int16-range input, train_biear.py:463-467).  Lets the host hand clips over in the audio wire format --
 * half the host->device bytes of float32 -- with the conversion on the device, next to the copy.
 *   in (n) int16, out (n) float32, both device pointers; any n >= 0.
 */
int biear_pcm16_to_f32(const int16_t* in, float* out, int64_t n, float scale, void* stream);

/* Both ears in ONE launch, frames in frame-major order (frame 0 of every row first): rows [0, rows_each) come from wavA,
 * [rows_each, 2 rows_each) from wavB; X (2 rows_each, T, n_fft/2+1, 2).  ready (nullable, 2 rows_each T + 4 int32, 8-byte
 * aligned, cleared by the caller / biear_adaptive_prepare; the trailing entries are the kernel's in-order work counter): entry [row][t] is set to 1, with release ordering, once X[row][t] is
 * complete -- the hand-over BiearSeqParams.x_ready describes. */
int biear_stft_fwd_pair(const float* wavA, const float* wavB, int64_t rows_each, int64_t nsamp, int64_t wav_row_stride,
                        const float* win_fn, int fs, int T, int win, int hop, int n_fft, float* X, int32_t* ready,
                        void* stream);
int biear_stft_fwd(const float* wav, int64_t rows, int64_t nsamp, int64_t wav_row_stride,
                   const float* win_fn, int fs, int T, int win, int hop, int n_fft,
                   float* X, void* stream);

/*
 * Gaussian band weighting + contraction for `items` independent (spectrum, Q-vector) pairs.
 * Replaces model_torch.py:340-346 (bw, W = exp(-0.5 u^2), row-normalise, nan_to_num,
 * einsum("bf,bnf->bn")) and, when phase != NULL, :1050-1060 (_subband_phase_from_X) in the same
 * pass; the (B,N,F) weight tensor is never materialised.  Item i reads spectrum X + i*x_stride
 * (F interleaved complex values; strides are in floats) and Q + i*q_stride (N values); a stride of 0
 * broadcasts (fixed-Q path).
 *   Y      (items, N)  band energies  nan_to_num(sum_f |X| W)              (required)
 *   phase  (items, N)  atan2(Im Z, Re Z), Z = sum_f W X                    (nullable)
 *   dYdQ   (items, N)  exact dY/dQ     = k (a2 - Y m2)                      (nullable)
 *   dPdQ   (items, N)  exact dphase/dQ = k (Re Z Im z2 - Im Z Re z2)/|Z|^2  (nullable)
 *          with m2 = sum W u^2, a2 = sum |X| W u^2, z2 = sum X W u^2, k = -fc/((Q+1e-8)^2 bw):
 *          what autograd derives in the reference, in closed form.
 *   cutoff: Gaussian support half-width in units of bw (|u| <= cutoff); <= 0 selects dense.
 */
int biear_band_fwd(const float* X, int64_t x_stride, const float* Q, int64_t q_stride,
                   const float* fc, int64_t items, int N, int F, float df, float cutoff,
                   float* Y, int64_t y_stride, float* phase, int64_t phase_stride,
                   float* dYdQ, float* dPdQ, int64_t jac_stride, void* stream);

/*
 * Fixed-Q band stage as one dense contraction: every item uses the SAME Q vector (N), so the Gaussian weights are a
 * single (N x F) matrix and Y = abs(X) W^T, Z = X W^T is a GEMM over all (row, frame) items at once.
 * Replaces model_torch.py:451-487 (FramewiseFixedGammatoneFB), :161-195 (AuralNetGammatoneFB) and, with phase != NULL,
 * :1039-1063 for fixed Q.  Same window truncation and normalisation as biear_band_fwd.
 *   X (items, F, 2) with item stride x_stride floats; Q, fc (N <= 128); Y / phase (items, N) with item strides;
 *   workspace: biear_band_fixed_workspace_floats(F) floats, 16-byte aligned (the packed weight matrix).
 */
int64_t biear_band_fixed_workspace_floats(int F);
int biear_band_fixed_fwd(const float* X, int64_t x_stride, const float* Q, const float* fc, int64_t items, int N, int F,
                         float df, float cutoff, float* Y, int64_t y_stride, float* phase, int64_t phase_stride,
                         float* workspace, void* stream);
/* The same contraction on the tcgen05 tensor cores: three 128 x 128 fp32 accumulators in TMEM, TF32 operands with the
 * 3-way hi/lo split (fp32-accurate: the 1e-4 contract holds), A tiles built on the fly from X, B tiles by bulk async
 * copy.  Same arguments; workspace: biear_band_fixed_tc_workspace_floats(F) floats, 16-byte aligned. */
int64_t biear_band_fixed_tc_workspace_floats(int F);
int biear_band_fixed_fwd_tc(const float* X, int64_t x_stride, const float* Q, const float* fc, int64_t items, int N, int F,
                            float df, float cutoff, float* Y, int64_t y_stride, float* phase, int64_t phase_stride,
                            float* workspace, void* stream);

/*
 * Backward of biear_band_fwd into Q by recomputation (nothing saved by the forward):
 *   dQ = gY * dY/dQ + gphase * dphase/dQ      per (item, band), written by one lane, no atomics.
 * This is the memory-lean form of what autograd derives from model_torch.py:340-346 / :1050-1060.
 *   gY, gphase (items, N) with item strides (either may be NULL, not both); dQ (items, N);
 *   accumulate != 0 adds into dQ instead of overwriting.
 */
int biear_band_bwd(const float* X, int64_t x_stride, const float* Q, int64_t q_stride,
                   const float* fc, int64_t items, int N, int F, float df, float cutoff,
                   const float* gY, int64_t gy_stride, const float* gphase, int64_t gphase_stride,
                   float* dQ, int64_t dq_stride, int accumulate, void* stream);

/*
 * Parameter block of the fused adaptive recurrence (dual front-end: one Q controller per ear).
 * Replaces the frame loop of FramewiseAdaptiveGammatoneFB.forward (model_torch.py:333-380) for both ears
 * and DeepEarActiveWaveform._subband_phase_from_X (:1039-1063), and their autograd backward.
 * Row order of the row-major tensors: ear-major, row = ear * B + clip.
 * "Tile layout" tensors are (G, T-1, tiles, D, R) with R = biear_adaptive_tile_rows() (16) and tiles = ceil(B / R):
 * R consecutive clips of one controller form a tile, stored feature-major ([D][R]); rows of the last tile beyond B are padding (their
 * gradient entries are exactly zero).  The forward writes them, the backward reads them and writes the
 * per-sample pre-activation gradients in the same layout; biear_ctrl_wgrad turns those into weight gradients.
 */
#define BIEAR_MAX_CTRL 2   /* controllers per call: one per ear in the dual front-end */
typedef struct BiearSeqParams {
    /* geometry */
    int32_t G, E, B, T, N, F, Kin;   /* controllers, ears, clips, frames, bands (<= 128), bins, controller input width */
    int32_t relative;                /* deltaQ_mode: 1 = Q0 (1 + dQ delta), 0 = Q0 + dQ delta */
    int32_t training;                /* dropout on (p = 0.1, Philox keyed by seed) */
    int32_t force_strict;            /* testing: skip the optimistic pass, run the batch-global (strict) pass only */
    uint64_t seed;
    float df, cutoff, q_min, q_max;
    /* constants (N) */
    const float *fc, *q0, *dq;
    /* controller weights: entry g < G of every array points at controller g's tensor in its torch layout (out, in),
       i.e. straight at the nn.GRU / nn.Linear / nn.LayerNorm parameters -- no stacking copy */
    const float *w_ih[BIEAR_MAX_CTRL], *w_hh[BIEAR_MAX_CTRL], *b_ih[BIEAR_MAX_CTRL], *b_hh[BIEAR_MAX_CTRL]; /* (384,Kin) (384,128) (384) (384) */
    const float *w1[BIEAR_MAX_CTRL], *b1[BIEAR_MAX_CTRL], *ln1_g[BIEAR_MAX_CTRL], *ln1_b[BIEAR_MAX_CTRL];   /* (128,128) (128) x3 */
    const float *w2[BIEAR_MAX_CTRL], *b2[BIEAR_MAX_CTRL], *ln2_g[BIEAR_MAX_CTRL], *ln2_b[BIEAR_MAX_CTRL];
    const float *w3[BIEAR_MAX_CTRL], *b3[BIEAR_MAX_CTRL];                                                   /* (N,128) (N) */
    /* spectra (E*B, T, F, 2) */
    const float* X;
    /* forward outputs, row-major (E*B, T, N); phase / dPdQ nullable together */
    float *Y, *phase, *dYdQ, *dPdQ;
    float *Q, *delta;                                /* (G*B, T, N): Q used for frame t; tanh output that produced it */
    /* saved by the forward for the backward, tile layout, D = 512 (r,z,n,hn), 128 x4, 2, N */
    float *gates, *xh1, *d1, *xh2, *d2, *rstd, *yc;
    /* GRU states (G, T, tiles, 128, R): step index 0 is h_{-1} = 0 (zeroed by the preparation launch), h_t is written
       at step index t+1 -- so H[:, :T-1] are the "previous states" and H[:, 1:] the "new states" of the T-1 steps */
    float* H;
    int32_t* flags;                                  /* ((T-1)*G + 1), cleared by the preparation launch: flags[t*G+g] != 0
                                                        <=> the non-finite-Q fallback (model_torch.py:378-380) was
                                                        taken after step t for controller g; last entry: any */
    /* backward inputs, ONE POINTER PER EAR / CONTROLLER g (each nullable, each a contiguous (B,T,N) tensor): dL/dY,
       dL/dphase, dL/dQ.  Separate pointers let a framework hand over the per-ear gradients as they are, without
       concatenating them first. */
    const float *gY[BIEAR_MAX_CTRL], *gP[BIEAR_MAX_CTRL], *gQ[BIEAR_MAX_CTRL];
    /* backward outputs, tile layout: dL/d[r_pre, z_pre, n_in_pre, hn] (D = 512); dL/d pre-LN and dL/d LN-output of
       layers 1, 2 (D = 128 each); dL/d(pre-tanh output) (D = N) */
    float *GG, *G_a1, *G_v1, *G_a2, *G_v2, *G_pre;
    /* scratch: biear_adaptive_workspace_floats(G, N) floats (packed per-CTA weight images, forward then backward);
       must stay untouched between the forward and the backward of one step */
    float* workspace;
    /* optional DEVICE location of the dropout seed; when non-NULL it overrides `seed` and is read by the kernels at
       run time, so that a captured CUDA graph draws fresh masks on every replay (the caller advances it on-stream) */
    const uint64_t* seed_ptr;
    /* optional fused log-energy features (model_torch.py:1080-1083): logY = clamp(log(Y + 1e-8), -12, 12), row-major
       (E*B, T, N), written by the forward band stage when non-NULL; gLogY[g] = dL/dlogY of ear g for the backward
       ((B,T,N), nullable): its contribution gLogY / (Y + 1e-8) (zero where the clamp is active) is added to dL/dY
       inside the kernel */
    float* logY;
    const float* gLogY[BIEAR_MAX_CTRL];
    /* non-zero: biear_adaptive_prepare has already run on this block's workspace / H / flags (stream-ordered before
       the forward), so biear_adaptive_fwd / _bwd skip their own preparation launch */
    int32_t prepared;
    /* optional streaming hand-over of the spectra: x_ready[(ear * B + clip) * T + t] != 0 <=> X[row][t] is complete
       (E*B*T + 4 int32: the last four are the STFT's work counter, cleared together with the flags).
       When non-NULL the preparation launch clears it, biear_stft_fwd_pair (frame-major order, running CONCURRENTLY on
       another stream, typically on the SMs the persistent kernel leaves idle) sets the entries, and the forward kernel
       waits for the entries of a frame before it reads that frame -- so the recurrence does not wait for the whole STFT.
       The caller must still order every later reader of X behind the STFT's completion. */
    int32_t* x_ready;
} BiearSeqParams;

/* 1 if the persistent recurrence kernels can take N bands and F bins (their weight slices, activations and spectrum
 * tiles must fit the 227 KB of shared memory of an SM), else 0: callers fall back to per-frame launches. */
int biear_adaptive_supported(int N, int F);
/* The SINGLE-controller front-end (BinauralAdaptiveGammatoneFB_SingleController, model_torch.py:579-776) goes through the
 * same entry points with G = 1, E = 2, Kin = 4N: one controller (GRU(4N -> 128) + MLP) whose input is
 * [log1p YL, memL, log1p YR, memR] and whose one Q per clip drives both ears' band stages.  Differences of the block:
 *   Q, delta (B,T,N); Y, phase, dYdQ, dPdQ, logY (2B,T,N) ear-major; gY / gP / gLogY indexed by EAR, gQ[0] only;
 *   `yc` saves the whole controller input, tile layout with D = 4N in weight_ih's column order; x_ready must be NULL;
 *   workspace = biear_single_workspace_floats(N) floats. */
int biear_single_supported(int N, int F);
int64_t biear_single_workspace_floats(int N);
/* Rows per tile (R) of the tile-layout tensors. */
int biear_adaptive_tile_rows(void);
/* Floats of scratch the calls below need in BiearSeqParams.workspace. */
int64_t biear_adaptive_workspace_floats(int G, int N);
/* Everything of a step that does not depend on the spectra, in ONE launch: the per-CTA weight images of the forward
 * and the backward kernel (from the controllers' nn.GRU / nn.Linear parameters), H[:, 0] = 0, flags = 0.  Needs only
 * the geometry, the weight pointers, workspace, H and flags of the block.  Optional: a caller that runs it on a forked
 * stream next to the STFT sets `prepared` = 1 afterwards; with `prepared` == 0 the forward / backward do the same work
 * themselves on their own stream. */
int biear_adaptive_prepare(const BiearSeqParams* p, void* stream);

/* Diagnostics: number of clusters of the forward / backward recurrence kernels that fit on the current device
 * at once (cudaOccupancyMaxActiveClusters); a batch needs G * ceil(B / R) clusters per pass. */
int biear_adaptive_occupancy(int N, int F, int* fwd_clusters, int* bwd_clusters);

/* Diagnostic builds only (-DBIEAR_PHASE_PROF, `make -C biear_b200/csrc prof`): per-phase clock64() totals of block 0 of
 * the forward [0][*] and backward [1][*] recurrence kernels, copied to out_host[2*16 + 8] (the last 8: inside the band stage) and cleared. */
int biear_debug_phase_cycles(unsigned long long* out_host);
/* ... and of the single-controller forward kernel (csrc/seq_single.cu): out_host[16]. */
int biear_debug_phase_cycles_single(unsigned long long* out_host);

/* Whole forward recurrence in ONE persistent cluster kernel (plus a weight-packing launch and a conditional
 * replay launch that exits immediately unless a non-finite Q was produced): each cluster of 4 CTAs carries R rows
 * through all T frames with the controller weights and the recurrent state resident in (distributed) shared
 * memory.  No host synchronisation.
 * Non-finite fallback: the reference resets Q to Q0 and the GRU state for the WHOLE batch of an ear when any
 * Q_{t+1} is non-finite (model_torch.py:378-380).  The persistent pass records such events in `flags`; if any
 * occurred, the replay launch recomputes the recurrence with exactly those batch-global semantics. */
int biear_adaptive_fwd(const BiearSeqParams* p, void* stream);
/* Whole backward recurrence (t = T-2 .. 0) in one persistent cluster kernel: applies the band stage's closed-form
 * dQ, carries dL/dQ_{t+1} -> dL/dY_t, dL/dh_{t-1} down the chain in shared memory / registers and
 * writes the per-sample pre-activation gradients GG / G_* (tile layout). */
int biear_adaptive_bwd(const BiearSeqParams* p, void* stream);

/*
 * Weight gradients of the controllers from tile-layout operands, all layers in two launches (split-K partials, then one
 * fixed-order reduction; deterministic, no atomics).  One job per parameter tensor:
 *   matrix job (Di > 0)   dW[g][o][i] = sum_{k < chunks, r < R} A[g][k][o][r] * Bm[g][k][i][r]      (o < Do, i < Di)
 *   diagonal job (Di = 0) dW[g][o]    = sum_{k, r} A[g][k][o][r] * Bm[g][k][o][r]                    (LayerNorm weight)
 *   both                  db[g][o]    = sum_{k, r} A[g][k][o][r]                                      (db nullable)
 * A is (G, chunks, >=Do, R) with chunk stride a_chunk_stride floats and group stride a_group_stride, likewise Bm;
 * R = tile_rows (16 or 32).  Replaces the weight-gradient GEMMs and
 * reductions autograd runs for model_torch.py:256-267.
 */
#define BIEAR_WGRAD_MAX_JOBS 8
typedef struct BiearWgradJob {
    const float* A; int64_t a_group_stride, a_chunk_stride; int32_t Do;
    const float* Bm; int64_t b_group_stride, b_chunk_stride; int32_t Di;
    int64_t chunks;
    float* dW; float* db;
    /* output strides in floats; 0 selects dense: dW (G, Do, Di) / (G, Do), db (G, Do).  Non-zero strides let a job write
       a row block of a larger parameter gradient (e.g. the r,z rows and the n rows of GRU weight_hh) in place. */
    int64_t dw_group_stride, dw_row_stride, db_group_stride;
    /* optional second copy of a matrix job's result, scaled: dW2[g][o][i] = scale2 * dW[g][o][i] with the same strides
       (NULL: none).  The dual controller's input is [c, 0.2 c.detach()] (model_torch.py:351-359), so the gradient of
       weight_ih[:, N:] is 0.2 x that of weight_ih[:, :N]: one job fills both halves of the parameter's gradient. */
    float* dW2; float scale2;
} BiearWgradJob;
/* Floats of scratch biear_ctrl_wgrad needs for these jobs (-1 on invalid arguments). */
int64_t biear_wgrad_scratch_floats(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows);
int biear_ctrl_wgrad(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows, float* scratch, void* stream);
/* The same jobs on the tcgen05 tensor cores (128 x Di output tiles accumulated in TMEM, TF32 operands with the 3-way
 * hi/lo split: fp32-accurate).  Matrix jobs need Di <= 128 and Di % 4 == 0; scratch 16-byte aligned, sized by
 * biear_wgrad_scratch_floats_tc. */
int64_t biear_wgrad_scratch_floats_tc(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows);
int biear_ctrl_wgrad_tc(const BiearWgradJob* jobs, int n_jobs, int G, int tile_rows, float* scratch, void* stream);

/*
 * Broadband interaural cross-correlation feature.  Replaces utils.py:390-420
 * (compute_cross_correlation_feature; byte-identical copy in create_h5_data/utils_save.py).
 *   wavL, wavR (B, nsamp) with row stride; lags k_min..k_max (inclusive) are the integer lags the
 *   reference keeps (|k|/fs <= max_lag); interp_idx/interp_frac (num_lags) give, for every output
 *   point, the left neighbour inside the cropped lag axis and the linear weight of its right
 *   neighbour (np.interp semantics, computed by the host in float64).
 *   cc (B, num_lags) out.
 */
int biear_cc_fwd(const float* wavL, const float* wavR, int64_t B, int64_t nsamp, int64_t row_stride,
                 int k_min, int k_max, const int32_t* interp_idx, const float* interp_frac,
                 int num_lags, float* cc, void* stream);

/*
 * Q regularisers of the training loss, value and gradient in one launch.  Replaces train_biear.py:476-490 on
 * Q = (QA + QB) / 2 (model.last_Q, model_torch.py:1076-1078; QB == NULL: Q = QA):
 *   reg_q = mean((log(Q + 1e-8) - log(Q0 + 1e-8))^2) over (rows, N);  reg_smooth = mean of the squared first difference
 *   of log(Q + 1e-8) along the band axis over (rows, N - 1).
 *   out[0] = w_reg * reg_q + w_smooth * reg_smooth, out[1] = reg_q, out[2] = reg_smooth.
 *   gQ (rows, N), nullable: d out[0] / d QA (identical to d out[0] / d QB when QB is given).
 *   workspace: biear_q_regularizers_workspace_floats() floats whose first 4 bytes the caller zeroes ONCE; the kernel
 *   leaves them zero again (launches sharing a workspace must be stream-ordered).  Deterministic for given sizes.
 */
int64_t biear_q_regularizers_workspace_floats(void);
int biear_q_regularizers(const float* QA, const float* QB, const float* Q0, int64_t rows, int N, float w_reg,
                         float w_smooth, float* out, float* gQ, float* workspace, void* stream);

/*
 * The per-sector heads of the back-end, all S heads in ONE forward and ONE backward launch (+ two
 * fixed-order reductions).  Replaces SubHead.forward and the loop over self.subheads, model_torch.py:869-906, 941-955,
 * 1096-1110 (80 small GEMMs and ~100 element-wise launches forward, three times that backward):
 *   h = Dropout_0.2(ReLU(Linear(D,100)(body)));  three branches Linear(100,50)-ReLU-Linear(50,10)-ReLU-Linear(10,k):
 *   sound (k = 1, logit), aoa (k = 1, sigmoid applied), dist (k = C, logits).
 * wptr: DEVICE array of S * biear_heads_tensors_per_head() (= 20) pointers to the heads' parameters in state-dict order
 *   (shared.0.weight, shared.0.bias, sound.0.weight, sound.0.bias, sound.2.*, sound.4.*, aoa.*, dist.*), torch (out, in)
 *   layout -- the nn.Linear parameters themselves, no packing copy.
 * Outputs sound (B,S), aoa (B,S), dist (B,S,C).  Dropout: Philox keyed by (seed | *seed_ptr, head, row, unit quad),
 *   regenerated by the backward (which recomputes the forward of its tile: no activations are saved).
 * Backward: g_sound / g_aoa (B,S), g_dist (B,S,C), each nullable -> d_body (B,D) and dw (S, biear_heads_flat_floats(D,C))
 *   = every head's parameter gradients concatenated in state-dict order; scratch d_body_part (S,B,D) and
 *   dw_part (ceil(B / biear_heads_tile_rows()), S, flat).  Deterministic (no atomics).
 */
typedef struct BiearHeadsParams {
    int32_t B, S, D, C;              /* clips, sectors, body width (multiple of 4, <= 200), distance classes (<= 8) */
    int32_t training;                /* dropout on */
    uint64_t seed;
    const uint64_t* seed_ptr;        /* optional device seed (overrides `seed`; for captured graphs) */
    const float* body;               /* (B, D) */
    const float* const* wptr;        /* device table of parameter pointers, see above */
    float *sound, *aoa, *dist;
    const float *g_sound, *g_aoa, *g_dist;
    float *d_body_part, *dw_part, *d_body, *dw;
} BiearHeadsParams;
int biear_heads_tile_rows(void);
int biear_heads_tensors_per_head(void);
int64_t biear_heads_flat_floats(int D, int C);
int biear_heads_fwd(const BiearHeadsParams* p, void* stream);
int biear_heads_bwd(const BiearHeadsParams* p, void* stream);

/*
 * The recurrence of ONE GRU layer (torch.nn.GRU semantics, gate order r, z, n; h_{-1} = 0) as a persistent cluster kernel,
 * forward and backward.  Used for the wide first layer of the back-end's ILD / IPD encoders (model_torch.py:828-867:
 * nn.GRU(100 -> 200) over the 19 frames), which the library runs as 19 x (GEMM + cell kernel) per direction.
 * The input projection and everything that is a plain GEMM stays with the caller (a library GEMM each):
 *   forward : gi (B,T,3H) = x W_ih^T + b_ih in  ->  h_seq (B,T,H), h_prev (B,T,H) = h_seq shifted by one step (h_prev[:,0] = 0),
 *             gates (B,T,4,H) = r, z, n, W_hn h + b_hn out
 *   backward: dh_seq (B,T,H) = dL/dh_t from the consumer of the sequence in  ->  dgi, dgh (B,T,3H) out, from which
 *             dx = dgi W_ih, dW_ih = dgi^T x, db_ih = sum dgi, dW_hh = dgh^T h_prev, db_hh = sum dgh.
 * workspace: biear_gru_workspace_floats(H) floats (the forward's weight images, packed by biear_gru_fwd itself; the backward
 * reads w_hh directly, so w_hh must be unchanged between the two calls of a step).  H % 4 == 0 and the shared-memory
 * budget (H <= 232): ask biear_gru_supported(H).
 */
typedef struct BiearGruParams {
    int32_t B, T, H, I;              /* rows, steps, hidden units, input width (informational) */
    const float* gi;
    const float *w_hh, *b_hh;        /* (3H, H), (3H): the nn.GRU parameters themselves */
    float *h_seq, *h_prev, *gates;
    const float* dh_seq;
    float *dgi, *dgh;
    float* workspace;
} BiearGruParams;
int biear_gru_supported(int H);
int64_t biear_gru_workspace_floats(int H);
int biear_gru_fwd(const BiearGruParams* p, void* stream);
int biear_gru_bwd(const BiearGruParams* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BIEAR_B200_H */
