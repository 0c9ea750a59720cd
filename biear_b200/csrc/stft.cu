// Framing + window + zero-padded 1024-point real FFT of every (row, frame): one launch for the
// whole batch, both ears, all frames (the spectrum does not depend on Q, so it is hoisted out of the
// reference's per-frame loop).
//
// Replaces model_torch.py:289-312 (_frame_1s) and :334-335 (frame * win_fn, torch.fft.rfft(n=n_fft))
// of the reference.
//
// Mapping: 64 threads cooperate on one frame (packed 512-point complex radix-8 Stockham FFT, three
// passes, one padded 4.6 KB shared buffer per frame; the first pass reads straight from HBM), four
// frames per CTA, CTAs stride over the frame list.  HBM traffic = the waveform once in, X once out.
#include "common.cuh"
#include "fft_dev.cuh"

namespace biear {

constexpr int kFramesPerCta = 4;
constexpr int kStftThreads = kFramesPerCta * kFftThreads;

struct StftArgs {
    const float* wav;
    const float* wav2;    // rows >= rows_first come from wav2 (the second ear), row index rebased; nullable
    long long rows_first;
    int32_t* ready;       // nullable: ready[row * T + t] = 1 once X[row][t] is complete (release order)
    void* counter;        // with ready: 8-byte work counter behind the flags (cleared with them)
    int frame_major;      // work order: 1 = frame 0 of every row first, then frame 1, ... (streaming consumers)
    long long rows, nsamp, row_stride;
    const float* win_fn;
    const float2* tw;
    int limit;        // padded clip length: max(fs, win)
    int n_avail;      // frames the reference's unfold() yields
    int T, win, hop;
    float2* X;        // (rows, T, 513)
    long long n_frames;
};

__global__ void __launch_bounds__(kStftThreads, 2) stft_fwd_kernel(const StftArgs a) {
    __shared__ float2 s_tw[kNfft];
    __shared__ float2 s_buf[kFramesPerCta][kFftSlots];

    for (int i = threadIdx.x; i < kNfft; i += kStftThreads) {
        s_tw[i] = a.tw[i];
    }

    const int g = threadIdx.x / kFftThreads;
    const int j = threadIdx.x % kFftThreads;
    float2* buf = s_buf[g];
    const int valid_len = min(min(a.win, kNfft), a.limit);
    const long long stride = (long long)gridDim.x * kFramesPerCta;

    // work item -> (row, frame)
    auto item_of = [&](long long fi, bool live, long long& row, int& t) {
        if (!live) {
            row = 0;
            t = 0;
        } else if (a.frame_major) {
            t = (int)(fi / a.rows);
            row = fi % a.rows;
        } else {
            row = fi / a.T;
            t = (int)(fi % a.T);
        }
    };
    // Raw samples of one frame (this thread's 16), NOT yet windowed: issued one trip ahead so that the HBM latency of
    // the next frame group hides behind the three FFT passes of the current one.
    auto fetch = [&](long long base, float2 (&raw)[8]) {
        const long long fi = base + g;
        const bool live = fi < a.n_frames;
        long long row;
        int t;
        item_of(fi, live, row, t);
        const bool frame_valid = live && t < a.n_avail;
        const long long start = (long long)t * a.hop;
        const float* wrow = (row < a.rows_first ? a.wav + row * a.row_stride : a.wav2 + (row - a.rows_first) * a.row_stride);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i0 = 2 * (j + r * 64);
            float x0 = 0.f, x1 = 0.f;
            if (frame_valid) {
                const long long s0 = start + i0;
                if (i0 < valid_len && s0 < a.limit && s0 < a.nsamp) x0 = __ldg(wrow + s0);
                if (i0 + 1 < valid_len && s0 + 1 < a.limit && s0 + 1 < a.nsamp) x1 = __ldg(wrow + s0 + 1);
            }
            raw[r] = make_float2(x0, x1);
        }
    };

    // Work distribution.  Static (no ready flags): CTA c takes frame groups c, c + grid, ...  Dynamic (streaming hand-over):
    // groups are claimed from a device counter in order, so that frame-major order holds no matter how many CTAs are
    // resident at a time (next to a persistent kernel only the SMs it leaves idle take CTAs of this one).
    __shared__ long long s_claim[2];
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(a.counter);
    const bool dynamic = counter != nullptr;
    auto claim = [&]() { return (long long)atomicAdd(counter, 1ull) * kFramesPerCta; };
    float2 nxt[8];
    long long base = (long long)blockIdx.x * kFramesPerCta, next_base = base + stride;
    if (dynamic) {
        if (threadIdx.x == 0) {
            s_claim[0] = claim();
            s_claim[1] = claim();
        }
        __syncthreads();
        base = s_claim[0];
        next_base = s_claim[1];
    }
    if (base < a.n_frames) fetch(base, nxt);
    // this thread's 16 window taps are the same for every frame it touches: registers, not shared memory
    float2 wreg[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i0 = 2 * (j + r * 64);
        wreg[r] = make_float2(i0 < a.win ? __ldg(a.win_fn + i0) : 0.0f, i0 + 1 < a.win ? __ldg(a.win_fn + i0 + 1) : 0.0f);
    }
    __syncthreads();   // twiddles in shared memory
    // the twiddles of passes 2 and 3 depend on the thread only: registers, not 14 shared loads per frame
    float2 tw8[7], tw64[7];
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        tw8[r - 1] = s_tw[r * (j & 7) * (kNfft / 64)];
        tw64[r - 1] = s_tw[r * (j & 63) * (kNfft / 512)];
    }
    for (int trip = 0; base < a.n_frames; ++trip) {
        const long long fi = base + g;
        const bool live = fi < a.n_frames;
        // pass 1 (Ns = 1): packed, windowed samples
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            v[r] = make_float2(nxt[r].x * wreg[r].x, nxt[r].y * wreg[r].y);
        }
        if (next_base < a.n_frames) fetch(next_base, nxt);
        fft512_butterfly(v, j, 1, s_tw);
        fft512_scatter<true>(v, buf, j, 1);
        __syncthreads();
        // (dynamic) claim the group after next: slot trip & 1 was consumed as this trip's base before the barrier above
        if (dynamic && threadIdx.x == 0) s_claim[trip & 1] = claim();
        fft512_gather<true>(v, buf, j);
        fft512_butterfly_reg(v, tw8);
        __syncthreads();
        fft512_scatter<true>(v, buf, j, 8);
        __syncthreads();
        fft512_gather<true>(v, buf, j);
        fft512_butterfly_reg(v, tw64);
        __syncthreads();
        fft512_scatter<true>(v, buf, j, 64);
        __syncthreads();

        long long orow;
        int ot;
        item_of(fi, live, orow, ot);
        if (live) {
            float2* out = a.X + (orow * a.T + ot) * kBins;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int k = j + r * 64;          // 0..255
                float2 xk, xm;
                rfft_unpack<true>(buf, k, s_tw, xk, xm);
                out[k] = xk;
                out[kNhalf - k] = xm;
            }
            if (j == 0) {
                float2 xk, xm;
                rfft_unpack<true>(buf, 256, s_tw, xk, xm);
                out[256] = xk;
            }
        }
        if (a.ready) __threadfence();   // this thread's part of the frame is visible device-wide ...
        __syncthreads();   // unpack reads done before the next trip scatters into the buffer
        if (a.ready && live && j == 0)   // ... every thread's is: publish the frame
            *reinterpret_cast<volatile int32_t*>(a.ready + orow * a.T + ot) = 1;
        base = next_base;
        next_base = dynamic ? s_claim[trip & 1] : next_base + stride;   // written before the barrier above
    }
}

}  // namespace biear

static int stft_launch(const float* wav, const float* wav2, int64_t rows_first, int64_t rows, int64_t nsamp,
                       int64_t wav_row_stride, const float* win_fn, int fs, int T, int win, int hop, int n_fft, float* X,
                       int32_t* ready, int frame_major, void* stream, const char* who) {
    using namespace biear;
    BIEAR_REQUIRE(n_fft == kNfft, "biear_stft_fwd: n_fft=%d unsupported (only 1024)", n_fft);
    BIEAR_REQUIRE(rows >= 0 && nsamp >= 0 && fs >= 1 && T >= 1 && win >= 1 && hop >= 1,
                  "biear_stft_fwd: bad shape rows=%lld nsamp=%lld fs=%d T=%d win=%d hop=%d", (long long)rows,
                  (long long)nsamp, fs, T, win, hop);
    BIEAR_REQUIRE(wav_row_stride >= nsamp, "biear_stft_fwd: row stride %lld < nsamp %lld",
                  (long long)wav_row_stride, (long long)nsamp);
    if (rows == 0) return 0;
    BIEAR_REQUIRE(wav && win_fn && X, "biear_stft_fwd: null pointer");
    int err = 0;
    cudaStream_t st = as_stream(stream);
    StftArgs a;
    a.tw = twiddle_table(st, &err);
    if (err) return err;
    a.wav = wav;
    a.wav2 = wav2;
    a.rows_first = rows_first;
    a.ready = ready;
    a.counter = ready ? reinterpret_cast<void*>(ready + ((rows * T + 1) & ~1LL)) : nullptr;
    a.frame_major = frame_major;
    a.rows = rows;
    a.nsamp = nsamp;
    a.row_stride = wav_row_stride;
    a.win_fn = win_fn;
    a.limit = fs > win ? fs : win;
    a.n_avail = (a.limit - win) / hop + 1;   // torch.Tensor.unfold(size=win, step=hop) frame count
    a.T = T;
    a.win = win;
    a.hop = hop;
    a.X = reinterpret_cast<float2*>(X);
    a.n_frames = rows * T;
    const long long ctas_needed = (a.n_frames + kFramesPerCta - 1) / kFramesPerCta;
    // 2 resident CTAs per SM (window taps and twiddles in registers, 26 KB of shared memory each), every CTA pipelining over the same number of
    // frame groups: one even wave instead of a ragged one
    const long long cap = (long long)kSmCountB200 * 2;
    const long long trips = (ctas_needed + cap - 1) / cap;
    const int grid = ready ? (int)(ctas_needed < cap ? ctas_needed : cap)      // dynamic claiming: any grid is in order
                           : (int)((ctas_needed + trips - 1) / trips);
    (void)who;
    stft_fwd_kernel<<<grid, kStftThreads, 0, st>>>(a);
    BIEAR_LAUNCH_CHECK("stft_fwd_kernel");
    return 0;
}

extern "C" int biear_stft_fwd(const float* wav, int64_t rows, int64_t nsamp, int64_t wav_row_stride,
                              const float* win_fn, int fs, int T, int win, int hop, int n_fft, float* X,
                              void* stream) {
    return stft_launch(wav, nullptr, rows, rows, nsamp, wav_row_stride, win_fn, fs, T, win, hop, n_fft, X, nullptr, 0, stream,
                       "biear_stft_fwd");
}

extern "C" int biear_stft_fwd_pair(const float* wavA, const float* wavB, int64_t rows_each, int64_t nsamp,
                                   int64_t wav_row_stride, const float* win_fn, int fs, int T, int win, int hop, int n_fft,
                                   float* X, int32_t* ready, void* stream) {
    BIEAR_REQUIRE(wavB != nullptr || rows_each == 0, "biear_stft_fwd_pair: null second waveform");
    return stft_launch(wavA, wavB, rows_each, 2 * rows_each, nsamp, wav_row_stride, win_fn, fs, T, win, hop, n_fft, X, ready,
                       1, stream, "biear_stft_fwd_pair");
}

// ------------------------------------------------------------------------------------------------------------------
// 16-bit PCM -> float32 (the end-to-end input path: clips travel host -> device as int16, half the bytes of float32)
// ------------------------------------------------------------------------------------------------------------------
namespace biear {
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const int16_t* __restrict__ in, float* __restrict__ out,
                                                           long long n, float scale) {
    // 8 samples per thread and step: one 128-bit load, two 128-bit stores; grid-stride
    const long long n8 = n >> 3;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(in) + i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[2 * j] = (float)(short)(w[j] & 0xffff) * scale;
            f[2 * j + 1] = (float)(short)(w[j] >> 16) * scale;
        }
        float4* o = reinterpret_cast<float4*>(out) + 2 * i;
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (float)in[i] * scale;
}
}  // namespace biear

extern "C" int biear_pcm16_to_f32(const int16_t* in, float* out, int64_t n, float scale, void* stream) {
    using namespace biear;
    BIEAR_REQUIRE(n >= 0, "biear_pcm16_to_f32: negative length");
    if (n == 0) return 0;
    BIEAR_REQUIRE(in && out, "biear_pcm16_to_f32: null pointer");
    BIEAR_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "biear_pcm16_to_f32: buffers must be 16-byte aligned");
    const long long blocks = (n / 8 + 255) / 256;
    const int grid = (int)(blocks < 1 ? 1 : (blocks > (long long)kSmCountB200 * 8 ? (long long)kSmCountB200 * 8 : blocks));
    pcm16_to_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, out, (long long)n, scale);
    BIEAR_LAUNCH_CHECK("pcm16_to_f32_kernel");
    return 0;
}
