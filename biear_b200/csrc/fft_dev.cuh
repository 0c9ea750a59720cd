// 1024-point real FFT of one Hann-windowed frame, as device-side building blocks.
//
// A frame is packed as 512 complex points z[n] = x[2n] + i x[2n+1], transformed with a 3-pass
// radix-8 Stockham FFT by a group of 64 threads, then unpacked into the 513 one-sided bins of the
// real spectrum.  Everything is HOST_DEVICE so the index arithmetic can be checked on the CPU
// (tests/test_host_emulation.py) before any GPU time is spent.
//
// Replaces torch.fft.rfft(frame * win_fn, n=1024) at model_torch.py:334-335 of the reference.
#pragma once
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define BIEAR_HD __host__ __device__ __forceinline__
#else
#define BIEAR_HD inline
#endif

namespace biear {

constexpr int kNfft = 1024;
constexpr int kNhalf = 512;          // complex FFT length
constexpr int kBins = 513;           // one-sided bins
constexpr int kFftThreads = 64;      // threads cooperating on one frame

BIEAR_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
BIEAR_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
BIEAR_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
BIEAR_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
BIEAR_HD float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

BIEAR_HD void bfly2(float2& a, float2& b) {
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// 4-point DFT; results come out in bit-reversed slots (0,2,1,3).
BIEAR_HD void bfly4(float2& v0, float2& v1, float2& v2, float2& v3) {
    bfly2(v0, v2);
    bfly2(v1, v3);
    v3 = mul_neg_i(v3);
    bfly2(v0, v1);
    bfly2(v2, v3);
}

// 8-point DFT in registers; natural-order output k lives in slot rev3(k) = {0,4,2,6,1,5,3,7}[k].
BIEAR_HD void bfly8(float2 v[8]) {
    const float s = 0.70710678118654752440f;
    bfly2(v[0], v[4]);
    bfly2(v[1], v[5]);
    bfly2(v[2], v[6]);
    bfly2(v[3], v[7]);
    v[5] = make_float2(s * (v[5].x + v[5].y), s * (v[5].y - v[5].x));      // * exp(-i pi/4)
    v[6] = mul_neg_i(v[6]);                                                  // * exp(-i pi/2)
    v[7] = make_float2(s * (v[7].y - v[7].x), -s * (v[7].x + v[7].y));     // * exp(-3i pi/4)
    bfly4(v[0], v[1], v[2], v[3]);
    bfly4(v[4], v[5], v[6], v[7]);
}

// Shared-memory slot of complex point i: one pad slot every 8 points keeps the stride-8 and
// stride-64 accesses of the Stockham passes off the same bank pairs.
BIEAR_HD int fft_slot(int i) { return i + (i >> 3); }
constexpr int kFftSlots = kNhalf + (kNhalf >> 3);   // 576 float2 per frame buffer

// Twiddle + radix-8 butterfly of one thread (j in [0,64)) for the Stockham pass with sub-transform
// length Ns in {1, 8, 64}; v[r] holds input point j + 64 r.  tw1024[m] = exp(-2 pi i m / 1024).
BIEAR_HD void fft512_butterfly(float2 v[8], int j, int Ns, const float2* tw1024) {
    const int k = j & (Ns - 1);
    if (Ns > 1) {
        // exp(-2 pi i r k / (8 Ns)) = tw1024[r k * (1024 / (8 Ns))]
        const int step = k * (kNfft / (8 * Ns));
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], tw1024[r * step]);
    }
    bfly8(v);
}

// The same with the 7 twiddles of the pass already in registers (they depend on the thread and the pass only, not on
// the frame: tw[r - 1] = tw1024[r * (j & (Ns - 1)) * (1024 / (8 Ns))]).
BIEAR_HD void fft512_butterfly_reg(float2 v[8], const float2 tw[7]) {
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], tw[r - 1]);
    bfly8(v);
}

// Scatter the butterfly outputs to their Stockham positions (through the slot map when Padded).
template <bool Padded>
BIEAR_HD void fft512_scatter(const float2 v[8], float2* out, int j, int Ns) {
    const int k = j & (Ns - 1);
    const int j0 = ((j - k) << 3) + k;
    const int rev[8] = {0, 4, 2, 6, 1, 5, 3, 7};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = j0 + r * Ns;
        out[Padded ? fft_slot(i) : i] = v[rev[r]];
    }
}

template <bool Padded>
BIEAR_HD void fft512_gather(float2 v[8], const float2* in, int j) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = j + r * 64;
        v[r] = in[Padded ? fft_slot(i) : i];
    }
}

// Whole pass on unpadded buffers (host emulation / reference form).
BIEAR_HD void fft512_pass(const float2* in, float2* out, int j, int Ns, const float2* tw1024) {
    float2 v[8];
    fft512_gather<false>(v, in, j);
    fft512_butterfly(v, j, Ns, tw1024);
    fft512_scatter<false>(v, out, j, Ns);
}

// Unpack bins k and 512-k (k in [0,256]) of the real spectrum from the packed transform Z.
template <bool Padded = false>
BIEAR_HD void rfft_unpack(const float2* Z, int k, const float2* tw1024, float2& Xk, float2& Xmk) {
    const int km = (kNhalf - k) & (kNhalf - 1);
    const float2 zk = Z[Padded ? fft_slot(k) : k];
    const float2 zc = cconj(Z[Padded ? fft_slot(km) : km]);
    const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
    const float2 o = mul_neg_i(d);
    const float2 wo = cmul(tw1024[k], o);
    Xk = cadd(e, wo);
    Xmk = cconj(csub(e, wo));
}

// Packed, windowed sample pair n (z[n] = x[2n] w[2n] + i x[2n+1] w[2n+1]) of frame t of a row.
// Samples past min(win, n_fft) are the rfft zero padding; samples past the clip (or past the
// `limit` = max(fs, win) padded length) are zero; frames past the available count are zero.
BIEAR_HD float frame_sample(const float* row, long long nsamp, int limit, long long start, int i, int win,
                            const float* win_fn, bool frame_valid) {
    if (!frame_valid || i >= win || i >= kNfft) return 0.0f;
    const long long s = start + i;
    if (s >= limit || s >= nsamp) return 0.0f;
    return row[s] * win_fn[i];
}

}  // namespace biear
