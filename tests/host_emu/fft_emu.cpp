// Host emulation of biear_b200/csrc/fft_dev.cuh: runs the 64-"thread" Stockham passes sequentially
// and checks the unpacked real spectrum against a float64 DFT.  Prints the max-abs-normalised error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fft_dev.cuh"
using namespace biear;

int main() {
    std::vector<float2> tw(1024);
    for (int m = 0; m < 1024; ++m) {
        double a = -2.0 * M_PI * m / 1024.0;
        tw[m] = make_float2((float)cos(a), (float)sin(a));
    }
    const int win = 842;
    std::vector<float> wav(16000), wfn(win);
    srand(7);
    for (auto& v : wav) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (int i = 0; i < win; ++i) wfn[i] = 0.5f - 0.5f * (float)cos(2.0 * M_PI * i / win);
    double worst = 0;
    for (int t : {0, 7, 18}) {
        std::vector<float2> a(512), b(512);
        for (int n = 0; n < 512; ++n)
            a[n] = make_float2(frame_sample(wav.data(), 16000, 16000, (long long)t * 842, 2 * n, win, wfn.data(), true),
                               frame_sample(wav.data(), 16000, 16000, (long long)t * 842, 2 * n + 1, win, wfn.data(), true));
        for (int j = 0; j < 64; ++j) fft512_pass(a.data(), b.data(), j, 1, tw.data());
        for (int j = 0; j < 64; ++j) fft512_pass(b.data(), a.data(), j, 8, tw.data());
        for (int j = 0; j < 64; ++j) fft512_pass(a.data(), b.data(), j, 64, tw.data());
        std::vector<float2> X(513);
        for (int k = 0; k <= 256; ++k) {
            float2 xk, xm;
            rfft_unpack(b.data(), k, tw.data(), xk, xm);
            X[k] = xk;
            X[512 - k] = xm;
        }
        double maxref = 0, maxerr = 0;
        for (int k = 0; k <= 512; ++k) {
            double re = 0, im = 0;
            for (int i = 0; i < win; ++i) {
                double x = (double)wav[t * 842 + i] * (double)wfn[i];
                double ang = -2.0 * M_PI * (double)k * i / 1024.0;
                re += x * cos(ang);
                im += x * sin(ang);
            }
            maxref = fmax(maxref, hypot(re, im));
            maxerr = fmax(maxerr, hypot(X[k].x - re, X[k].y - im));
        }
        worst = fmax(worst, maxerr / maxref);
    }
    printf("%.3e\n", worst);
    return worst < 1e-6 ? 0 : 1;
}
