// Host emulation of biear_b200/csrc/cc_dev.cuh + the chunk loop of cc.cu: runs every "thread" of one
// CTA sequentially and checks the raw lag sums against a float64 direct evaluation.
// usage: cc_emu <nsamp> <k_min> <k_max>; prints the max-abs-normalised error of c[k].
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "cc_dev.cuh"
using namespace biear;

int main(int argc, char** argv) {
    const long long nsamp = argc > 1 ? atoll(argv[1]) : 16000;
    const int k_min = argc > 2 ? atoi(argv[2]) : -48;
    const int k_max = argc > 3 ? atoi(argv[3]) : 48;
    const CcPlan p = cc_make_plan(nsamp, k_min, k_max);
    std::vector<float> L(nsamp), R(nsamp);
    srand(11);
    float ar = 0.f;
    for (long long i = 0; i < nsamp; ++i) {
        ar = 0.9f * ar + ((float)rand() / RAND_MAX - 0.5f);
        L[i] = ar + 0.3f;
        R[i] = (i >= 5 ? L[i - 5] * 0.7f : 0.f) + 0.05f * ((float)rand() / RAND_MAX - 0.5f) - 0.1f;
    }
    double ml = 0, mr = 0;
    for (long long i = 0; i < nsamp; ++i) { ml += L[i]; mr += R[i]; }
    ml /= nsamp; mr /= nsamp;
    const float mean_l = (float)ml, mean_r = (float)mr;

    const int lag_span = kCcLagBlock * p.lag_blocks;
    std::vector<float> sR(cc_r_floats(p)), sL(cc_l_floats(p));
    std::vector<float> acc((size_t)kCcThreads * 16, 0.f);
    for (int c = 0; c < p.n_chunks; ++c) {
        const long long c0 = (long long)c * p.chunk;
        for (int e = 0; e < cc_r_floats(p); ++e) {
            const long long n = c0 + e;
            sR[cc_slot(e)] = n < nsamp ? R[n] - mean_r : 0.f;
        }
        for (int e = 0; e < cc_l_floats(p); ++e) {
            const long long n = c0 + k_min + e;
            sL[cc_slot(e)] = (n >= 0 && n < nsamp) ? L[n] - mean_l : 0.f;
        }
        for (int tid = 0; tid < kCcThreads; ++tid) {
            const int lb = tid / p.strips, s = tid - lb * p.strips;
            if (lb >= p.lag_blocks) continue;
            for (int i = 0; i < p.m; ++i)
                cc_unit((const float4*)sL.data(), (const float4*)sR.data(), s + p.strips * i, lb, &acc[(size_t)tid * 16]);
        }
    }
    std::vector<double> ck(p.nlags, 0.0);
    for (int tid = 0; tid < kCcThreads; ++tid) {
        const int lb = tid / p.strips;
        if (lb >= p.lag_blocks) continue;
        for (int i = 0; i < 16; ++i) {
            const int li = lb * 16 + i;
            if (li < p.nlags) ck[li] += acc[(size_t)tid * 16 + i];
        }
    }
    (void)lag_span;
    double maxref = 0, maxerr = 0;
    for (int li = 0; li < p.nlags; ++li) {
        const int k = k_min + li;
        double ref = 0;
        for (long long n = 0; n < nsamp; ++n) {
            const long long j = n + k;
            if (j >= 0 && j < nsamp) ref += ((double)L[j] - ml) * ((double)R[n] - mr);
        }
        maxref = fmax(maxref, fabs(ref));
        maxerr = fmax(maxerr, fabs(ref - ck[li]));
    }
    printf("%.3e plan: lags=%d blocks=%d strips=%d m=%d chunk=%d n_chunks=%d\n", maxerr / maxref, p.nlags, p.lag_blocks,
           p.strips, p.m, p.chunk, p.n_chunks);
    return maxerr / maxref < 2e-6 ? 0 : 1;
}
