// Host harness for the lane-uniform pass 2 of asrk_fft.cuh (split_lane): emulates the ten
// roles of one 400-sample frame read from stdin, with the generated tables of
// asrk_tables.inc, and prints 4*|X[k]|^2, k<200.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../asr_dfcnn_transformer_b200/csrc/asrk_fft.cuh"
using namespace asrk;

static const double kTab[1200] = {
#include "../../asr_dfcnn_transformer_b200/csrc/asrk_tables.inc"
};

int main() {
    std::vector<double> x(400);
    for (int i = 0; i < 400; ++i)
        if (scanf("%lf", &x[i]) != 1) return 1;
    const double* W = kTab;
    const cplx* TW = reinterpret_cast<const cplx*>(kTab + 400);   // [r][k1]
    const cplx* P = reinterpret_cast<const cplx*>(kTab + 800);
    std::vector<cplx> Y(200);   // [k1][n2]
    for (int r = 0; r < 10; ++r) {
        cplx z[20], y[20];
        for (int n1 = 0; n1 < 20; ++n1) {
            int m = 10 * n1 + r;
            z[n1] = cplx{x[2 * m] * W[2 * m], x[2 * m + 1] * W[2 * m + 1]};
        }
        fft200_pass1(z, TW + r * 20, y);
        for (int k1 = 0; k1 < 20; ++k1) Y[k1 * 10 + r] = y[k1];
    }
    std::vector<double> out(204, -1.0);
    std::vector<int> cnt(204, 0);
    for (int j = 0; j < 10; ++j) {
        cplx in[10], za[10], zb[10];
        for (int n2 = 0; n2 < 10; ++n2) in[n2] = Y[lane_k1a(j) * 10 + n2];
        dft10(in, za);
        for (int n2 = 0; n2 < 10; ++n2) in[n2] = Y[lane_k1b(j) * 10 + n2];
        dft10(in, zb);
        const bool j0 = (j == 0);
        auto loadP = [&](int s) { return P[(s == 10 && !j0) ? 0 : lane_bin(j, s)]; };
        auto emit = [&](int s, double pk, double pm) {
            if (s == 10 && !j0) return;
            const int k = lane_bin(j, s);
            out[200 - k] = pm; cnt[200 - k]++;
            out[k] = pk; cnt[k]++;
        };
        split_lane(j0, za, zb, loadP, emit);
    }
    for (int k = 0; k < 200; ++k) {
        // bin 100 is written twice by role 0 (slot 5, pk last), everything else once
        if (cnt[k] != (k == 100 ? 2 : 1)) { fprintf(stderr, "bin %d emitted %d times\n", k, cnt[k]); return 2; }
        printf("%.17g\n", out[k]);
    }
    return 0;
}
