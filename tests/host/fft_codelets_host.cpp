// Host harness for asrk_fft.cuh + asrk_tables.inc: runs the exact pass1/pass2 code
// and the generated constant tables the CUDA kernel uses on one 400-sample frame
// (raw samples, read from stdin) and prints 4*|X[k]|^2, k<200.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../asr_dfcnn_transformer_b200/csrc/asrk_fft.cuh"
using namespace asrk;

static const double kTab[1200] = {
#include "../../asr_dfcnn_transformer_b200/csrc/asrk_tables.inc"
};

int main() {
    std::vector<double> xw(400);
    for (int i = 0; i < 400; ++i) {
        if (scanf("%lf", &xw[i]) != 1) return 1;
        xw[i] *= kTab[i];                                      // the window
    }
    const cplx* TW = reinterpret_cast<const cplx*>(kTab + 400);   // [r][k1]
    const cplx* PT = reinterpret_cast<const cplx*>(kTab + 800);
    std::vector<cplx> Y(200);   // [k1][n2]
    for (int r = 0; r < 10; ++r) {
        cplx z[20], y[20];
        for (int n1 = 0; n1 < 20; ++n1) {
            int m = 10 * n1 + r;
            z[n1] = cplx{xw[2 * m], xw[2 * m + 1]};
        }
        fft200_pass1(z, TW + r * 20, y);
        for (int k1 = 0; k1 < 20; ++k1) Y[k1 * 10 + r] = y[k1];
    }
    std::vector<cplx> P(PT, PT + 200);
    std::vector<double> out(200, -1.0);
    std::vector<int> cnt(200, 0);
    for (int j = 0; j < 10; ++j) {
        auto loadY = [&](int k1, int n2) { return Y[k1 * 10 + n2]; };
        auto emit = [&](int k, double p) { out[k] = p; cnt[k]++; };
        fft200_pass2(j, loadY, P.data(), emit);
    }
    for (int k = 0; k < 200; ++k) {
        if (cnt[k] != 1) { fprintf(stderr, "bin %d emitted %d times\n", k, cnt[k]); return 2; }
        printf("%.17g\n", out[k]);
    }
    return 0;
}
