// Host harness for the 256-point complex transform of the mel front end (csrc/logfbank.cu): the same
// two passes of dft16 with the W256 twiddle in between, on 256 complex points read from stdin
// (re im pairs); prints Z[k], k < 256.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../asr_dfcnn_transformer_b200/csrc/asrk_fft.cuh"
using namespace asrk;

int main() {
    std::vector<cplx> z(256), ex(16 * 17), Z(256);
    for (int i = 0; i < 256; ++i)
        if (scanf("%lf %lf", &z[i].x, &z[i].y) != 2) return 1;
    const double kPi = 3.14159265358979323846;
    for (int r = 0; r < 16; ++r) {                       // pass 1, "lane" r
        cplx a[16];
        for (int m = 0; m < 16; ++m) a[m] = z[r + 16 * m];
        dft16(a);
        for (int k1 = 0; k1 < 16; ++k1) {
            const cplx v = a[4 * (k1 & 3) + (k1 >> 2)];
            const double ang = -2.0 * kPi * (double)(r * k1) / 256.0;
            ex[r * 17 + k1] = (k1 == 0) ? v : cmul(v, cplx{std::cos(ang), std::sin(ang)});
        }
    }
    for (int k1 = 0; k1 < 16; ++k1) {                    // pass 2, "lane" k1
        cplx a[16];
        for (int n2 = 0; n2 < 16; ++n2) a[n2] = ex[n2 * 17 + k1];
        dft16(a);
        for (int k2 = 0; k2 < 16; ++k2) Z[k1 + 16 * k2] = a[4 * (k2 & 3) + (k2 >> 2)];
    }
    for (int k = 0; k < 256; ++k) printf("%.17g %.17g\n", Z[k].x, Z[k].y);
    return 0;
}
