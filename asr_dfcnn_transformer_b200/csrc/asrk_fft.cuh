// Straight-line fp64 DFT codelets used by the spectrogram kernel.
//
// The 400-point real transform of one windowed frame (util/wav_util.py:70-75 in
// the reference: fft(data_line * w), bins 0..199 kept) is computed as a
// 200-point complex transform of z[m] = x[2m] + i x[2m+1] followed by the
// real-input split.  200 = 20 x 10 (Cooley-Tukey), and both factors are
// themselves done twiddle-free with the Good-Thomas prime-factor map
// (20 = 4 x 5, 10 = 2 x 5).  Everything here is register-resident and fully
// unrolled; indices are compile-time constants.
//
// The header compiles for the host as well (tests/test_fft_codelets.py builds a
// tiny g++ harness around it) so the index maps are checked on the CPU.
#pragma once

#if defined(__CUDACC__)
#define ASRK_HD __host__ __device__ __forceinline__
#else
#define ASRK_HD inline
#endif

namespace asrk {

struct cplx {
    double x, y;
};

ASRK_HD cplx cadd(cplx a, cplx b) { return cplx{a.x + b.x, a.y + b.y}; }
ASRK_HD cplx csub(cplx a, cplx b) { return cplx{a.x - b.x, a.y - b.y}; }
ASRK_HD cplx cmul(cplx a, cplx b) {
    return cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}

// cos(2 pi/5), cos(4 pi/5), sin(2 pi/5), sin(4 pi/5)
#define ASRK_C1 0.30901699437494742410
#define ASRK_C2 -0.80901699437494742410
#define ASRK_S1 0.95105651629515357212
#define ASRK_S2 0.58778525229247312917

// forward 5-point DFT (kernel exp(-2 pi i nk/5)), in place.  32 fp64 instructions: the sine
// terms are folded into the output FMAs (a1 = m1 - i s1 as two chained FMAs per component) and
// the mirrored output is formed as 2 m1 - a1 (one FMA), which saves the separate s1 / s2 values.
#if defined(__CUDACC__)
#define ASRK_FMA(a, b, c) fma((a), (b), (c))
#else
#define ASRK_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif
ASRK_HD void dft5(cplx& a0, cplx& a1, cplx& a2, cplx& a3, cplx& a4) {
    const cplx t1 = cadd(a1, a4), t2 = cadd(a2, a3);
    const cplx t3 = csub(a1, a4), t4 = csub(a2, a3);
    const cplx m1{ASRK_FMA(ASRK_C2, t2.x, ASRK_FMA(ASRK_C1, t1.x, a0.x)), ASRK_FMA(ASRK_C2, t2.y, ASRK_FMA(ASRK_C1, t1.y, a0.y))};
    const cplx m2{ASRK_FMA(ASRK_C1, t2.x, ASRK_FMA(ASRK_C2, t1.x, a0.x)), ASRK_FMA(ASRK_C1, t2.y, ASRK_FMA(ASRK_C2, t1.y, a0.y))};
    a0 = cplx{a0.x + t1.x + t2.x, a0.y + t1.y + t2.y};
    // s1 = S1 t3 + S2 t4, s2 = S2 t3 - S1 t4;  a1 = m1 - i s1, a4 = m1 + i s1, a2 = m2 - i s2, a3 = m2 + i s2
    a1 = cplx{ASRK_FMA(ASRK_S2, t4.y, ASRK_FMA(ASRK_S1, t3.y, m1.x)), ASRK_FMA(-ASRK_S2, t4.x, ASRK_FMA(-ASRK_S1, t3.x, m1.y))};
    a4 = cplx{ASRK_FMA(2.0, m1.x, -a1.x), ASRK_FMA(2.0, m1.y, -a1.y)};
    a2 = cplx{ASRK_FMA(-ASRK_S1, t4.y, ASRK_FMA(ASRK_S2, t3.y, m2.x)), ASRK_FMA(ASRK_S1, t4.x, ASRK_FMA(-ASRK_S2, t3.x, m2.y))};
    a3 = cplx{ASRK_FMA(2.0, m2.x, -a2.x), ASRK_FMA(2.0, m2.y, -a2.y)};
}

// forward 4-point DFT, in place.
ASRK_HD void dft4(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {
    cplx s02 = cadd(a0, a2), d02 = csub(a0, a2);
    cplx s13 = cadd(a1, a3), d13 = csub(a1, a3);
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cplx{d02.x + d13.y, d02.y - d13.x};   // d02 - i d13
    a3 = cplx{d02.x - d13.y, d02.y + d13.x};   // d02 + i d13
}

ASRK_HD void dft2(cplx& a0, cplx& a1) {
    cplx s = cadd(a0, a1), d = csub(a0, a1);
    a0 = s;
    a1 = d;
}

// 20-point forward DFT, in -> out (natural order both sides), prime-factor map
//   n = (5 n1 + 4 n2) mod 20,  k = (5 k1 + 16 k2) mod 20,  n1,k1 < 4, n2,k2 < 5.
ASRK_HD void dft20(const cplx (&in)[20], cplx (&out)[20]) {
    cplx a[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1) a[n1][n2] = in[(5 * n1 + 4 * n2) % 20];
        dft4(a[0][n2], a[1][n2], a[2][n2], a[3][n2]);
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        dft5(a[k1][0], a[k1][1], a[k1][2], a[k1][3], a[k1][4]);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) out[(5 * k1 + 16 * k2) % 20] = a[k1][k2];
    }
}

// 10-point forward DFT: n = (5 n1 + 2 n2) mod 10, k = (5 k1 + 6 k2) mod 10.
ASRK_HD void dft10(const cplx (&in)[10], cplx (&out)[10]) {
    cplx a[2][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        a[0][n2] = in[(2 * n2) % 10];
        a[1][n2] = in[(5 + 2 * n2) % 10];
        dft2(a[0][n2], a[1][n2]);
    }
#pragma unroll
    for (int k1 = 0; k1 < 2; ++k1) {
        dft5(a[k1][0], a[k1][1], a[k1][2], a[k1][3], a[k1][4]);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) out[(5 * k1 + 6 * k2) % 10] = a[k1][k2];
    }
}

// forward 16-point DFT in place (kernel exp(-2 pi i nk/16)): n = 4 n1 + n2, k = k1 + 4 k2;
// on return X[k1 + 4 k2] sits in a[4 k1 + k2]
ASRK_HD void dft16(cplx (&a)[16]) {
    constexpr double C = 0.92387953251128675613, S = 0.38268343236508977173, R = 0.70710678118654752440;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);     // a[4 k1 + n2]
    // twiddles W16^(n2 k1)
    a[5] = cmul(a[5], cplx{C, -S});                                  // (1,1): W^1
    a[6] = cplx{R * (a[6].x + a[6].y), R * (a[6].y - a[6].x)};       // (k1=1,n2=2): W^2
    a[7] = cmul(a[7], cplx{S, -C});                                  // (1,3): W^3
    a[9] = cplx{R * (a[9].x + a[9].y), R * (a[9].y - a[9].x)};       // (2,1): W^2
    a[10] = cplx{a[10].y, -a[10].x};                                 // (2,2): W^4 = -i
    a[11] = cplx{R * (a[11].y - a[11].x), -R * (a[11].x + a[11].y)}; // (2,3): W^6 = (-R, -R)
    a[13] = cmul(a[13], cplx{S, -C});                                // (3,1): W^3
    a[14] = cplx{R * (a[14].y - a[14].x), -R * (a[14].x + a[14].y)}; // (3,2): W^6
    a[15] = cmul(a[15], cplx{-C, S});                                // (3,3): W^9 = -W^1
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
}

}  // namespace asrk

// ---------------------------------------------------------------------------
// 200-point complex FFT split in two "roles passes" (see spectrogram.cu):
//   Z[k1 + 20 k2] = sum_{n2<10} W200^{n2 k1} ( sum_{n1<20} z[10 n1 + n2] W20^{n1 k1} ) W10^{n2 k2}
// pass 1, role r = n2:  y[k1] = W200^{r k1} * DFT20_{n1}( z[10 n1 + r] )
// pass 2, role j     :  DFT10 over n2 for k1 = j and k1 = 20 - j (role 0: k1 = 0
//                       and 10) -- the two k1 rows that hold each other's mirror
//                       bins, so the real-input split
//   2 X[k]         = (A + B) + O,   A = Z[k], B = conj(Z[200-k]),
//   2 conj(X[200-k]) = (A + B) - O, O = (-i W400^k) (A - B)
//                       runs on registers of one thread.
// ---------------------------------------------------------------------------
namespace asrk {

// tw[k1] = W200^{r k1}
ASRK_HD void fft200_pass1(const cplx (&z)[20], const cplx* tw, cplx (&y)[20]) {
    dft20(z, y);
#pragma unroll
    for (int k1 = 1; k1 < 20; ++k1) y[k1] = cmul(y[k1], tw[k1]);
}

// |X|^2 * 4 for the bin pair (k, 200-k): 2 X[k] = E + O, 2 conj(X[200-k]) = E - O with
// E = A + conj(Zm), O = P (A - conj(Zm)).  14 fp64 instructions: O is never formed, E + O is
// two chained FMAs per component and E - O = 2 E - (E + O) one.
ASRK_HD void split_pair(cplx A, cplx Zm, cplx P, double& pk, double& pm) {
    const double Ex = A.x + Zm.x, Ey = A.y - Zm.y;
    const double Dx = A.x - Zm.x, Dy = A.y + Zm.y;
    const double ex = ASRK_FMA(P.x, Dx, ASRK_FMA(-P.y, Dy, Ex));
    const double ey = ASRK_FMA(P.x, Dy, ASRK_FMA(P.y, Dx, Ey));
    const double fx = ASRK_FMA(2.0, Ex, -ex), fy = ASRK_FMA(2.0, Ey, -ey);
    pk = ASRK_FMA(ex, ex, ey * ey);
    pm = ASRK_FMA(fx, fx, fy * fy);
}

// LoadY(k1, n2) -> cplx ; P[k] = -i W400^k = (-sin(2 pi k/400), -cos(2 pi k/400));
// Emit(k, four_times_power) is called exactly once for every bin k in 0..199
// owned by role j (20 bins per role).
template <class LoadY, class Emit>
ASRK_HD void fft200_pass2(int j, const LoadY& loadY, const cplx* P, const Emit& emit) {
    const int k1a = (j == 0) ? 0 : j;
    const int k1b = (j == 0) ? 10 : 20 - j;
    cplx in[10], za[10], zb[10];
#pragma unroll
    for (int n2 = 0; n2 < 10; ++n2) in[n2] = loadY(k1a, n2);
    dft10(in, za);   // za[k2] = Z[k1a + 20 k2]
#pragma unroll
    for (int n2 = 0; n2 < 10; ++n2) in[n2] = loadY(k1b, n2);
    dft10(in, zb);   // zb[k2] = Z[k1b + 20 k2]
    if (j != 0) {
#pragma unroll
        for (int k2 = 0; k2 < 10; ++k2) {
            const int k = j + 20 * k2;
            double pk, pm;
            split_pair(za[k2], zb[9 - k2], P[k], pk, pm);
            emit(k, pk);
            emit(200 - k, pm);
        }
    } else {
        // row k1 = 0: bins 20 k2, mirror 20 (10 - k2); k2 = 0 and 5 are their own mirror
#pragma unroll
        for (int k2 = 0; k2 <= 5; ++k2) {
            const int k = 20 * k2;
            double pk, pm;
            split_pair(za[k2], za[(10 - k2) % 10], P[k], pk, pm);
            emit(k, pk);
            if (k2 != 0 && k2 != 5) emit(200 - k, pm);
        }
        // row k1 = 10: bins 10 + 20 k2, mirror 10 + 20 (9 - k2)
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) {
            const int k = 10 + 20 * k2;
            double pk, pm;
            split_pair(zb[k2], zb[9 - k2], P[k], pk, pm);
            emit(k, pk);
            emit(200 - k, pm);
        }
    }
}

// ---------------------------------------------------------------------------
// Lane-uniform pass 2: the same code for every role j, for warps whose lanes carry
// different roles (spectrogram.cu: the two half-warps of a warp own two roles).  Role j
// owns rows k1a = j and k1b = 20 - j of Z[k1 + 20 k2] (role 0: the two self-mirrored
// rows 0 and 10), i.e. the bins k = j + 20 s and 200 - k.  Eleven "slots" s = 0..10,
// operands picked with selects, so lanes never diverge:
//   j != 0 : slot s <= 9 pairs za[s] with zb[9-s];                    slot 10 unused
//   j == 0 : slots 0..5 pair za[s] with za[(10-s)%10]   (k = 20 s),
//            slots 6..10 pair zb[s-6] with zb[15-s]     (k = 10 + 20 (s-6))
// Emit(s, pk, pm): 4|X[k]|^2 and 4|X[200-k]|^2 of slot s (s is a constant after
// unrolling).  k = j + 20 s, except for j == 0, s >= 6: k = 20 s - 110.
// ---------------------------------------------------------------------------
ASRK_HD cplx csel(bool c, cplx a, cplx b) { return cplx{c ? a.x : b.x, c ? a.y : b.y}; }

ASRK_HD int lane_k1a(int j) { return j; }
ASRK_HD int lane_k1b(int j) { return j ? 20 - j : 10; }
// bin of slot s for role j
ASRK_HD int lane_bin(int j, int s) { return (j == 0 && s >= 6) ? 20 * s - 110 : j + 20 * s; }

template <class LoadP, class Emit>
ASRK_HD void split_lane(bool j0, const cplx (&za)[10], const cplx (&zb)[10], const LoadP& loadP,
                        const Emit& emit) {
#pragma unroll
    for (int s = 0; s < 11; ++s) {
        cplx A, B;
        if (s <= 5) {
            A = za[s];
            B = csel(j0, za[(10 - s) % 10], zb[9 - s]);
        } else if (s <= 9) {
            A = csel(j0, zb[s - 6], za[s]);
            B = csel(j0, zb[15 - s], zb[9 - s]);
        } else {
            A = zb[4];
            B = zb[5];
        }
        double pk, pm;
        split_pair(A, B, loadP(s), pk, pm);
        emit(s, pk, pm);
    }
}

}  // namespace asrk
