// Coloured-noise generation of the reference's augmentation, on the device (SURVEY.md section
// 8a row A4 / 8f row 1):  /root/reference/util/noise.py:17-34  color_noise(len_noise, type_noise)
//     x = N(0,1)[N];  X = fft(x);  half = X[:ceil((N+1)/2)] * k^colour, k = 1..;
//     Hermitian rebuild (even / odd N branches);  y = real(ifft(.));  y -= mean(y);  y /= max(y)
// The normal deviates x are an INPUT (numpy's global Mersenne twister cannot be reproduced on
// the device; the Python surface draws them exactly like the reference does), everything after
// them runs here in fp64.  N is arbitrary (whatever the utterance length is), so the two
// length-N transforms are Bluestein chirp-z transforms on a power-of-two length M >= 2N-1:
//     DFT_N(a)[k] = conj(c_k) * sum_n (a_n conj(c_n)) c_(k-n),   c_n = exp(+i pi n^2 / N)
// i.e. a cyclic convolution of length M done with radix-2 FFTs (decimation in frequency
// forward, decimation in time backward: no bit-reversal pass is ever needed).
// The radix-2 stages are blocked in shared memory (blocked_stages_kernel): a CTA holds 4096
// complex points (64 KB) and runs up to 12 consecutive stages on them -- the stages whose span is
// below 4096 on a contiguous block, the others on a tile of 2^S rows x (4096 >> S) consecutive
// columns -- so a length-2^18 transform crosses HBM twice instead of 18 times.  Same butterflies,
// same twiddle table and same order of operations as the one-launch-per-stage kernels of the first
// version (kept below: they still serve transforms shorter than one block's worth of sense and as
// the reference the blocked kernel was checked against), so the bits are unchanged.
#include <math.h>

#include "asrk_common.cuh"

namespace asrk {
namespace cnoise {

struct cd {
    double x, y;
};
__device__ __forceinline__ cd cmul(cd a, cd b) { return cd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd conj(cd a) { return cd{a.x, -a.y}; }

// c_n = exp(+i pi n^2 / N): the phase is reduced exactly in integers (n^2 mod 2N) first
__device__ __forceinline__ cd chirp(long long n, long long N) {
    const unsigned long long q = ((unsigned long long)n * (unsigned long long)n) % (unsigned long long)(2 * N);
    double s, c;
    sincospi((double)q / (double)N, &s, &c);
    return cd{c, s};
}

struct Params {
    const double* x;                 // [total] normal deviates, ragged
    const long long* offsets;        // [B]
    const long long* counts;         // [B]
    const double* colour;            // [B]
    float* out;                      // [total]
    cd* A;                           // [B][M]
    cd* Bc;                          // [B][M]  FFT of the chirp
    cd* tw;                          // [M/2]   exp(-2 pi i k / M)
    double* y;                       // [B][M]  (first N used)
    double* red;                     // [B][2]  mean, max
    int batch;
    int log2M;
};

__global__ void twiddle_kernel(cd* tw, int log2M) {
    const long long M = 1LL << log2M;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M / 2) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)M, &s, &c);
    tw[k] = cd{c, -s};
}

// A = x conj(c) (zero padded), Bc = chirp wrapped around M
__global__ void init_kernel(Params p) {
    const int b = blockIdx.y;
    const long long M = 1LL << p.log2M, N = p.counts[b];
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= M) return;
    cd a = cd{0.0, 0.0}, bb = cd{0.0, 0.0};
    if (n < N) {
        const cd c = chirp(n, N);
        const double xv = p.x[p.offsets[b] + n];
        a = cd{xv * c.x, -xv * c.y};
        bb = c;
    } else if (M - n < N) {
        bb = chirp(M - n, N);
    }
    p.A[(size_t)b * M + n] = a;
    p.Bc[(size_t)b * M + n] = bb;
}

// one radix-2 decimation-in-frequency stage (natural order in -> bit-reversed order out after the last one)
__global__ void dif_stage_kernel(cd* buf, const cd* tw, int log2M, int stage) {
    const long long M = 1LL << log2M;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M / 2) return;
    cd* v = buf + (size_t)blockIdx.y * M;
    const long long span = M >> (stage + 1);
    const long long j = t & (span - 1);
    const long long i = ((t >> (log2M - stage - 1)) << (log2M - stage)) + j;
    const cd a = v[i], b = v[i + span];
    v[i] = cd{a.x + b.x, a.y + b.y};
    v[i + span] = cmul(cd{a.x - b.x, a.y - b.y}, tw[j << stage]);
}

// one radix-2 decimation-in-time stage of the INVERSE transform (bit-reversed in -> natural out)
__global__ void dit_inv_stage_kernel(cd* buf, const cd* tw, int log2M, int stage) {
    const long long M = 1LL << log2M;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M / 2) return;
    cd* v = buf + (size_t)blockIdx.y * M;
    const int s = log2M - 1 - stage;             // stages run span = 1, 2, 4, ...
    const long long span = M >> (s + 1);
    const long long j = t & (span - 1);
    const long long i = ((t >> (log2M - s - 1)) << (log2M - s)) + j;
    const cd a = v[i], b = cmul(v[i + span], conj(tw[j << s]));
    v[i] = cd{a.x + b.x, a.y + b.y};
    v[i + span] = cd{a.x - b.x, a.y - b.y};
}

// ---------------------------------------------------------------------------
// S consecutive radix-2 stages [s0, s0 + S) of the length-2^L transform on 4096-point tiles in
// shared memory.  Tile element (a, c), a < 2^S, c < C = 4096 >> S, is the global point
//     i = h 2^(L - s0) + a 2^(L - s0 - S) + c0 + c
// (for the last stages, L - s0 - S = 0 and C = 2^(L - s0 - S): the tile is a contiguous block).
// Global stage s pairs points that differ in bit L - 1 - s, i.e. bit (s0 + S - 1 - s) of a.
// INVERSE = false: decimation in frequency, stages in increasing order;
// INVERSE = true : decimation in time of the inverse transform, stages in DEcreasing order.
// ---------------------------------------------------------------------------
constexpr int kTileLog2 = 12, kTile = 1 << kTileLog2;

template <bool INVERSE>
__global__ void __launch_bounds__(256) blocked_stages_kernel(cd* buf, const cd* __restrict__ tw, int L, int s0, int S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* t = reinterpret_cast<cd*>(smem_raw);
    const int low = L - s0 - S;                         // bits of the global index below the active ones
    const int cb = kTileLog2 - S;                       // log2 of the tile's column count (<= low)
    const long long M = 1LL << L;
    cd* v = buf + (size_t)blockIdx.y * M;
    // tile id -> (h, c0)
    const long long tiles_per_h = 1LL << (low - cb);
    const long long h = (long long)blockIdx.x / tiles_per_h;
    const long long c0 = ((long long)blockIdx.x % tiles_per_h) << cb;
    const long long base = (h << (L - s0)) + c0;
    const int C = 1 << cb;
    for (int e = threadIdx.x; e < kTile; e += blockDim.x) {
        const int a = e >> cb, c = e & (C - 1);
        t[e] = v[base + ((long long)a << low) + c];
    }
    __syncthreads();
    // twiddle index of the butterfly whose first point is tile element i0, at global stage s, when the
    // paired bit is bit `abits` of a: the global index below bit L - 1 - s is (low bits of a, column)
    auto tw_of = [&](int i0, int abits, int s) -> cd {
        const long long j = ((long long)((i0 >> cb) & ((1 << abits) - 1)) << low) + c0 + (i0 & (C - 1));
        const double2 w = __ldg(reinterpret_cast<const double2*>(tw) + (j << s));
        return cd{w.x, w.y};
    };
    // two stages per round on four points held in registers (half the shared-memory traffic and
    // barriers of one stage per round); a leftover single stage when S is odd.  k counts stages in
    // processing order: forward s = s0 + k (paired a-bit S-1-k, going down), inverse s = s0+S-1-k
    // (paired a-bit k, going up).
    int k = 0;
    while (k < S) {
        if (S - k >= 2) {
            const int bit_first = INVERSE ? k : (S - 1 - k);              // a-bit paired by the round's first stage
            const int bit_hi = INVERSE ? k + 1 : bit_first, bit_lo = bit_hi - 1;
            const int pb_lo = cb + bit_lo, pb_hi = pb_lo + 1;
            const int s_hi = s0 + S - 1 - bit_hi, s_lo = s_hi + 1;        // global stages that pair bit_hi / bit_lo
#pragma unroll 4
            for (int q = threadIdx.x; q < kTile / 4; q += blockDim.x) {
                const int lo = q & ((1 << pb_lo) - 1);
                const int i00 = ((q >> pb_lo) << (pb_lo + 2)) | lo;
                const int i01 = i00 | (1 << pb_lo), i10 = i00 | (1 << pb_hi), i11 = i10 | (1 << pb_lo);
                cd e0 = t[i00], e1 = t[i01], e2 = t[i10], e3 = t[i11];
                const cd w_lo = tw_of(i00, bit_lo, s_lo);                 // (e0,e1) and (e2,e3)
                const cd w_h0 = tw_of(i00, bit_hi, s_hi);                 // (e0,e2)
                const cd w_h1 = tw_of(i01, bit_hi, s_hi);                 // (e1,e3)
                if (!INVERSE) {
                    cd a = e0, b = e2;
                    e0 = cd{a.x + b.x, a.y + b.y};
                    e2 = cmul(cd{a.x - b.x, a.y - b.y}, w_h0);
                    a = e1; b = e3;
                    e1 = cd{a.x + b.x, a.y + b.y};
                    e3 = cmul(cd{a.x - b.x, a.y - b.y}, w_h1);
                    a = e0; b = e1;
                    e0 = cd{a.x + b.x, a.y + b.y};
                    e1 = cmul(cd{a.x - b.x, a.y - b.y}, w_lo);
                    a = e2; b = e3;
                    e2 = cd{a.x + b.x, a.y + b.y};
                    e3 = cmul(cd{a.x - b.x, a.y - b.y}, w_lo);
                } else {
                    cd a = e0, b = cmul(e1, conj(w_lo));
                    e0 = cd{a.x + b.x, a.y + b.y};
                    e1 = cd{a.x - b.x, a.y - b.y};
                    a = e2; b = cmul(e3, conj(w_lo));
                    e2 = cd{a.x + b.x, a.y + b.y};
                    e3 = cd{a.x - b.x, a.y - b.y};
                    a = e0; b = cmul(e2, conj(w_h0));
                    e0 = cd{a.x + b.x, a.y + b.y};
                    e2 = cd{a.x - b.x, a.y - b.y};
                    a = e1; b = cmul(e3, conj(w_h1));
                    e1 = cd{a.x + b.x, a.y + b.y};
                    e3 = cd{a.x - b.x, a.y - b.y};
                }
                t[i00] = e0; t[i01] = e1; t[i10] = e2; t[i11] = e3;
            }
            k += 2;
        } else {
            const int abits = INVERSE ? k : (S - 1 - k);
            const int s = s0 + S - 1 - abits;
            const int pb = cb + abits;
            for (int q = threadIdx.x; q < kTile / 2; q += blockDim.x) {
                const int lo = q & ((1 << pb) - 1);
                const int i0 = ((q >> pb) << (pb + 1)) | lo, i1 = i0 | (1 << pb);
                const cd w = tw_of(i0, abits, s);
                if (!INVERSE) {
                    const cd a = t[i0], b = t[i1];
                    t[i0] = cd{a.x + b.x, a.y + b.y};
                    t[i1] = cmul(cd{a.x - b.x, a.y - b.y}, w);
                } else {
                    const cd a = t[i0], b = cmul(t[i1], conj(w));
                    t[i0] = cd{a.x + b.x, a.y + b.y};
                    t[i1] = cd{a.x - b.x, a.y - b.y};
                }
            }
            k += 1;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < kTile; e += blockDim.x) {
        const int a = e >> cb, c = e & (C - 1);
        v[base + ((long long)a << low) + c] = t[e];
    }
}

// all L stages of one transform over the batch: blocked when the transform has at least one tile
template <bool INVERSE>
static void run_fft(cd* buf, const cd* tw, int L, int batch, cudaStream_t stream) {
    const long long M = 1LL << L;
    if (L < kTileLog2) {
        const dim3 gH((unsigned)((M / 2 + 255) / 256), batch);
        for (int s = 0; s < L; ++s) {
            if (INVERSE) dit_inv_stage_kernel<<<gH, 256, 0, stream>>>(buf, tw, L, s), asrk::note_launch();
            else dif_stage_kernel<<<gH, 256, 0, stream>>>(buf, tw, L, s), asrk::note_launch();
        }
        return;
    }
    const size_t smem = sizeof(cd) * kTile;
    cudaFuncSetAttribute(blocked_stages_kernel<INVERSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const dim3 grid((unsigned)(M >> kTileLog2), batch);
    // stage groups in forward order: the top L - 12 stages in strided passes of <= 6, then the last 12
    int starts[8], counts[8], n = 0;
    int top = L - kTileLog2;
    for (int s0 = 0; s0 < top;) {
        const int S = (top - s0) < 6 ? (top - s0) : 6;
        starts[n] = s0; counts[n] = S; ++n;
        s0 += S;
    }
    starts[n] = top; counts[n] = kTileLog2; ++n;
    if (!INVERSE) {
        for (int g = 0; g < n; ++g) blocked_stages_kernel<false><<<grid, 256, smem, stream>>>(buf, tw, L, starts[g], counts[g]), asrk::note_launch();
    } else {
        for (int g = n - 1; g >= 0; --g) blocked_stages_kernel<true><<<grid, 256, smem, stream>>>(buf, tw, L, starts[g], counts[g]), asrk::note_launch();
    }
}

__global__ void pointwise_kernel(cd* A, const cd* Bc, int log2M) {
    const long long M = 1LL << log2M;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= M) return;
    const size_t i = (size_t)blockIdx.y * M + n;
    A[i] = cmul(A[i], Bc[i]);
}

// after the first convolution: X[k] = conj(c_k) conv[k] / M.  Spectral shaping of noise.py:19-27 and
// the input of the inverse transform, which is run as conj(DFT(conj Y)): A = conj(Y) conj(c).
__global__ void shape_kernel(Params p) {
    const int b = blockIdx.y;
    const long long M = 1LL << p.log2M, N = p.counts[b];
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    cd* A = p.A + (size_t)b * M;
    // every thread needs conv[k] and conv[N-k]: read both BEFORE anything is overwritten (two-kernel
    // hazard avoided by writing the result to Bc's spare... no: Bc is still needed) -> the result goes to y/imag
    cd out = cd{0.0, 0.0};
    if (k < N) {
        const long long half = N / 2;                          // ceil((N+1)/2) - 1
        const long long src = (k <= half) ? k : N - k;         // upper bins are the conjugates of the lower ones
        const cd c = chirp(src, N);
        const cd v = A[src];
        const double inv = 1.0 / (double)M;
        cd X = cmul(conj(c), cd{v.x * inv, v.y * inv});
        const double g = pow((double)(src + 1), p.colour[b]);
        cd Y = cd{X.x * g, X.y * g};
        if (k > half) Y = conj(Y);
        const cd ck = chirp(k, N);
        out = cmul(conj(Y), conj(ck));
    }
    // staged through y (real) and red-free scratch: write into a second buffer to avoid the read/write race
    reinterpret_cast<cd*>(p.y)[(size_t)b * M + k] = out;
}

// copy the staged input of the second transform into A
__global__ void restage_kernel(Params p) {
    const long long M = 1LL << p.log2M;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const size_t i = (size_t)blockIdx.y * M + k;
    p.A[i] = reinterpret_cast<const cd*>(p.y)[i];
}

// y[m] = Re(conj(c_m) conv2[m]) / (M N), into the real buffer; then mean / max / scaling
__global__ void finish_real_kernel(Params p, double* yreal) {
    const int b = blockIdx.y;
    const long long M = 1LL << p.log2M, N = p.counts[b];
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N) return;
    const cd c = chirp(m, N);
    const cd v = p.A[(size_t)b * M + m];
    yreal[(size_t)b * M + m] = (c.x * v.x + c.y * v.y) / ((double)M * (double)N);
}

__global__ void __launch_bounds__(1024) mean_max_kernel(Params p, const double* yreal) {
    const int b = blockIdx.x;
    const long long M = 1LL << p.log2M, N = p.counts[b];
    const double* y = yreal + (size_t)b * M;
    __shared__ double sh[1024];
    double a = 0.0;
    for (long long i = threadIdx.x; i < N; i += 1024) a += y[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    const double mean = sh[0] / (double)N;
    __syncthreads();
    double mx = -INFINITY;
    for (long long i = threadIdx.x; i < N; i += 1024) mx = fmax(mx, y[i] - mean);
    sh[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) { p.red[2 * b] = mean; p.red[2 * b + 1] = sh[0]; }
}

__global__ void scale_out_kernel(Params p, const double* yreal) {
    const int b = blockIdx.y;
    const long long M = 1LL << p.log2M, N = p.counts[b];
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N) return;
    // noise.py:29-32: (noise - mean) / max(noise - mean), then float32
    p.out[p.offsets[b] + m] = (float)((yreal[(size_t)b * M + m] - p.red[2 * b]) / p.red[2 * b + 1]);
}

struct WsLayout {
    size_t A, Bc, Y, tw, red, total;
};
static WsLayout ws_layout(int batch, int log2M) {
    const size_t M = (size_t)1 << log2M;
    WsLayout l;
    size_t o = 0;
    l.A = o;   o = align_up(o + sizeof(cd) * M * (size_t)batch, 256);
    l.Bc = o;  o = align_up(o + sizeof(cd) * M * (size_t)batch, 256);
    l.Y = o;   o = align_up(o + sizeof(cd) * M * (size_t)batch, 256);
    l.tw = o;  o = align_up(o + sizeof(cd) * (M / 2), 256);
    l.red = o; o = align_up(o + sizeof(double) * 2 * (size_t)batch, 256);
    l.total = o;
    return l;
}

static int log2_fft_len(long long max_count) {
    int l = 1;
    while ((1LL << l) < 2 * max_count - 1) ++l;
    return l;
}

}  // namespace cnoise
}  // namespace asrk

using namespace asrk;
using namespace asrk::cnoise;

extern "C" size_t asrk_color_noise_workspace_bytes(int batch, long long max_count) {
    if (batch <= 0 || max_count <= 0) return 0;
    return ws_layout(batch, log2_fft_len(max_count)).total;
}

extern "C" int asrk_color_noise_run(const double* normals, const long long* offsets, const long long* counts,
                                    const double* colour, int batch, long long max_count, float* out,
                                    void* workspace, size_t workspace_bytes, asrk_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (batch < 0 || max_count < 0) return ASRK_E_BADARG;
    if (batch == 0 || max_count == 0) return ASRK_OK;
    if (!normals || !offsets || !counts || !colour || !out || !workspace) return ASRK_E_BADARG;
    if (max_count > (1LL << 27)) return ASRK_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return ASRK_E_WORKSPACE;
    const int log2M = log2_fft_len(max_count);
    const WsLayout l = ws_layout(batch, log2M);
    if (workspace_bytes < l.total) return ASRK_E_WORKSPACE;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    Params p;
    p.x = normals; p.offsets = offsets; p.counts = counts; p.colour = colour; p.out = out;
    p.A = reinterpret_cast<cd*>(ws + l.A);
    p.Bc = reinterpret_cast<cd*>(ws + l.Bc);
    p.y = reinterpret_cast<double*>(ws + l.Y);
    p.tw = reinterpret_cast<cd*>(ws + l.tw);
    p.red = reinterpret_cast<double*>(ws + l.red);
    p.batch = batch; p.log2M = log2M;
    const long long M = 1LL << log2M;
    const dim3 gM((unsigned)((M + 255) / 256), batch), gH((unsigned)((M / 2 + 255) / 256), batch);
    twiddle_kernel<<<(unsigned)((M / 2 + 255) / 256), 256, 0, stream>>>(p.tw, log2M), asrk::note_launch();
    init_kernel<<<gM, 256, 0, stream>>>(p), asrk::note_launch();
    run_fft<false>(p.A, p.tw, log2M, batch, stream);
    run_fft<false>(p.Bc, p.tw, log2M, batch, stream);
    pointwise_kernel<<<gM, 256, 0, stream>>>(p.A, p.Bc, log2M), asrk::note_launch();
    run_fft<true>(p.A, p.tw, log2M, batch, stream);
    shape_kernel<<<gM, 256, 0, stream>>>(p), asrk::note_launch();
    restage_kernel<<<gM, 256, 0, stream>>>(p), asrk::note_launch();
    run_fft<false>(p.A, p.tw, log2M, batch, stream);
    pointwise_kernel<<<gM, 256, 0, stream>>>(p.A, p.Bc, log2M), asrk::note_launch();
    run_fft<true>(p.A, p.tw, log2M, batch, stream);
    double* yreal = p.y;     // the staging buffer is free again: [B][M] doubles fit in its first half
    finish_real_kernel<<<gM, 256, 0, stream>>>(p, yreal), asrk::note_launch();
    mean_max_kernel<<<batch, 1024, 0, stream>>>(p, yreal), asrk::note_launch();
    scale_out_kernel<<<gM, 256, 0, stream>>>(p, yreal), asrk::note_launch();
    return launch_status();
}
