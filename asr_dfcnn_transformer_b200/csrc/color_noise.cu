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
// First correct version (round 1): one kernel launch per radix-2 stage over the whole batch in
// global memory -- HBM-bound at 2 * 16 B * M per utterance and stage; not yet blocked in shared
// memory.
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
    twiddle_kernel<<<(unsigned)((M / 2 + 255) / 256), 256, 0, stream>>>(p.tw, log2M);
    init_kernel<<<gM, 256, 0, stream>>>(p);
    for (int s = 0; s < log2M; ++s) {
        dif_stage_kernel<<<gH, 256, 0, stream>>>(p.A, p.tw, log2M, s);
        dif_stage_kernel<<<gH, 256, 0, stream>>>(p.Bc, p.tw, log2M, s);
    }
    pointwise_kernel<<<gM, 256, 0, stream>>>(p.A, p.Bc, log2M);
    for (int s = 0; s < log2M; ++s) dit_inv_stage_kernel<<<gH, 256, 0, stream>>>(p.A, p.tw, log2M, s);
    shape_kernel<<<gM, 256, 0, stream>>>(p);
    restage_kernel<<<gM, 256, 0, stream>>>(p);
    for (int s = 0; s < log2M; ++s) dif_stage_kernel<<<gH, 256, 0, stream>>>(p.A, p.tw, log2M, s);
    pointwise_kernel<<<gM, 256, 0, stream>>>(p.A, p.Bc, log2M);
    for (int s = 0; s < log2M; ++s) dit_inv_stage_kernel<<<gH, 256, 0, stream>>>(p.A, p.tw, log2M, s);
    double* yreal = p.y;     // the staging buffer is free again: [B][M] doubles fit in its first half
    finish_real_kernel<<<gM, 256, 0, stream>>>(p, yreal);
    mean_max_kernel<<<batch, 1024, 0, stream>>>(p, yreal);
    scale_out_kernel<<<gM, 256, 0, stream>>>(p, yreal);
    return launch_status();
}
