// Mel filterbank front end of the reference's LIVE loaders:
//   compute_fbank_from_api  -- /root/reference/util/wav_util.py:22-31
//     = python_speech_features.logfbank(signal, fs, nfilt=200)  (pre-emphasis 0.97, rectangular
//       400-sample frames at hop 160 zero-padded at the end, 512-point power spectrum / 512,
//       triangular mel filters, log with 0 -> eps)  +  sklearn.preprocessing.scale
//   called at lm_and_am/data_loader.py:129, data_loader2.py:130, end2end/data_loader.py:126.
// One half-warp per frame.  The 512-point real transform is a 256-point complex one
// (z[m] = y[2m] + i y[2m+1]) done as 16 x 16 with register-resident straight-line DFT16 codelets
// in fp64 (the reference is float64 and log() has no +1 floor here, so small bins matter):
//   pass 1, lane r:  DFT16 over m of z[r + 16 m], times W256^(r k1), to a padded exchange tile
//   pass 2, lane k1: DFT16 over r  ->  Z[k1 + 16 k2]
// then the real-input split, power / 512, the sparse triangular mel sums and the log, all from
// shared memory; the per-utterance z-score is a second kernel.  (The first version ran eight
// radix-2 stages through shared memory, one warp per frame: 82 KB of shared-memory traffic per
// frame, 1.11 ms per C2 batch.)
#include <math.h>

#include "asrk_common.cuh"
#include "asrk_fft.cuh"

namespace asrk {
namespace lfb {

constexpr int kNfft = 512, kHalf = 256, kSpec = 257;
constexpr int kWarps = 16;                          // 32 frames in flight per CTA, one CTA per SM
constexpr int kMaxFilt = 224;                       // 14 filters per lane of a half-warp
constexpr int kExch = 16 * 17;                      // padded 16 x 16 exchange tile (complex)
constexpr int kPsDoubles = 264;                     // power spectrum [257]
// per frame: one buffer that is y[512], then the exchange tile, then Z[256]; and the power spectrum
constexpr int kFrameDoubles = 2 * kExch + kPsDoubles;

typedef cplx cd;

struct Params {
    const double* samples;
    const long long* sample_offsets;
    const long long* sample_counts;
    const long long* frame_offsets;
    const long long* out_row_offsets;
    const int* mel_bins;        // [nfilt + 2]
    int batch, nfilt, frame_len, frame_step;
    long long total_frames;
    float* out;
    double preemph;
};

__global__ void __launch_bounds__(kWarps * 32, 1) logfbank_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tw512 = reinterpret_cast<cd*>(smem_raw);                   // W512^k, k <= 256 (real-input split)
    cd* tw256 = tw512 + kSpec + 1;                                 // [k1][r] = W256^(r k1)
    int* bins = reinterpret_cast<int*>(tw256 + 256);               // [kMaxFilt + 2]
    double* w_up = reinterpret_cast<double*>(bins + kMaxFilt + 4); // [260] weight of bin i on the rising edge that covers it
    double* w_dn = w_up + 260;                                     // [260] ... on the falling edge that covers it
    double* frames = w_dn + 260;                                   // [2 kWarps][kFrameDoubles]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < kSpec; k += blockDim.x) {
        double sn, cs;
        sincospi(2.0 * (double)k / (double)kNfft, &sn, &cs);
        tw512[k] = cd{cs, -sn};
    }
    for (int i = tid; i < 256; i += blockDim.x) {
        double sn, cs;
        sincospi(2.0 * (double)((i >> 4) * (i & 15)) / 256.0, &sn, &cs);
        tw256[i] = cd{cs, -sn};
    }
    for (int k = tid; k < p.nfilt + 2; k += blockDim.x) bins[k] = p.mel_bins[k];
    // get_filterbanks: filter j rises over [bin j, bin j+1) as (i - bin j) / (bin j+1 - bin j) and falls over
    // [bin j+1, bin j+2) as (bin j+2 - i) / (bin j+2 - bin j+1); the bins are non-decreasing, so a spectrum
    // bin lies on at most one rising and one falling edge
    for (int j = tid; j < p.nfilt; j += blockDim.x) {
        const int b0 = p.mel_bins[j], b1 = p.mel_bins[j + 1], b2 = p.mel_bins[j + 2];
        for (int i = b0; i < b1 && i < 260; ++i) w_up[i] = (double)(i - b0) / (double)(b1 - b0);
        for (int i = b1; i < b2 && i < 260; ++i) w_dn[i] = (double)(b2 - i) / (double)(b2 - b1);
    }
    __syncthreads();
    const int h = lane >> 4, r = lane & 15;                        // frame slot of the warp, lane of the half-warp
    double* ys = frames + (size_t)(2 * warp + h) * kFrameDoubles;  // y[512] ...
    cd* ex = reinterpret_cast<cd*>(ys);                            // ... then the exchange tile, then Z[256]
    double* ps = ys + 2 * kExch;                                   // power spectrum [257]

    const long long pairs = (p.total_frames + 1) >> 1;
    for (long long gp = (long long)blockIdx.x * kWarps + warp; gp < pairs; gp += (long long)gridDim.x * kWarps) {
        const long long g = 2 * gp + h;
        const bool active = g < p.total_frames;
        int b = 0;
        long long fidx = 0;
        if (active) {
            int lo = 0, hi = p.batch - 1;                          // utterance of the frame
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (p.frame_offsets[mid] <= g) lo = mid; else hi = mid - 1;
            }
            b = lo;
            while (g >= p.frame_offsets[b + 1]) ++b;
            fidx = g - p.frame_offsets[b];
        }
        // pre-emphasised frame (sigproc.preemphasis, then framesig: zeros are appended AFTER the filter):
        // every sample is loaded once, all loads of the frame in flight together; x[s - 1] comes from the
        // neighbouring lane (lane 0: from lane 15 of the previous group of 16)
        {
            const double* x = p.samples + (active ? p.sample_offsets[b] : 0);
            const long long N = active ? p.sample_counts[b] : 0;
            const long long s0 = fidx * p.frame_step;
            double xv[kNfft / 16];
#pragma unroll
            for (int i = 0; i < kNfft / 16; ++i) {
                const int n = r + 16 * i;
                xv[i] = (n < p.frame_len && s0 + n < N) ? x[s0 + n] : 0.0;
            }
            const int groups = (p.frame_len + 15) >> 4;              // groups of 16 samples that hold data
            double carry = (active && s0 > 0 && lane == 16 * h) ? x[s0 - 1] : 0.0;   // sample before the frame
#pragma unroll
            for (int i = 0; i < kNfft / 16; ++i) {
                const int n = r + 16 * i;
                double y = 0.0;
                if (i < groups) {                                     // uniform
                    double prev = __shfl_up_sync(0xffffffffu, xv[i], 1, 16);
                    if (r == 0) prev = carry;
                    carry = __shfl_sync(0xffffffffu, xv[i], 15, 16);
                    if (n < p.frame_len && s0 + n < N) y = (s0 + n == 0) ? xv[i] : xv[i] - p.preemph * prev;
                }
                ys[n] = y;
            }
        }
        __syncwarp();
        cd a[16];
        // pass 1: z[r + 16 m] = (y[2 (r + 16 m)], y[2 (r + 16 m) + 1])
        {
            const cd* zs = reinterpret_cast<const cd*>(ys);
#pragma unroll
            for (int m = 0; m < 16; ++m) a[m] = zs[r + 16 * m];
            __syncwarp();                                         // the buffer becomes the exchange tile
            dft16(a);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const cd v = a[4 * (k1 & 3) + (k1 >> 2)];          // X[k1], k1 = k1lo + 4 k1hi
                ex[r * 17 + k1] = (k1 == 0) ? v : cmul(v, tw256[k1 * 16 + r]);
            }
        }
        __syncwarp();
        // pass 2: lane k1 = r takes column k1 of the tile
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) a[n2] = ex[n2 * 17 + r];
        __syncwarp();                                             // the tile becomes Z[256]
        dft16(a);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) ex[r + 16 * k2] = a[4 * (k2 & 3) + (k2 >> 2)];
        __syncwarp();
        // real-input split, bins k and 256 - k together: with A = Z[k], B = conj(Z[256-k]), E = A + B,
        // O = W512^k (A - B):  X[k] = (E - i O)/2,  conj(X[256-k]) = (E + i O)/2;  power / 512
        for (int k = r; k <= 128; k += 16) {
            const cd A = ex[k], Zm = ex[(256 - k) & 255];
            const cd E = cd{A.x + Zm.x, A.y - Zm.y}, D = cd{A.x - Zm.x, A.y + Zm.y};
            const cd O = cmul(tw512[k], D);
            const double ar = E.x + O.y, ai = E.y - O.x;          // 2 X[k]
            const double br = E.x - O.y, bi = E.y + O.x;          // 2 conj(X[256-k])
            ps[k] = (ar * ar + ai * ai) * (0.25 / kNfft);
            ps[256 - k] = (br * br + bi * bi) * (0.25 / kNfft);
        }
        // mel filters (get_filterbanks): rising edge over [bin j, bin j+1), falling over [bin j+1, bin j+2)
        if (active) {
            const long long row = (p.out_row_offsets ? p.out_row_offsets[b] : p.frame_offsets[b]) + fidx;
            for (int j = r; j < p.nfilt; j += 16) {
                const int b0 = bins[j], b1 = bins[j + 1], b2 = bins[j + 2];
                double feat = 0.0;
                for (int i = b0; i < b1; ++i) feat += ps[i] * w_up[i];
                for (int i = b1; i < b2; ++i) feat += ps[i] * w_dn[i];
                if (feat == 0.0) feat = 2.220446049250313e-16;   // np.finfo(float).eps
                // log(feat) = log(m) + e ln 2 with feat = m 2^e, m in [0.5, 1) split off exactly (the double
                // range of a power spectrum does not fit a float); fp32 lg2.approx of the mantissa: ~2e-7 absolute
                const int hi = __double2hiint(feat);
                int e = ((hi >> 20) & 0x7ff) - 1022;
                double mant = __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(feat));
                if (((hi >> 20) & 0x7ff) == 0) mant = frexp(feat, &e);          // subnormal
                p.out[row * p.nfilt + j] = fmaf((float)e, 0.6931471805599453f, __logf((float)mant));
            }
        }
        __syncwarp();
    }
}

// sklearn.preprocessing.scale per utterance, in place: one CTA per utterance, thread = (row group,
// column): up to five row groups stride through the frames with eight loads in flight each, their
// shifted sums are combined in a fixed order.  A column that is constant (std == 0 in exact
// arithmetic: the empty mel filters) comes out as 0.
constexpr int kZGroups = 5;
__global__ void __launch_bounds__(1024) zscore_kernel(Params p) {
    __shared__ double s_s[kZGroups][kMaxFilt], s_q[kZGroups][kMaxFilt];
    __shared__ double s_mean[kMaxFilt], s_inv[kMaxFilt];
    const int b = blockIdx.x;
    const long long fo = p.frame_offsets[b];
    const long long T = p.frame_offsets[b + 1] - fo;
    if (T <= 0) return;
    const long long row0 = p.out_row_offsets ? p.out_row_offsets[b] : fo;
    float* base = p.out + row0 * p.nfilt;
    int G = (int)blockDim.x / p.nfilt;
    if (G > kZGroups) G = kZGroups;
    const int g = (int)threadIdx.x / p.nfilt, k = (int)threadIdx.x - g * p.nfilt;
    const bool on = g < G;
    double c = 0.0;
    if (on) {
        c = (double)base[k];                                   // shift: the column's first value
        double s = 0.0, q = 0.0;
        for (long long r0 = g; r0 < T; r0 += 8 * G) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (r0 + e * G < T) ? base[(r0 + e * G) * p.nfilt + k] : (float)c;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double d = (double)v[e] - c;
                s += d;
                q = fma(d, d, q);
            }
        }
        s_s[g][k] = s;
        s_q[g][k] = q;
    }
    __syncthreads();
    if (g == 0) {
        double s = 0.0, q = 0.0;
        for (int i = 0; i < G; ++i) { s += s_s[i][k]; q += s_q[i][k]; }      // fixed order
        const double md = s / (double)T;
        double var = q / (double)T - md * md;
        if (var < 0.0) var = 0.0;
        double sd = sqrt(var);
        if (sd < 10.0 * 2.220446049250313e-16) sd = 1.0;
        s_mean[k] = c + md;
        s_inv[k] = 1.0 / sd;
    }
    __syncthreads();
    if (on) {
        const double mean = s_mean[k], inv = s_inv[k];
        for (long long r0 = g; r0 < T; r0 += 8 * G) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (r0 + e * G < T) ? base[(r0 + e * G) * p.nfilt + k] : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (r0 + e * G < T) base[(r0 + e * G) * p.nfilt + k] = (float)(((double)v[e] - mean) * inv);
        }
    }
}

}  // namespace lfb
}  // namespace asrk

using namespace asrk;

extern "C" int asrk_logfbank_run(const double* samples, const long long* sample_offsets,
                                 const long long* sample_counts, const long long* frame_offsets,
                                 const long long* out_row_offsets, const int* mel_bins, int batch,
                                 long long total_frames, int nfilt, int frame_len, int frame_step,
                                 double preemph, int normalise, float* out, asrk_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (batch < 0 || total_frames < 0 || nfilt < 1 || frame_len < 1 || frame_step < 1) return ASRK_E_BADARG;
    if (batch == 0 || total_frames == 0) return ASRK_OK;
    if (!samples || !sample_offsets || !sample_counts || !frame_offsets || !mel_bins || !out) return ASRK_E_BADARG;
    if (nfilt > lfb::kMaxFilt || frame_len > lfb::kNfft) return ASRK_E_SHAPE;
    lfb::Params p;
    p.samples = samples; p.sample_offsets = sample_offsets; p.sample_counts = sample_counts;
    p.frame_offsets = frame_offsets; p.out_row_offsets = out_row_offsets; p.mel_bins = mel_bins;
    p.batch = batch; p.nfilt = nfilt; p.frame_len = frame_len; p.frame_step = frame_step;
    p.total_frames = total_frames; p.out = out; p.preemph = preemph;
    const size_t smem = sizeof(lfb::cd) * (lfb::kSpec + 1 + 256) + sizeof(int) * (lfb::kMaxFilt + 4) +
                        sizeof(double) * 2 * 260 +
                        sizeof(double) * 2 * lfb::kWarps * lfb::kFrameDoubles;
    cudaFuncSetAttribute(lfb::logfbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long blocks = (total_frames + 2 * lfb::kWarps - 1) / (2 * lfb::kWarps);
    const long long cap = (long long)sm_count();
    if (blocks > cap) blocks = cap;
    lfb::logfbank_kernel<<<(unsigned)blocks, lfb::kWarps * 32, smem, stream>>>(p), asrk::note_launch();
    if (normalise) lfb::zscore_kernel<<<batch, 1024, 0, stream>>>(p), asrk::note_launch();
    return launch_status();
}
