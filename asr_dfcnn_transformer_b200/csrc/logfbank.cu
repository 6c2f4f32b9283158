// Mel filterbank front end of the reference's LIVE loaders:
//   compute_fbank_from_api  -- /root/reference/util/wav_util.py:22-31
//     = python_speech_features.logfbank(signal, fs, nfilt=200)  (pre-emphasis 0.97, rectangular
//       400-sample frames at hop 160 zero-padded at the end, 512-point power spectrum / 512,
//       triangular mel filters, log with 0 -> eps)  +  sklearn.preprocessing.scale
//   called at lm_and_am/data_loader.py:129, data_loader2.py:130, end2end/data_loader.py:126.
// First correct version (round 1): one warp per frame, a 256-point complex radix-2 FFT in the
// warp's shared-memory buffer in fp64 (the reference is float64 and log() has no +1 floor here,
// so small bins matter), real-input split, sparse mel sums, log; the per-utterance z-score is a
// second kernel.  Not yet tuned like spectrogram.cu (no register-resident codelets).
#include <math.h>

#include "asrk_common.cuh"

namespace asrk {
namespace lfb {

constexpr int kNfft = 512, kHalf = 256, kSpec = 257;
constexpr int kWarps = 8;
constexpr int kMaxFilt = 224;                       // 7 filters per lane

struct cd {
    double x, y;
};

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

__device__ __forceinline__ cd cmul(cd a, cd b) { return cd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

__global__ void __launch_bounds__(kWarps * 32) logfbank_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tw = reinterpret_cast<cd*>(smem_raw);                     // W512^k, k <= 256
    int* bins = reinterpret_cast<int*>(tw + kSpec);               // [kMaxFilt + 2]
    double* inv_up = reinterpret_cast<double*>(bins + kMaxFilt + 4);   // [kMaxFilt] 1 / (bin j+1 - bin j)
    double* inv_dn = inv_up + kMaxFilt;                            // [kMaxFilt] 1 / (bin j+2 - bin j+1)
    cd* zbuf = reinterpret_cast<cd*>(inv_dn + kMaxFilt);           // [kWarps][256]
    double* pbuf = reinterpret_cast<double*>(zbuf + kWarps * kHalf);   // [kWarps][260]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < kSpec; k += blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)k / (double)kNfft, &s, &c);
        tw[k] = cd{c, -s};
    }
    for (int k = tid; k < p.nfilt + 2; k += blockDim.x) bins[k] = p.mel_bins[k];
    for (int j = tid; j < p.nfilt; j += blockDim.x) {
        const int d1 = p.mel_bins[j + 1] - p.mel_bins[j], d2 = p.mel_bins[j + 2] - p.mel_bins[j + 1];
        inv_up[j] = d1 > 0 ? 1.0 / (double)d1 : 0.0;
        inv_dn[j] = d2 > 0 ? 1.0 / (double)d2 : 0.0;
    }
    __syncthreads();
    cd* z = zbuf + warp * kHalf;
    double* ps = pbuf + warp * 260;

    for (long long g = (long long)blockIdx.x * kWarps + warp; g < p.total_frames; g += (long long)gridDim.x * kWarps) {
        // utterance of the frame
        int lo = 0, hi = p.batch - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.frame_offsets[mid] <= g) lo = mid; else hi = mid - 1;
        }
        int b = lo;
        while (g >= p.frame_offsets[b + 1]) ++b;
        const long long fidx = g - p.frame_offsets[b];
        const double* x = p.samples + p.sample_offsets[b];
        const long long N = p.sample_counts[b];
        const long long s0 = fidx * p.frame_step;
        // pre-emphasised frame (sigproc.preemphasis then framesig: zeros are appended AFTER the filter),
        // packed as z[m] = y[2m] + i y[2m+1], bit-reversed on the way in
        for (int m = lane; m < kHalf; m += 32) {
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = 2 * m + e;
                const long long s = s0 + n;
                double y = 0.0;
                if (n < p.frame_len && s < N) y = (s == 0) ? x[0] : x[s] - p.preemph * x[s - 1];
                v[e] = y;
            }
            z[__brev((unsigned)m) >> 24] = cd{v[0], v[1]};
        }
        __syncwarp();
        // 256-point radix-2 DIT, 8 stages, 128 butterflies each
#pragma unroll 1
        for (int st = 1; st <= 8; ++st) {
            const int half = 1 << (st - 1);
            const int tstep = kNfft >> st;                 // W256^(j * 128/half) = W512^(j * 256/half)
            // four butterflies per lane: all loads first, so that their latencies overlap
            cd a[4], b[4], w[4];
            int idx[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int q = lane + 32 * e;
                const int j = q & (half - 1);
                idx[e] = ((q >> (st - 1)) << st) + j;
                a[e] = z[idx[e]];
                b[e] = z[idx[e] + half];
                w[e] = tw[j * tstep];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const cd bb = cmul(b[e], w[e]);
                z[idx[e]] = cd{a[e].x + bb.x, a[e].y + bb.y};
                z[idx[e] + half] = cd{a[e].x - bb.x, a[e].y - bb.y};
            }
            __syncwarp();
        }
        // real-input split: X[k] = (A + B)/2 - (i/2) W512^k (A - B), A = Z[k], B = conj(Z[256-k]); power / 512
        for (int k = lane; k < kSpec; k += 32) {
            const cd A = z[k & 255], Zm = z[(256 - k) & 255];
            const cd B = cd{Zm.x, -Zm.y};
            const cd E = cd{A.x + B.x, A.y + B.y}, D = cd{A.x - B.x, A.y - B.y};
            const cd O = cmul(tw[k], D);                     // W512^k D; -i O = (O.y, -O.x)
            const double xr = 0.5 * (E.x + O.y), xi = 0.5 * (E.y - O.x);
            ps[k] = (xr * xr + xi * xi) * (1.0 / kNfft);
        }
        __syncwarp();
        // mel filters (get_filterbanks): rising edge over [bin j, bin j+1), falling over [bin j+1, bin j+2)
        const long long row = (p.out_row_offsets ? p.out_row_offsets[b] : p.frame_offsets[b]) + fidx;
        for (int j = lane; j < p.nfilt; j += 32) {
            const int b0 = bins[j], b1 = bins[j + 1], b2 = bins[j + 2];
            const double iu = inv_up[j], id = inv_dn[j];
            double feat = 0.0;
            for (int i = b0; i < b1; ++i) feat += ps[i] * ((double)(i - b0) * iu);
            for (int i = b1; i < b2; ++i) feat += ps[i] * ((double)(b2 - i) * id);
            if (feat == 0.0) feat = 2.220446049250313e-16;   // np.finfo(float).eps
            // log(feat) = log(m) + e ln 2 with feat = m 2^e, m in [0.5, 1): fp32 log of the mantissa (the
            // double range of a power spectrum does not fit a float), error ~1e-7 absolute
            int e;
            const double mant = frexp(feat, &e);
            p.out[row * p.nfilt + j] = logf((float)mant) + (float)e * 0.6931471805599453f;
        }
        __syncwarp();
    }
}

// sklearn.preprocessing.scale per utterance, in place: one CTA per utterance, thread = column.
// A column that is constant (std == 0 in exact arithmetic: the empty mel filters) comes out as 0.
__global__ void __launch_bounds__(256) zscore_kernel(Params p) {
    const int b = blockIdx.x;
    const long long fo = p.frame_offsets[b];
    const long long T = p.frame_offsets[b + 1] - fo;
    const long long row0 = p.out_row_offsets ? p.out_row_offsets[b] : fo;
    float* base = p.out + row0 * p.nfilt;
    for (int k = threadIdx.x; k < p.nfilt; k += blockDim.x) {
        if (T <= 0) continue;
        const double c = (double)base[k];
        double s = 0.0, q = 0.0;
        for (long long r0 = 0; r0 < T; r0 += 8) {            // eight loads in flight
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (r0 + e < T) ? base[(r0 + e) * p.nfilt + k] : (float)c;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double d = (double)v[e] - c;
                s += d;
                q = fma(d, d, q);
            }
        }
        const double md = s / (double)T;
        double var = q / (double)T - md * md;
        if (var < 0.0) var = 0.0;
        double sd = sqrt(var);
        if (sd < 10.0 * 2.220446049250313e-16) sd = 1.0;
        const double mean = c + md, inv = 1.0 / sd;
        for (long long r0 = 0; r0 < T; r0 += 8) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (r0 + e < T) ? base[(r0 + e) * p.nfilt + k] : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (r0 + e < T) base[(r0 + e) * p.nfilt + k] = (float)(((double)v[e] - mean) * inv);
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
    const size_t smem = sizeof(lfb::cd) * lfb::kSpec + sizeof(int) * (lfb::kMaxFilt + 4) +
                        sizeof(double) * 2 * lfb::kMaxFilt +
                        sizeof(lfb::cd) * lfb::kWarps * lfb::kHalf + sizeof(double) * lfb::kWarps * 260;
    cudaFuncSetAttribute(lfb::logfbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long blocks = (total_frames + lfb::kWarps - 1) / lfb::kWarps;
    const long long cap = (long long)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    lfb::logfbank_kernel<<<(unsigned)blocks, lfb::kWarps * 32, smem, stream>>>(p);
    if (normalise) lfb::zscore_kernel<<<batch, 256, 0, stream>>>(p);
    return launch_status();
}
