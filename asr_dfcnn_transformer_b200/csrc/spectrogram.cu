// Part 1 of the hot path: spectrogram features (+ fused noise mix, + z-score).
//
// Reference behaviour reproduced: /root/reference/util/wav_util.py:49-79
// (compute_fbank), :82-112 (compute_fbank_from_asrt), util/noise.py:48-52,108.
//
// B200 design (DESIGN.md "spectrogram kernel"):
//   * the reference computes in float64 and the tolerance (1e-4 on log|X|, then
//     amplified by the z-score) is not reachable with fp32 butterflies on
//     high-dynamic-range audio, so the 400-point transform runs in fp64 -- B200
//     has a half-rate fp64 pipe (64 lanes/SM/clk); everything after |X|^2
//     (sqrt, log, the output) is fp32.
//   * one persistent CTA per SM; a tile is 32 consecutive frames of one
//     utterance.  LANE = FRAME, WARP = ROLE: ten "FFT warps" each own one of the
//     ten residues of the 200 = 20 x 10 Cooley-Tukey split, so window values and
//     twiddles are warp-uniform (shared-memory broadcasts), the PCM tile and the
//     32x200 output tile are staged in padded shared memory, and the only
//     exchange between the two passes is one conflict-free 100 KB fp64 buffer.
//   * four "helper warps" run concurrently: they stage the next PCM tile with
//     16-byte loads (mixing in K*noise on the fly), and turn the previous tile's
//     |X|^2 into log(|X|+1), store it coalesced, and accumulate the per-bin
//     column sums for the z-score in fp64.  Column sums are combined in a fixed
//     order (no float atomics) and the CTA that retires the last tile of an
//     utterance publishes mean / 1/std; a second, purely streaming kernel
//     normalises in place.
#include <math.h>

#include "asrk_common.cuh"
#include "asrk_fft.cuh"

namespace asrk {
namespace spec {

constexpr int kFftWarps = 10;
constexpr int kHelperWarps = 4;
constexpr int kFftThreads = kFftWarps * 32;          // 320
constexpr int kHelperThreads = kHelperWarps * 32;    // 128
constexpr int kThreads = kFftThreads + kHelperThreads;
constexpr int kTile = 32;                            // frames per tile (= lanes)
constexpr int kHop = 160, kFrameLen = 400, kBins = 200;
constexpr int kHopRows = 34;                         // 31*160+400 = 5360 samples -> 34 hops
constexpr int kOutStride = 201;                      // padded row of the |X|^2 tile
// hop rows stay 16-byte aligned (for 16-byte async copies) and are padded by 16
// bytes: lane f reads word 84 f + c -> 8 distinct banks, a 4-way conflict on the
// 20 sample loads of a thread per tile (the rest of its ~110 shared accesses are
// conflict-free), instead of the 16-way conflict of the unpadded layout.
constexpr int kHopWordsI16 = 84;                     // 80 words of int16 pairs + 4 pad
constexpr int kHopWordsF32 = 164;                    // 160 words + 4 pad
constexpr int kTabDoubles = 1200;                    // window[400] | W200 table[200 cplx] | P[200 cplx]

struct TileRec {            // written by the main kernel, read by the normalise kernel
    int b;
    int nf;
    long long row0;
};

struct Meta {
    int b;
    int f0;
    int nf;
    int tiles_b;             // tiles of utterance b
    long long sbase;         // first sample of the utterance in the ragged buffer
    long long nsamp;         // samples of the utterance
    long long row0;          // output row of frame f0
    long long nfr;           // frames of the utterance
    float mag;
    float gain;
};

struct Params {
    const void* samples;
    const float* noise;
    const float* gain;
    const long long* sample_offsets;
    const long long* sample_counts;
    const long long* frame_offsets;
    const long long* out_row_offsets;
    int batch;
    int mode;
    float* out;
    // workspace
    const double* tables;
    int* tile_offsets;       // [B+1]
    int* done;               // [B]
    double* stats;           // [B][400]: mean[200], inv_std[200]
    double* partials;        // [G+B][400]
    TileRec* tile_rec;       // [n_tiles]
};

struct WsLayout {
    size_t tables, tile_offsets, done, gains, stats, partials, tile_rec, total;
};

static WsLayout ws_layout(int batch, long long total_frames, int grid) {
    WsLayout l;
    size_t o = 0;
    l.tables = o;        o = align_up(o + sizeof(double) * kTabDoubles, 256);
    l.tile_offsets = o;  o = align_up(o + sizeof(int) * (size_t)(batch + 1), 256);
    l.done = o;          o = align_up(o + sizeof(int) * (size_t)batch, 256);
    l.gains = o;         o = align_up(o + sizeof(float) * (size_t)batch, 256);
    l.stats = o;         o = align_up(o + sizeof(double) * 400 * (size_t)batch, 256);
    l.partials = o;      o = align_up(o + sizeof(double) * 400 * kHelperWarps * (size_t)(grid + batch + 1), 256);
    l.tile_rec = o;
    size_t max_tiles = (size_t)(total_frames / kTile) + (size_t)batch + 1;
    o = align_up(o + sizeof(TileRec) * max_tiles, 256);
    l.total = o;
    return l;
}

// ---------------------------------------------------------------------------
// setup: constant tables, tile prefix sums, counters
// ---------------------------------------------------------------------------
__global__ void setup_kernel(double* tables, const long long* frame_offsets, int batch,
                             int* tile_offsets, int* done) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 400; i += nt) {
        // wav_util.py:51-52  w = 0.54 - 0.46 cos(2 pi x / 399)
        tables[i] = 0.54 - 0.46 * cospi(2.0 * (double)i / 399.0);
    }
    for (int i = tid; i < 200; i += nt) {
        // pass-1 twiddles, [r][k1]: W200^{r k1}
        const int r = i / 20, k1 = i % 20;
        double s, c;
        sincospi(2.0 * (double)(r * k1) / 200.0, &s, &c);
        tables[400 + 2 * i] = c;
        tables[400 + 2 * i + 1] = -s;
        // split twiddles P[k] = -i W400^k
        sincospi(2.0 * (double)i / 400.0, &s, &c);
        tables[800 + 2 * i] = -s;
        tables[800 + 2 * i + 1] = -c;
    }
    for (int i = tid; i < batch; i += nt) done[i] = 0;
    // exclusive scan of ceil(n_frames / 32); B is small (hundreds..thousands)
    __shared__ int carry;
    __shared__ int scan[1024];
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < batch; base += nt) {
        const int i = base + tid;
        int v = 0;
        if (i < batch) {
            long long n = frame_offsets[i + 1] - frame_offsets[i];
            if (n < 0) n = 0;
            v = (int)((n + kTile - 1) / kTile);
        }
        scan[tid] = v;
        __syncthreads();
        for (int o = 1; o < nt; o <<= 1) {
            int add = (tid >= o) ? scan[tid - o] : 0;
            __syncthreads();
            scan[tid] += add;
            __syncthreads();
        }
        if (i < batch) tile_offsets[i] = carry + scan[tid] - v;
        __syncthreads();
        if (tid == nt - 1) carry += scan[tid];
        __syncthreads();
    }
    if (tid == 0) tile_offsets[batch] = carry;
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double i16_to_f64(int v) {
    // exact int -> double with one integer op and one DADD (the I2F.F64 path is
    // a quarter-rate conversion): 2^52 + 2^31 + v has v in its low mantissa bits.
    return __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)v)) - 4503601774854144.0;
}


template <bool F32>
__device__ __forceinline__ void load_pcm_tile(const Params& p, const Meta& m, uint32_t* dst, int hth) {
    const long long t0 = (long long)m.f0 * kHop;   // utterance-local first sample of the tile
    if (!F32) {
        const short* src = reinterpret_cast<const short*>(p.samples) + m.sbase;
        constexpr int kChunks = kHopRows * 20;   // 16-byte chunks of 8 samples
        for (int c = hth; c < kChunks; c += kHelperThreads) {
            const long long us = t0 + (long long)c * 8;
            uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            const short* g = src + us;
            if (us + 8 <= m.nsamp && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
                const int4 v = __ldg(reinterpret_cast<const int4*>(g));
                w0 = (uint32_t)v.x; w1 = (uint32_t)v.y; w2 = (uint32_t)v.z; w3 = (uint32_t)v.w;
            } else {
                uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (us + e < m.nsamp) {
                        const uint32_t s = (uint16_t)g[e];
                        w[e >> 1] |= s << (16 * (e & 1));
                    }
                }
                w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
            }
            uint32_t* d = dst + (c / 20) * kHopWordsI16 + (c % 20) * 4;
            d[0] = w0; d[1] = w1; d[2] = w2; d[3] = w3;
        }
    } else {
        const float* src = reinterpret_cast<const float*>(p.samples) + m.sbase;
        const float* nz = p.noise ? p.noise + m.sbase : nullptr;
        const float K = m.gain;
        constexpr int kChunks = kHopRows * 40;   // 16-byte chunks of 4 samples
        for (int c = hth; c < kChunks; c += kHelperThreads) {
            const long long us = t0 + (long long)c * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            const float* g = src + us;
            if (us + 4 <= m.nsamp && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(g));
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                if (nz) {
                    const float4 n4 = __ldg(reinterpret_cast<const float4*>(nz + us));
                    // noise.py:108  (signal + K * noise).astype(float32): two roundings, no FMA
                    v[0] = __fadd_rn(v[0], __fmul_rn(K, n4.x));
                    v[1] = __fadd_rn(v[1], __fmul_rn(K, n4.y));
                    v[2] = __fadd_rn(v[2], __fmul_rn(K, n4.z));
                    v[3] = __fadd_rn(v[3], __fmul_rn(K, n4.w));
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (us + e < m.nsamp) {
                        float s = g[e];
                        if (nz) s = __fadd_rn(s, __fmul_rn(K, nz[us + e]));
                        v[e] = s;
                    }
                }
            }
            float* d = reinterpret_cast<float*>(dst) + (c / 40) * kHopWordsF32 + (c % 40) * 4;
            *reinterpret_cast<float2*>(d) = make_float2(v[0], v[1]);
            *reinterpret_cast<float2*>(d + 2) = make_float2(v[2], v[3]);
        }
    }
}

// Locate tile `tile`: b such that tile_offsets[b] <= tile < tile_offsets[b+1].
__device__ void fill_meta(const Params& p, int tile, int hint_b, Meta& m) {
    int b;
    if (hint_b >= 0) {
        b = hint_b;
        while (tile >= p.tile_offsets[b + 1]) ++b;
    } else {
        int lo = 0, hi = p.batch - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.tile_offsets[mid] <= tile) lo = mid; else hi = mid - 1;
        }
        b = lo;
        while (tile >= p.tile_offsets[b + 1]) ++b;   // skip utterances without tiles
    }
    const int t0 = p.tile_offsets[b];
    m.b = b;
    m.tiles_b = p.tile_offsets[b + 1] - t0;
    m.f0 = (tile - t0) * kTile;
    const long long fo = p.frame_offsets[b];
    m.nfr = p.frame_offsets[b + 1] - fo;
    const long long rem = m.nfr - m.f0;
    m.nf = rem < kTile ? (int)rem : kTile;
    m.sbase = p.sample_offsets[b];
    m.nsamp = p.sample_counts[b];
    m.row0 = (p.out_row_offsets ? p.out_row_offsets[b] : fo) + m.f0;
    m.mag = (p.mode == ASRK_SPEC_ASRT) ? (1.0f / (float)m.nsamp) : 1.0f;
    m.gain = (p.noise && p.gain) ? p.gain[b] : 0.0f;
}

// owner CTA of a tile under the contiguous chunking start_c = c * n / G
__device__ __forceinline__ int chunk_of(long long tile, long long n, long long G) {
    return (int)(((tile + 1) * G - 1) / n);
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
// Named barriers (id 0 is __syncthreads).  FULL/EMPTY pairs hand the three PCM
// buffers and the two |X|^2 tiles between the FFT warps and the helper warps so
// that neither side waits for the other unless it is a whole tile behind.
enum : int {
    kBarHelpers = 1,      // helper warps only
    kBarExchA = 2,        // FFT warps only: pass 1 -> pass 2
    kBarExchB = 3,        // FFT warps only: pass 2 -> next pass 1
    kBarPcmFull = 4,      // +stage (4,5,6)
    kBarPcmEmpty = 7,     // +stage (7,8,9)
    kBarOutFull = 10,     // +slot (10,11)
    kBarOutEmpty = 12,    // +slot (12,13)
};
constexpr int kPipeThreads = kFftThreads + kHelperThreads;   // threads on a FULL/EMPTY barrier

__device__ __forceinline__ void bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int n) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// log(|X| * mag + 1) from p4 = 4 |X|^2, natural log (wav_util.py:76,107,111), fp32:
// v = 1 + m is split into 2^e * f exactly; lg2.approx on f in [1,2) has an absolute
// error of 2^-22, so the result is good to ~1 ulp at any magnitude; for small m the
// rounding of 1 + m is put back with the first-order term (m - (v - 1)) / v.
__device__ __forceinline__ float log_mag(float p4, float half_mag) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p4));
    const float m = r * half_mag;
    const float v = 1.0f + m;
    const int vi = __float_as_int(v);
    const float e = (float)((vi >> 23) - 127);
    const float f = __int_as_float((vi & 0x007fffff) | 0x3f800000);
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(f));
    // ln2 = 0.693145751953125 (exact in 12 bits) + 1.42860682e-6
    float y = fmaf(e, 0.693145751953125f, fmaf(l2, 0.69314718056f, e * 1.42860682e-6f));
    // (called from divergent code: no warp votes here)
    if (m < 0.5f) y += __fdividef(m - (v - 1.0f), v);
    return y;
}

// Asynchronous staging of one PCM tile (no arithmetic on the way): 16-byte LDGSTS
// copies into the padded hop rows, zero-filled past the end of the utterance.
// Needs the utterance start to be 16-byte aligned.
template <bool F32>
__device__ __forceinline__ void issue_pcm_tile_async(const Params& p, const Meta& m, uint32_t* dst, int hth) {
    const long long t0 = (long long)m.f0 * kHop;
    constexpr int kPerChunk = F32 ? 4 : 8;             // samples per 16 bytes
    constexpr int kChunksPerHop = kHop / kPerChunk;    // 40 / 20
    constexpr int kChunks = kHopRows * kChunksPerHop;
    constexpr int kBytes = F32 ? 4 : 2;
    constexpr int kRowWords = F32 ? kHopWordsF32 : kHopWordsI16;
    const char* src = reinterpret_cast<const char*>(p.samples) + m.sbase * kBytes;
    for (int c = hth; c < kChunks; c += kHelperThreads) {
        const long long us = t0 + (long long)c * kPerChunk;
        const long long rem = (m.nsamp - us) * kBytes;
        const int nb = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
        cp_async16(dst + (c / kChunksPerHop) * kRowWords + (c % kChunksPerHop) * 4,
                   nb ? src + us * kBytes : src, nb);
    }
}

template <bool F32>
__global__ void __launch_bounds__(kThreads, 1) spectrogram_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kPcmWords = kHopRows * (F32 ? kHopWordsF32 : kHopWordsI16);
    double* tab = reinterpret_cast<double*>(smem_raw);                       // 1200 doubles
    cplx* exch = reinterpret_cast<cplx*>(tab + kTabDoubles);                  // [200][32]
    float* outt = reinterpret_cast<float*>(exch + 200 * kTile);               // [2][32*201]
    uint32_t* pcm = reinterpret_cast<uint32_t*>(outt + 2 * kTile * kOutStride);  // [3][kPcmWords]
    __shared__ Meta meta[4];
    __shared__ int s_flag;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const long long n_tiles = p.tile_offsets[p.batch];
    const long long G = gridDim.x;
    const int tile_begin = (int)((long long)blockIdx.x * n_tiles / G);
    const int tile_end = (int)((long long)(blockIdx.x + 1) * n_tiles / G);
    const int my_tiles = tile_end - tile_begin;
    if (my_tiles <= 0) return;   // uniform per CTA

    for (int i = tid; i < kTabDoubles; i += kThreads) tab[i] = p.tables[i];
    __syncthreads();

    const double* tabW = tab;
    const cplx* tabT = reinterpret_cast<const cplx*>(tab + 400);
    const cplx* tabP = reinterpret_cast<const cplx*>(tab + 800);

    // warp roles: 0..9 FFT (one residue of the 20 x 10 split each); 10..13 helpers
    if (warp < kFftWarps) {
        // ------------------------------ FFT warps ------------------------------
        const int r = warp;
        int st = 0;   // PCM stage = i % 3
        for (int i = 0; i < my_tiles; ++i) {
            const int s = i & 1;
            const uint32_t* pc = pcm + st * kPcmWords;
            bar_sync(kBarPcmFull + st, kPipeThreads);
            {
                cplx z[20], y[20];
#pragma unroll
                for (int n1 = 0; n1 < 20; ++n1) {
                    const int q = n1 / 8;                 // hop row offset of sample 2*(10 n1 + r)
                    const int wq = 10 * (n1 % 8) + r;     // int16-pair index inside the hop
                    double x0, x1;
                    if (!F32) {
                        const uint32_t w = pc[(lane + q) * kHopWordsI16 + wq];
                        x0 = i16_to_f64((int)(short)(w & 0xffffu));
                        x1 = i16_to_f64((int)(short)(w >> 16));
                    } else {
                        const float2 v = *reinterpret_cast<const float2*>(
                            reinterpret_cast<const float*>(pc) + (lane + q) * kHopWordsF32 + 2 * wq);
                        x0 = (double)v.x;
                        x1 = (double)v.y;
                    }
                    const double2 w2 = *reinterpret_cast<const double2*>(tabW + 20 * n1 + 2 * r);
                    z[n1] = cplx{x0 * w2.x, x1 * w2.y};   // wav_util.py:71 data_line * w
                }
                if (i + 3 < my_tiles) bar_arrive(kBarPcmEmpty + st, kPipeThreads);
                fft200_pass1(z, tabT + r * 20, y);
#pragma unroll
                for (int k1 = 0; k1 < 20; ++k1) exch[(k1 * 10 + r) * kTile + lane] = y[k1];
            }
            bar_sync(kBarExchA, kFftThreads);
            if (i >= 2) bar_sync(kBarOutEmpty + s, kPipeThreads);
            {
                float* ot = outt + s * (kTile * kOutStride) + lane * kOutStride;
                const float half_mag = 0.5f * meta[i & 3].mag;
                auto loadY = [&](int k1, int n2) { return exch[(k1 * 10 + n2) * kTile + lane]; };
                auto emit = [&](int k, double p4) { ot[k] = log_mag((float)p4, half_mag); };
                fft200_pass2(r, loadY, tabP, emit);
            }
            bar_arrive(kBarOutFull + s, kPipeThreads);
            bar_sync(kBarExchB, kFftThreads);
            st = (st == 2) ? 0 : st + 1;
        }
    } else {
        // ----------------------------- helper warps ----------------------------
        const int hw = warp - kFftWarps;                          // 0..3
        const int hth = hw * 32 + lane;
        const bool want_stats = (p.mode == ASRK_SPEC_FBANK);
        const bool mix = (p.noise != nullptr);
        double acc1[7], acc2[7];
#pragma unroll
        for (int s = 0; s < 7; ++s) { acc1[s] = 0.0; acc2[s] = 0.0; }
        int acc_b = -1;        // utterance the accumulators belong to
        int acc_tiles = 0;     // tiles accumulated since the last flush
        int acc_tiles_b = 0;   // total tiles of utterance acc_b
        long long acc_nfr = 0;

        auto flush = [&]() {
            // publish this warp's partial column sums; the CTA that retires the last
            // tile of the utterance adds all partials in a fixed order (no atomics on
            // data) and writes mean and 1/std
            double* slot = p.partials + ((size_t)(blockIdx.x + acc_b) * kHelperWarps + hw) * 400;
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                const int k = lane + 32 * s;
                if (k < kBins) { slot[k] = acc1[s]; slot[200 + k] = acc2[s]; }
                acc1[s] = 0.0; acc2[s] = 0.0;
            }
            __threadfence();
            bar_sync(kBarHelpers, kHelperThreads);
            if (hth == 0) {
                const int old = atomicAdd(p.done + acc_b, acc_tiles);
                s_flag = (old + acc_tiles == acc_tiles_b);
            }
            bar_sync(kBarHelpers, kHelperThreads);
            if (s_flag) {
                __threadfence();
                const long long t_lo = p.tile_offsets[acc_b];
                const int c_lo = chunk_of(t_lo, n_tiles, G);
                const int c_hi = chunk_of(t_lo + acc_tiles_b - 1, n_tiles, G);
                for (int k = hth; k < kBins; k += kHelperThreads) {
                    double s1 = 0.0, s2 = 0.0;
                    for (int c = c_lo; c <= c_hi; ++c) {
                        // CTAs whose chunk is empty never flushed anything
                        if ((long long)(c + 1) * n_tiles / G == (long long)c * n_tiles / G) continue;
                        const double* q = p.partials + (size_t)(c + acc_b) * kHelperWarps * 400;
#pragma unroll
                        for (int h = 0; h < kHelperWarps; ++h) {
                            s1 += __ldcg(q + h * 400 + k);
                            s2 += __ldcg(q + h * 400 + 200 + k);
                        }
                    }
                    // sklearn.preprocessing.scale (wav_util.py:79): mean, std (ddof=0),
                    // std < 10 eps -> 1
                    const double n = (double)acc_nfr;
                    const double mean = s1 / n;
                    double var = s2 / n - mean * mean;
                    if (var < 0.0) var = 0.0;
                    double sd = sqrt(var);
                    if (sd < 10.0 * 2.220446049250313e-16) sd = 1.0;
                    p.stats[(size_t)acc_b * 400 + k] = mean;
                    p.stats[(size_t)acc_b * 400 + 200 + k] = 1.0 / sd;
                }
            }
            bar_sync(kBarHelpers, kHelperThreads);
            acc_tiles = 0;
        };

        auto epilogue = [&](int i) {
            const Meta m = meta[i & 3];
            if (m.b != acc_b) {
                if (want_stats && acc_b >= 0 && acc_tiles > 0) flush();
                acc_b = m.b;
                acc_tiles_b = m.tiles_b;
                acc_nfr = m.nfr;
            }
            const float* ot = outt + (i & 1) * (kTile * kOutStride);
#pragma unroll 2
            for (int ff = 0; ff < 8; ++ff) {
                const int f = hw * 8 + ff;
                if (f >= m.nf) break;
                float* orow = p.out + (size_t)(m.row0 + f) * kBins;
#pragma unroll
                for (int s = 0; s < 7; ++s) {
                    const int k = lane + 32 * s;
                    if (k < kBins) {
                        const float y = ot[f * kOutStride + k];
                        orow[k] = y;
                        if (want_stats) {
                            const double yd = (double)y;
                            acc1[s] += yd;
                            acc2[s] = fma(yd, yd, acc2[s]);
                        }
                    }
                }
            }
            acc_tiles += 1;
        };

        // tile t (chunk-local): metadata, then start filling PCM stage t % 3.  The
        // async path returns with the copies in flight (one commit group per tile).
        auto issue = [&](int t) {
            if (t < my_tiles) {
                if (hth == 0) {
                    Meta& mn = meta[t & 3];
                    fill_meta(p, tile_begin + t, t == 0 ? -1 : meta[(t - 1) & 3].b, mn);
                    TileRec rec{mn.b, mn.nf, mn.row0};
                    p.tile_rec[tile_begin + t] = rec;
                }
                bar_sync(kBarHelpers, kHelperThreads);
                if (t >= 3) bar_sync(kBarPcmEmpty + (t % 3), kPipeThreads);
                const Meta& m = meta[t & 3];
                uint32_t* dst = pcm + (t % 3) * kPcmWords;
                const bool aligned = (((reinterpret_cast<uintptr_t>(p.samples) + m.sbase * (F32 ? 4 : 2)) & 15) == 0);
                if (!mix && aligned) issue_pcm_tile_async<F32>(p, m, dst, hth);
                else load_pcm_tile<F32>(p, m, dst, hth);
            }
            cp_async_commit();
        };

        issue(0);
        issue(1);
        cp_async_wait<1>();
        bar_arrive(kBarPcmFull + 0, kPipeThreads);
        for (int i = 0; i < my_tiles; ++i) {
            issue(i + 2);
            if (i + 1 < my_tiles) {
                cp_async_wait<1>();          // everything but the newest group: tile i+1 has landed
                bar_arrive(kBarPcmFull + ((i + 1) % 3), kPipeThreads);
            }
            bar_sync(kBarOutFull + (i & 1), kPipeThreads);
            epilogue(i);
            if (i + 2 < my_tiles) bar_arrive(kBarOutEmpty + (i & 1), kPipeThreads);
        }
        if (want_stats) flush();
    }
}

// ---------------------------------------------------------------------------
// z-score: out = (y - mean) * inv_std, in place, one CTA per tile (grid-stride)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_kernel(Params p) {
    const int n_tiles = p.tile_offsets[p.batch];
    __shared__ double s_stats[400];
    int cur_b = -1;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileRec rec = p.tile_rec[tile];
        if (rec.b != cur_b) {
            __syncthreads();
            for (int k = threadIdx.x; k < 400; k += blockDim.x)
                s_stats[k] = p.stats[(size_t)rec.b * 400 + k];
            cur_b = rec.b;
            __syncthreads();
        }
        float4* base = reinterpret_cast<float4*>(p.out + (size_t)rec.row0 * kBins);
        const int n4 = rec.nf * (kBins / 4);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const int k = (i % (kBins / 4)) * 4;
            float4 v = base[i];
            v.x = (float)(((double)v.x - s_stats[k]) * s_stats[200 + k]);
            v.y = (float)(((double)v.y - s_stats[k + 1]) * s_stats[200 + k + 1]);
            v.z = (float)(((double)v.z - s_stats[k + 2]) * s_stats[200 + k + 2]);
            v.w = (float)(((double)v.w - s_stats[k + 3]) * s_stats[200 + k + 3]);
            base[i] = v;
        }
    }
}

template <bool F32>
static size_t main_smem_bytes() {
    const size_t pcm_words = (size_t)kHopRows * (F32 ? kHopWordsF32 : kHopWordsI16);
    size_t b = sizeof(double) * kTabDoubles + sizeof(cplx) * 200 * kTile +
               sizeof(float) * 2 * kTile * kOutStride + sizeof(uint32_t) * (3 * pcm_words);
    return b + 16;
}

}  // namespace spec
}  // namespace asrk

using namespace asrk;
using namespace asrk::spec;

extern "C" size_t asrk_spectrogram_workspace_bytes(int batch, long long total_frames) {
    if (batch < 0 || total_frames < 0) return 0;
    return ws_layout(batch, total_frames, 256).total;
}

extern "C" int asrk_spectrogram_run_phases(const void* samples, int sample_dtype, const float* noise,
                                           const float* gain, const int* snr_db,
                                           const long long* sample_offsets, const long long* sample_counts,
                                           const long long* frame_offsets, const long long* out_row_offsets,
                                           int batch, long long total_frames,
                                           int mode, float* out, void* workspace, size_t workspace_bytes,
                                           asrk_stream_t stream_, int phases) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (batch < 0 || total_frames < 0) return ASRK_E_BADARG;
    if (batch == 0) return ASRK_OK;
    if (!samples || !sample_offsets || !sample_counts || !frame_offsets || !out || !workspace)
        return ASRK_E_BADARG;
    if (sample_dtype != ASRK_DTYPE_I16 && sample_dtype != ASRK_DTYPE_F32) return ASRK_E_BADARG;
    if (mode != ASRK_SPEC_FBANK && mode != ASRK_SPEC_ASRT && mode != ASRK_SPEC_FBANK_RAW) return ASRK_E_BADARG;
    if (noise && sample_dtype != ASRK_DTYPE_F32) return ASRK_E_BADARG;
    if (noise && !gain && !snr_db) return ASRK_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return ASRK_E_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return ASRK_E_ALIGN;
    int grid = sm_count() > 256 ? 256 : sm_count();
    const int cta_limit = (phases >> 16) & 0x7fff;
    if (cta_limit > 0 && cta_limit < grid) grid = cta_limit;
    const WsLayout l = ws_layout(batch, total_frames, 256);
    if (workspace_bytes < l.total) return ASRK_E_WORKSPACE;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);

    Params p;
    p.samples = samples;
    p.noise = noise;
    p.gain = gain;
    p.sample_offsets = sample_offsets;
    p.sample_counts = sample_counts;
    p.frame_offsets = frame_offsets;
    p.out_row_offsets = out_row_offsets;
    p.batch = batch;
    p.mode = mode;
    p.out = out;
    p.tables = reinterpret_cast<double*>(ws + l.tables);
    p.tile_offsets = reinterpret_cast<int*>(ws + l.tile_offsets);
    p.done = reinterpret_cast<int*>(ws + l.done);
    p.stats = reinterpret_cast<double*>(ws + l.stats);
    p.partials = reinterpret_cast<double*>(ws + l.partials);
    p.tile_rec = reinterpret_cast<TileRec*>(ws + l.tile_rec);

    if (noise && !gain && !(phases & ASRK_PHASE_SPEC_SETUP)) {
        p.gain = reinterpret_cast<float*>(ws + l.gains);   // computed by an earlier SETUP phase
    } else if (noise && !gain) {
        // K from snr_db with the reference's float32 arithmetic (noise.cu)
        float* g = reinterpret_cast<float*>(ws + l.gains);
        const int st = asrk_snr2k_run(reinterpret_cast<const float*>(samples), noise, sample_offsets,
                                      sample_counts, snr_db, batch, g, stream_);
        if (st != ASRK_OK) return st;
        p.gain = g;
    }

    if (phases & ASRK_PHASE_SPEC_SETUP)
        setup_kernel<<<1, 1024, 0, stream>>>(const_cast<double*>(p.tables), frame_offsets, batch,
                                             p.tile_offsets, p.done);
    if (!(phases & ASRK_PHASE_SPEC_MAIN)) {
    } else if (sample_dtype == ASRK_DTYPE_F32) {
        const size_t smem = main_smem_bytes<true>();
        cudaFuncSetAttribute(spectrogram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        spectrogram_kernel<true><<<grid, kThreads, smem, stream>>>(p);
    } else {
        const size_t smem = main_smem_bytes<false>();
        cudaFuncSetAttribute(spectrogram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        spectrogram_kernel<false><<<grid, kThreads, smem, stream>>>(p);
    }
    if (mode == ASRK_SPEC_FBANK && (phases & ASRK_PHASE_SPEC_NORMALIZE))
        normalize_kernel<<<sm_count() * 4, 256, 0, stream>>>(p);
    return launch_status();
}

extern "C" int asrk_spectrogram_run(const void* samples, int sample_dtype, const float* noise,
                                    const float* gain, const int* snr_db,
                                    const long long* sample_offsets, const long long* sample_counts,
                                    const long long* frame_offsets, const long long* out_row_offsets,
                                    int batch, long long total_frames,
                                    int mode, float* out, void* workspace, size_t workspace_bytes,
                                    asrk_stream_t stream_) {
    return asrk_spectrogram_run_phases(samples, sample_dtype, noise, gain, snr_db, sample_offsets,
                                       sample_counts, frame_offsets, out_row_offsets, batch, total_frames,
                                       mode, out, workspace, workspace_bytes, stream_, ASRK_PHASE_ALL);
}
