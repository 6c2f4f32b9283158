// Part 1 of the hot path: spectrogram features (+ fused noise mix, + z-score).
//
// Reference behaviour reproduced: /root/reference/util/wav_util.py:49-79
// (compute_fbank), :82-112 (compute_fbank_from_asrt), util/noise.py:48-52,108.
//
// B200 design (DESIGN.md "spectrogram kernel"):
//   * the reference computes in float64 and the tolerance (1e-4 on log|X|, then
//     amplified by the z-score) is not reachable with fp32 butterflies on
//     high-dynamic-range audio, so the 400-point transform runs in fp64 -- B200
//     has a half-rate fp64 pipe (64 lanes/SM/clk), which is the bound of this
//     kernel (register-resident codelets reach > 90 % of it, tools/ubench);
//     everything after |X|^2 (sqrt, log, z-score) is fp32.
//   * ONE persistent kernel, one CTA per SM.  A tile is 32 consecutive frames of
//     one utterance: LANE = FRAME, WARP = ROLE.  The 400-point real DFT is a
//     200-point complex DFT (20 x 10 Cooley-Tukey, both factors twiddle-free
//     prime-factor codelets) + the real-input split; ten "FFT warps" each own one
//     of the ten residues, so window values and twiddles are warp-uniform and come
//     from the constant bank (no load/store-unit traffic), the PCM tile is staged
//     in padded shared memory by cp.async, and the only exchange between the two
//     passes is one conflict-free 100 KB fp64 buffer.
//   * four "helper warps" run concurrently: they claim tiles from an atomic counter
//     (in utterance order), stage the PCM two tiles ahead (mixing in K*noise on the
//     fly in noise mode), and move finished 32 x 200 log-magnitude tiles to global
//     memory with coalesced stores.  FULL/EMPTY named barriers decouple the two
//     groups by up to a tile.
//   * z-score: the helpers accumulate the per-bin column sums of every tile while
//     they copy it out (fp32, shifted by the tile's first row so that nothing
//     cancels), combine them per tile in fp64 in a fixed order (reproducible, no float
//     atomics).  A tiny kernel turns the tile partials of every utterance into mean and
//     1/std, and a purely streaming kernel with the whole chip's memory parallelism
//     normalises in place (the rows are largely still in L2).  [Measured alternatives:
//     an in-kernel z-score by the CTA that retires an utterance's last tile -- one
//     CTA's 448 threads cannot keep enough L2 requests in flight, 40 us per utterance;
//     per-tile release fence + counter in the helpers -- the membar costs 19 us.]
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
constexpr int kOutStride = 201;                      // padded row of the out tile
// hop rows stay 16-byte aligned (for 16-byte async copies) and are padded by 16
// bytes: lane f reads word 84 f + c -> 8 distinct banks, a 4-way conflict on the
// 20 sample loads of a thread per tile (its only other shared accesses are the
// conflict-free exchange), instead of the 16-way conflict of the unpadded layout.
constexpr int kHopWordsI16 = 84;                     // 80 words of int16 pairs + 4 pad
constexpr int kHopWordsF32 = 164;                    // 160 words + 4 pad
constexpr int kTabDoubles = 1200;                    // window[400] | tw[r][k1] | P[k]
constexpr int kMaxBatch = 2047;                      // utterances per launch (tile prefix in smem)
constexpr int kRing = 8;                             // tile metadata ring

__device__ const double g_tab[kTabDoubles] = {
#include "asrk_tables.inc"
};

struct Meta {               // one tile, written by helper thread 0
    int valid;
    int b;
    int f0;
    int nf;
    int ntiles_b;            // tiles of utterance b
    int tile;                // global tile index (row of the partial sums)
    long long sbase;         // first sample of the utterance in the ragged buffer
    long long nsamp;         // samples of the utterance
    long long row0;          // output row of frame 0 of the utterance
    long long nfr;           // frames of the utterance
    float half_mag;
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
    int* counters;           // [0] next tile (zeroed per launch)
    int* tile_off_g;         // [B + 1] first tile of every utterance (written by CTA 0)
    double2* partials;       // [tiles][200]: per-tile column sums (sum y, sum y^2)
    float* stats;            // [B][3][200]: mean (hi, lo), 1/std
};

struct WsLayout {
    size_t counters, tile_off, gains, stats, partials, total;
};

static WsLayout ws_layout(int batch, long long total_frames) {
    WsLayout l;
    size_t o = 0;
    l.counters = o;  o = align_up(o + sizeof(int) * 4, 256);
    l.tile_off = o;  o = align_up(o + sizeof(int) * (size_t)(batch + 1), 256);
    l.gains = o;     o = align_up(o + sizeof(float) * (size_t)batch, 256);
    l.stats = o;     o = align_up(o + sizeof(float) * 3 * kBins * (size_t)batch, 256);
    l.partials = o;
    const size_t max_tiles = (size_t)(total_frames / kTile) + (size_t)batch + 1;
    o = align_up(o + sizeof(double2) * kBins * max_tiles, 256);
    l.total = o;
    return l;
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double i16_to_f64(int v) {
    // exact int -> double with one integer op and one DADD (the I2F.F64 path is
    // a quarter-rate conversion): 2^52 + 2^31 + v has v in its low mantissa bits.
    return __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)v)) - 4503601774854144.0;
}

// log(|X| * mag + 1) from p4 = 4 |X|^2, natural log (wav_util.py:76,107,111), fp32:
// v = 1 + m is split into 2^e * f exactly; lg2.approx on f in [1,2) has an absolute
// error of 2^-22, so the result carries an absolute error of ~2e-7 at any magnitude
// (the tolerance is 1e-4 * max(|ref|, 1)); sqrt.approx is good to 2^-23 relative.
__device__ __forceinline__ float log_mag(float p4, float half_mag) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p4));
    const float v = fmaf(r, half_mag, 1.0f);
    const int vi = __float_as_int(v);
    const float e = (float)((vi >> 23) - 127);
    const float f = __int_as_float((vi & 0x007fffff) | 0x3f800000);
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(f));
    return (e + l2) * 0.69314718056f;
}

// Synchronous staging of one PCM tile (unaligned utterances, and the noise mix:
// fl32(signal + fl32(K * noise)) is formed on the way in).
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
            if (us + 4 <= m.nsamp && ((reinterpret_cast<uintptr_t>(g) & 15) == 0) &&
                (!nz || ((reinterpret_cast<uintptr_t>(nz + us) & 15) == 0))) {
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

// Tile -> utterance, frame range and the utterance's constants (helper thread 0).
__device__ void fill_meta(const Params& p, const int* tile_off, int tile, Meta& m) {
    int lo = 0, hi = p.batch - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_off[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    int b = lo;
    while (tile >= tile_off[b + 1]) ++b;          // skip utterances without tiles
    m.valid = 1;
    m.tile = tile;
    m.b = b;
    m.ntiles_b = tile_off[b + 1] - tile_off[b];
    m.f0 = (tile - tile_off[b]) * kTile;
    const long long fo = p.frame_offsets[b];
    m.nfr = p.frame_offsets[b + 1] - fo;
    const long long rem = m.nfr - m.f0;
    m.nf = rem < kTile ? (int)rem : kTile;
    m.sbase = p.sample_offsets[b];
    m.nsamp = p.sample_counts[b];
    m.row0 = p.out_row_offsets ? p.out_row_offsets[b] : fo;
    m.half_mag = 0.5f * ((p.mode == ASRK_SPEC_ASRT) ? (1.0f / (float)m.nsamp) : 1.0f);
    m.gain = (p.noise && p.gain) ? p.gain[b] : 0.0f;
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
// Named barriers (id 0 is __syncthreads).  FULL/EMPTY pairs hand the PCM stages and
// the two out tiles between the FFT warps and the helper warps so that neither side
// waits for the other unless it is a whole tile behind.
enum : int {
    kBarHelpers = 1,      // helper warps only
    kBarExchA = 2,        // FFT warps only: pass 1 stores -> pass 2 loads
    kBarExchB = 3,        // FFT warps only: pass 2 loads -> next pass 1 stores
    kBarPcmFull = 4,      // +stage (4,5,6)
    kBarPcmEmpty = 7,     // +stage (7,8,9)
    kBarOutFull = 10,     // +slot (10,11)
    kBarOutEmpty = 12,    // +slot (12,13)
};
constexpr int kPipeThreads = kThreads;   // threads on a FULL/EMPTY barrier

__device__ __forceinline__ void bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int n) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

template <bool F32>
struct Cfg {
    static constexpr int kStages = F32 ? 2 : 3;          // PCM stages (tiles staged ahead: kStages - 1)
    static constexpr int kPcmWords = kHopRows * (F32 ? kHopWordsF32 : kHopWordsI16);
};

// Synchronisation protocol.  Both groups walk the same iteration space t = 0, 1, ...
// The tiles a CTA claims are valid for t < V and invalid from V on (the counter ran
// out); both groups stop in iteration V:
//   helpers, iteration t : claim + stage tile t + kAhead (waits PcmEmpty of the stage),
//                          arrive PcmFull(t + 1), [valid] wait OutFull(t), store rows +
//                          column sums, arrive OutEmpty(t), publish the tile's partial
//                          sums, retire the tile
//   FFT,     iteration t : wait PcmFull(t), [valid] pass 1 (arrive PcmEmpty after the
//                          loads), ExchA, pass-2 loads, ExchB, wait OutEmpty(t - 2),
//                          pass 2, arrive OutFull(t)
template <bool F32>
__global__ void __launch_bounds__(kThreads, 1) spectrogram_kernel(Params p) {
    constexpr int kStages = Cfg<F32>::kStages;
    constexpr int kAhead = kStages - 1;
    constexpr int kPcmWords = Cfg<F32>::kPcmWords;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* exch = reinterpret_cast<cplx*>(smem_raw);                           // [200][32]
    float* outt = reinterpret_cast<float*>(exch + 200 * kTile);               // [2][32*201]
    uint32_t* pcm = reinterpret_cast<uint32_t*>(outt + 2 * kTile * kOutStride);  // [kStages][kPcmWords]
    int* tile_off = reinterpret_cast<int*>(pcm + kStages * kPcmWords);        // [kMaxBatch + 1]
    float* s_col = reinterpret_cast<float*>(tile_off + kMaxBatch + 1);        // [4][3][200] helper column sums
    double* tab = reinterpret_cast<double*>(s_col + kHelperWarps * 3 * kBins);   // [1200], 16-byte aligned
    __shared__ Meta meta[kRing];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < kTabDoubles; i += kThreads) tab[i] = g_tab[i];
    if (warp == 0) {
        // exclusive scan of ceil(n_frames / 32) over the utterances
        int carry = 0;
        for (int base = 0; base < p.batch; base += 32) {
            const int i = base + lane;
            int v = 0;
            if (i < p.batch) {
                long long n = p.frame_offsets[i + 1] - p.frame_offsets[i];
                if (n < 0) n = 0;
                v = (int)((n + kTile - 1) / kTile);
            }
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            if (i < p.batch) tile_off[i] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) tile_off[p.batch] = carry;
    }
    __syncthreads();
    const int total_tiles = tile_off[p.batch];
    const bool want_stats = (p.mode == ASRK_SPEC_FBANK);
    if (blockIdx.x == 0 && want_stats)
        for (int i = tid; i <= p.batch; i += kThreads) p.tile_off_g[i] = tile_off[i];

    if (warp < kFftWarps) {
        // ------------------------------ FFT warps ------------------------------
        // the role: residue class in pass 1, row pair in pass 2; window values and twiddles
        // are warp-uniform shared-memory broadcasts
        const int r = warp;
        const double2* tabW2 = reinterpret_cast<const double2*>(tab);
        const cplx* tabT = reinterpret_cast<const cplx*>(tab + 400) + r * 20;
        const cplx* tabP = reinterpret_cast<const cplx*>(tab + 800);
        for (int t = 0;; ++t) {
            const int st = t % kStages;
            const int s = t & 1;
            bar_sync(kBarPcmFull + st, kPipeThreads);
            if (!meta[t % kRing].valid) break;
            const uint32_t* pc = pcm + st * kPcmWords;
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
                    const double2 w2 = tabW2[10 * n1 + r];
                    z[n1] = cplx{x0 * w2.x, x1 * w2.y};   // wav_util.py:71 data_line * w
                }
                bar_arrive(kBarPcmEmpty + st, kPipeThreads);
                dft20(z, y);
                exch[r * kTile + lane] = y[0];
#pragma unroll
                for (int k1 = 1; k1 < 20; ++k1) exch[(k1 * 10 + r) * kTile + lane] = cmul(y[k1], tabT[k1]);
            }
            bar_sync(kBarExchA, kFftThreads);
            {
                const int k1a = r, k1b = (r == 0) ? 10 : 20 - r;
                cplx ia[10], ib[10], za[10], zb[10];
#pragma unroll
                for (int n2 = 0; n2 < 10; ++n2) ia[n2] = exch[(k1a * 10 + n2) * kTile + lane];
#pragma unroll
                for (int n2 = 0; n2 < 10; ++n2) ib[n2] = exch[(k1b * 10 + n2) * kTile + lane];
                bar_sync(kBarExchB, kFftThreads);     // every warp holds its rows: the exchange is free
                dft10(ia, za);   // za[k2] = Z[k1a + 20 k2]
                dft10(ib, zb);   // zb[k2] = Z[k1b + 20 k2]
                if (t >= 2) bar_sync(kBarOutEmpty + s, kPipeThreads);
                float* ot = outt + s * (kTile * kOutStride) + lane * kOutStride;
                const float hm = meta[t % kRing].half_mag;
                if (r != 0) {
#pragma unroll
                    for (int k2 = 0; k2 < 10; ++k2) {
                        const int k = r + 20 * k2;
                        double pk, pm;
                        split_pair(za[k2], zb[9 - k2], tabP[k], pk, pm);
                        ot[k] = log_mag((float)pk, hm);
                        ot[200 - k] = log_mag((float)pm, hm);
                    }
                } else {
                    // row k1 = 0: bins 20 k2, mirror 20 (10 - k2); k2 = 0 and 5 are their own mirror
#pragma unroll
                    for (int k2 = 0; k2 <= 5; ++k2) {
                        const int k = 20 * k2;
                        double pk, pm;
                        split_pair(za[k2], za[(10 - k2) % 10], tabP[k], pk, pm);
                        ot[k] = log_mag((float)pk, hm);
                        if (k2 != 0 && k2 != 5) ot[200 - k] = log_mag((float)pm, hm);
                    }
                    // row k1 = 10: bins 10 + 20 k2, mirror 10 + 20 (9 - k2)
#pragma unroll
                    for (int k2 = 0; k2 < 5; ++k2) {
                        const int k = 10 + 20 * k2;
                        double pk, pm;
                        split_pair(zb[k2], zb[9 - k2], tabP[k], pk, pm);
                        ot[k] = log_mag((float)pk, hm);
                        ot[200 - k] = log_mag((float)pm, hm);
                    }
                }
            }
            bar_arrive(kBarOutFull + s, kPipeThreads);
        }
    } else {
        // ----------------------------- helper warps ----------------------------
        const int hw = warp - kFftWarps;                          // 0..3
        const int hth = hw * 32 + lane;
        const bool mix = (p.noise != nullptr);
        // Tile metadata is prepared one iteration before it is needed by the first lane of
        // helper warp 1 (claim from the atomic counter, look the utterance up), the retire
        // bookkeeping below runs on the first lane of helper warp 0: neither serialises the
        // other, and neither round trip is on the staging path.
        constexpr int kMetaThread = 32;
        int next_tile = total_tiles;
        if (hth == kMetaThread) next_tile = atomicAdd(p.counters, 1);
        auto prepare_meta = [&](int t) {
            if (hth == kMetaThread) {
                Meta& mn = meta[t % kRing];
                const int tile = next_tile;
                if (tile < total_tiles) {
                    next_tile = atomicAdd(p.counters, 1);
                    fill_meta(p, tile_off, tile, mn);
                } else {
                    mn.valid = 0;
                }
            }
        };

        // tile t (metadata prepared earlier): start filling PCM stage t % kStages.  The
        // async path returns with the copies in flight (one commit group per tile).
        auto issue = [&](int t) {
            bar_sync(kBarHelpers, kHelperThreads);
            if (t >= kStages) bar_sync(kBarPcmEmpty + (t % kStages), kPipeThreads);
            const Meta& m = meta[t % kRing];
            if (m.valid) {
                uint32_t* dst = pcm + (t % kStages) * kPcmWords;
                const bool aligned = (((reinterpret_cast<uintptr_t>(p.samples) + m.sbase * (F32 ? 4 : 2)) & 15) == 0);
                if (!mix && aligned) issue_pcm_tile_async<F32>(p, m, dst, hth);
                else load_pcm_tile<F32>(p, m, dst, hth);
            }
            cp_async_commit();
        };

        // rows of tile t to global memory (coalesced), and this warp's column sums of
        // (y - c), (y - c)^2 with c = the warp's first row, into shared memory
        // the un-normalised rows are read again by the z-score kernel: keep them in L2
        const uint64_t keep = want_stats ? l2_policy_evict_last() : l2_policy_evict_first();
        auto epilogue = [&](int t) {
            const Meta& m = meta[t % kRing];
            const float* ot = outt + (t & 1) * (kTile * kOutStride);
            float c[7], sm[7], sq[7];
#pragma unroll
            for (int e = 0; e < 7; ++e) { c[e] = 0.f; sm[e] = 0.f; sq[e] = 0.f; }
#pragma unroll 2
            for (int ff = 0; ff < 8; ++ff) {
                const int f = hw * 8 + ff;
                if (f >= m.nf) break;
                float* orow = p.out + (size_t)(m.row0 + m.f0 + f) * kBins;
#pragma unroll
                for (int e = 0; e < 7; ++e) {
                    const int k = lane + 32 * e;
                    if (k < kBins) {
                        const float y = ot[f * kOutStride + k];
                        stg_hint(orow + k, y, keep);
                        if (ff == 0) c[e] = y;
                        const float d = y - c[e];
                        sm[e] += d;
                        sq[e] = fmaf(d, d, sq[e]);
                    }
                }
            }
            if (want_stats) {
                float* sc = s_col + hw * 3 * kBins;
#pragma unroll
                for (int e = 0; e < 7; ++e) {
                    const int k = lane + 32 * e;
                    if (k < kBins) { sc[k] = c[e]; sc[kBins + k] = sm[e]; sc[2 * kBins + k] = sq[e]; }
                }
            }
        };

#pragma unroll 1
        for (int t = 0; t <= kAhead; ++t) prepare_meta(t);
#pragma unroll 1
        for (int t = 0; t < kAhead; ++t) issue(t);
        cp_async_wait<kAhead - 1>();
        bar_arrive(kBarPcmFull + 0, kPipeThreads);
        for (int t = 0;; ++t) {
            issue(t + kAhead);
            prepare_meta(t + kAhead + 1);
            cp_async_wait<kAhead - 1>();          // everything but the newest group(s): tile t+1 has landed
            bar_arrive(kBarPcmFull + ((t + 1) % kStages), kPipeThreads);
            if (!meta[t % kRing].valid) break;
            bar_sync(kBarOutFull + (t & 1), kPipeThreads);
            epilogue(t);
            bar_sync(kBarHelpers, kHelperThreads);          // the out tile has been read by all four warps
            bar_arrive(kBarOutEmpty + (t & 1), kPipeThreads);
            if (!want_stats) continue;
            // per-tile column sums, un-shifted, in fp64 and a fixed order:
            //   sum y = s + n c,  sum y^2 = q + 2 c s + n c^2   over the four warps' row groups
            const Meta& m = meta[t % kRing];
            for (int k = hth; k < kBins; k += kHelperThreads) {
                double a1 = 0.0, a2 = 0.0;
#pragma unroll
                for (int w = 0; w < kHelperWarps; ++w) {
                    int nw = m.nf - 8 * w;
                    nw = nw < 0 ? 0 : (nw > 8 ? 8 : nw);
                    const double n = (double)nw;
                    const double cc = (double)s_col[w * 3 * kBins + k];
                    const double ss = (double)s_col[w * 3 * kBins + kBins + k];
                    const double qq = (double)s_col[w * 3 * kBins + 2 * kBins + k];
                    a1 += fma(n, cc, ss);
                    a2 += fma(cc, fma(n, cc, 2.0 * ss), qq);
                }
                p.partials[(size_t)m.tile * kBins + k] = make_double2(a1, a2);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// mean and 1/std of every utterance from its tiles' partial sums, one CTA each
// (sklearn.preprocessing.scale, wav_util.py:79: std with ddof = 0, std < 10 eps -> 1)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_kernel(Params p) {
    const int b = blockIdx.x;
    const int t_lo = p.tile_off_g[b], t_hi = p.tile_off_g[b + 1];
    const long long nfr = p.frame_offsets[b + 1] - p.frame_offsets[b];
    for (int k = threadIdx.x; k < kBins; k += blockDim.x) {
        double a1 = 0.0, a2 = 0.0;
        for (int q = t_lo; q < t_hi; ++q) {               // fixed order: reproducible
            const double2 v = p.partials[(size_t)q * kBins + k];
            a1 += v.x;
            a2 += v.y;
        }
        const double n = (double)(nfr > 0 ? nfr : 1);
        const double mean = a1 / n;
        double var = a2 / n - mean * mean;
        if (var < 0.0) var = 0.0;
        double sd = sqrt(var);
        if (sd < 10.0 * 2.220446049250313e-16) sd = 1.0;
        const float mh = (float)mean;
        float* st = p.stats + (size_t)b * 3 * kBins;
        st[k] = mh;
        st[kBins + k] = (float)(mean - (double)mh);
        st[2 * kBins + k] = (float)(1.0 / sd);
    }
}

// ---------------------------------------------------------------------------
// z-score: out = (y - mean) / std, in place; grid (x, utterance), pure streaming
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_kernel(Params p) {
    const int b = blockIdx.y;
    __shared__ __align__(16) float s_stat[3 * kBins];
    for (int k = threadIdx.x; k < 3 * kBins; k += blockDim.x) s_stat[k] = p.stats[(size_t)b * 3 * kBins + k];
    __syncthreads();
    const long long fo = p.frame_offsets[b];
    const long long nfr = p.frame_offsets[b + 1] - fo;
    const long long row0 = p.out_row_offsets ? p.out_row_offsets[b] : fo;
    float4* base = reinterpret_cast<float4*>(p.out + (size_t)row0 * kBins);
    const long long n4 = nfr * (kBins / 4);
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int kU = 4;
    const uint64_t drop = l2_policy_evict_first();
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += kU * stride) {
        float4 v[kU];
#pragma unroll
        for (int e = 0; e < kU; ++e) {
            const long long i = i0 + e * stride;
            if (i < n4) v[e] = ldg_hint(base + i, drop);
        }
#pragma unroll
        for (int e = 0; e < kU; ++e) {
            const long long i = i0 + e * stride;
            if (i < n4) {
                const int k = (int)(i % (kBins / 4)) * 4;
                const float4 mh = *reinterpret_cast<const float4*>(s_stat + k);
                const float4 ml = *reinterpret_cast<const float4*>(s_stat + kBins + k);
                const float4 iv = *reinterpret_cast<const float4*>(s_stat + 2 * kBins + k);
                float4 o;
                o.x = ((v[e].x - mh.x) - ml.x) * iv.x;
                o.y = ((v[e].y - mh.y) - ml.y) * iv.y;
                o.z = ((v[e].z - mh.z) - ml.z) * iv.z;
                o.w = ((v[e].w - mh.w) - ml.w) * iv.w;
                stg_evict_first(base + i, o);
            }
        }
    }
}

template <bool F32>
static size_t main_smem_bytes() {
    return sizeof(cplx) * 200 * kTile + sizeof(float) * 2 * kTile * kOutStride +
           sizeof(uint32_t) * (size_t)Cfg<F32>::kStages * Cfg<F32>::kPcmWords + sizeof(int) * (kMaxBatch + 1) +
           sizeof(float) * kHelperWarps * 3 * kBins + sizeof(double) * kTabDoubles + 16;
}

template <bool F32>
static void launch_main(const Params& p, int grid, cudaStream_t stream) {
    const size_t smem = main_smem_bytes<F32>();
    cudaFuncSetAttribute(spectrogram_kernel<F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spectrogram_kernel<F32><<<grid, kThreads, smem, stream>>>(p);
}

}  // namespace spec
}  // namespace asrk

using namespace asrk;
using namespace asrk::spec;

extern "C" size_t asrk_spectrogram_workspace_bytes(int batch, long long total_frames) {
    if (batch < 0 || total_frames < 0) return 0;
    return ws_layout(batch, total_frames).total;
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
    int grid = sm_count();
    const int cta_limit = (phases >> 16) & 0x7fff;
    if (cta_limit > 0 && cta_limit < grid) grid = cta_limit;
    const WsLayout l = ws_layout(batch, total_frames);
    if (workspace_bytes < l.total) return ASRK_E_WORKSPACE;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);

    const float* gains = gain;
    if (noise && !gain) {
        // K from snr_db with the reference's float32 arithmetic (noise.cu)
        float* gw = reinterpret_cast<float*>(ws + l.gains);
        if (phases & ASRK_PHASE_SPEC_SETUP) {
            const int st = asrk_snr2k_run(reinterpret_cast<const float*>(samples), noise, sample_offsets,
                                          sample_counts, snr_db, batch, gw, stream_);
            if (st != ASRK_OK) return st;
        }
        gains = gw;
    }
    if (!(phases & (ASRK_PHASE_SPEC_MAIN | ASRK_PHASE_SPEC_NORMALIZE))) return launch_status();

    // the kernel locates tiles through a prefix array in shared memory: at most
    // kMaxBatch utterances per launch, larger batches go in slices
    for (int b0 = 0; b0 < batch; b0 += kMaxBatch) {
        const int nb = (batch - b0) < kMaxBatch ? (batch - b0) : kMaxBatch;
        Params p;
        p.samples = samples;
        p.noise = noise;
        p.gain = gains ? gains + b0 : nullptr;
        p.sample_offsets = sample_offsets + b0;
        p.sample_counts = sample_counts + b0;
        p.frame_offsets = frame_offsets + b0;
        p.out_row_offsets = out_row_offsets ? out_row_offsets + b0 : nullptr;
        p.batch = nb;
        p.mode = mode;
        p.out = out;
        p.counters = reinterpret_cast<int*>(ws + l.counters);
        p.tile_off_g = reinterpret_cast<int*>(ws + l.tile_off);
        p.partials = reinterpret_cast<double2*>(ws + l.partials);
        p.stats = reinterpret_cast<float*>(ws + l.stats) + (size_t)b0 * 3 * kBins;
        if ((phases & ASRK_PHASE_SPEC_MAIN) &&
            cudaMemsetAsync(p.counters, 0, sizeof(int) * 4, stream) != cudaSuccess)
            return ASRK_E_CUDA;
        if (phases & ASRK_PHASE_SPEC_MAIN) {
            if (sample_dtype == ASRK_DTYPE_I16) launch_main<false>(p, grid, stream);
            else launch_main<true>(p, grid, stream);
        }
        if (mode == ASRK_SPEC_FBANK && (phases & ASRK_PHASE_SPEC_NORMALIZE)) {
            stats_kernel<<<nb, 256, 0, stream>>>(p);
            normalize_kernel<<<dim3(16, nb), 256, 0, stream>>>(p);
        }
    }
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
