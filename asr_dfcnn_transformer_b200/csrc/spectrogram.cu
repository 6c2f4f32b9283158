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
//     kernel; everything after |X|^2 (sqrt, log, z-score) is fp32.
//   * ONE persistent kernel, one CTA per SM: kTeams independent teams of five warps.
//     A team transforms 16 consecutive frames at a time: lane = (frame, role half), the
//     two half-warps of warp q own roles q and q + 5 of the 20 x 10 Cooley-Tukey split
//     of the 200-point complex DFT (400-point real DFT = 200-point complex DFT + real-
//     input split; both factors twiddle-free prime-factor codelets), so window values
//     and twiddles are half-warp broadcasts from shared memory and the two passes
//     exchange through the team's own conflict-free 50 KB fp64 buffer with team-wide
//     named barriers only.  While one team loads, exchanges or copies out, the others
//     keep the fp64 pipe busy.
//   * instruction diet (round 2; the kernel is co-bound by issue slots): fewer fp64
//     instructions in the codelets (asrk_fft.cuh), a cheaper int16 -> fp64 step (see
//     ASRK_SPEC_CONV below); 16-byte exchange accesses; the out tile is copied to global memory
//     and summed for the z-score with 16-byte accesses (lane = (4 bins, row phase));
//     log(|X| + 1) is one MUFU.SQRT + FFMA + MUFU.LG2 + FMUL; role 0's operand selects
//     run only in the warp that owns role 0.
//   * a team claims units of 32 consecutive frames of one utterance from an atomic
//     counter (in utterance order) and stages the PCM of a 16-frame sub-tile with
//     16-byte cp.async behind the previous sub-tile's second pass (zero-filled past the
//     end; the noise mix fl32(s + fl32(K n)) is applied on the way in).
//   * z-score: per-bin sums of every unit are accumulated during the copy-out (fp32,
//     shifted by the unit's first row so that nothing cancels; un-shifted in fp64, fixed
//     order: reproducible, no float atomics).  The team that completes an utterance
//     (release fence one unit late + per-utterance counter: the fence then has nothing to
//     wait for) turns the unit sums into mean / 1/std inside the kernel; one streaming
//     kernel with the whole chip's memory parallelism normalises in place behind it
//     (newest utterances first: their rows are still in L2) and runs next to the CTC
//     kernel.  Two ways of normalising INSIDE the transform kernel were built and measured
//     this round (a dedicated z-score warp per CTA fed by TMA bulk copies; one 16-row
//     chunk per team and sub-tile, also through TMA) -- both parity-green, both slower
//     than the trailing kernel because the transform is latency bound and every
//     instruction added to its warps costs more than the HBM round trip it saves
//     (profiles/r2_spectrogram.md).
#include <math.h>
#include <stdlib.h>

#include "asrk_common.cuh"
#include "asrk_fft.cuh"

namespace asrk {
namespace spec {

constexpr int kTeamWarps = 5;
constexpr int kTeamThreads = kTeamWarps * 32;        // 160
constexpr int kSub = 16;                             // frames per sub-tile (= lanes of a half-warp)
constexpr int kUnit = 32;                            // frames per claimed unit
constexpr int kHop = 160, kFrameLen = 400, kBins = 200;
constexpr int kOutStride = 204;                      // padded, 16-byte aligned row of the out tile
constexpr int kSubHopRows = kSub + 2;                // (kSub-1)*160+400 samples
// hop rows stay 16-byte aligned (for 16-byte async copies) and are padded by 16
// bytes: frame f reads word 84 f + c -> the 16 frames of a half-warp hit 8 distinct
// banks (2-way conflict on the 20 sample loads of a thread per sub-tile).
constexpr int kHopWordsI16 = 84;                     // 80 words of int16 pairs + 4 pad
constexpr int kHopWordsF32 = 164;                    // 160 words + 4 pad
constexpr int kTabDoubles = 1200;                    // window[400] | tw[r][k1] | P[k]
constexpr int kMaxBatch = 1023;                      // utterances per launch (unit prefix in smem)
constexpr int kZRows = 16;                           // rows per z-score chunk
constexpr int kZBytes = kZRows * kBins * 4;          // 12 800

__device__ const double g_tab[kTabDoubles] = {
#include "asrk_tables.inc"
};

struct __align__(16) cplx16 {   // cplx with the alignment that makes exchange accesses LDS/STS.128
    double x, y;
};

struct Meta {               // one unit, written by the team's first thread one unit ahead
    int valid;
    int b;
    int f0;
    int ntiles_b;            // units of utterance b
    int tile;                // unit index (row of the partial sums)
    float gain;
    // raw per-utterance constants, copied global -> shared with cp.async (no registers, no stall of the
    // claiming thread); complete at the team's next "PCM is in place" barrier
    long long fo, fo1;       // frame_offsets[b], frame_offsets[b + 1]
    long long sbase;         // first sample of the utterance in the ragged buffer
    long long nsamp;         // samples of the utterance
    long long row0;          // out_row_offsets[b] (when given)
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
    int zscore_in_kernel;
    float* out;
    // workspace
    int* counters;           // [0] next unit                                   (zeroed per launch)
    int* done;               // [B] published units per utterance               (zeroed per launch)
    int* tile_off_g;         // [B + 1] first unit of every utterance (written by CTA 0)
    double2* partials;       // [units][200]: per-unit column sums (sum y, sum y^2)
    float* stats;            // [B][3][200]: mean (hi, lo), 1/std
};

struct WsLayout {
    size_t counters, done, zero_bytes, tile_off, gains, stats, partials, total;
};

static size_t max_units(int batch, long long total_frames) {
    return (size_t)(total_frames / kUnit) + (size_t)batch + 1;
}

static WsLayout ws_layout(int batch, long long total_frames) {
    WsLayout l;
    size_t o = 0;
    // counters | done are contiguous: one memset per launch
    l.counters = o;  o = align_up(o + sizeof(int) * 4, 256);
    l.done = o;      o = align_up(o + sizeof(int) * (size_t)(batch + 1), 256);
    l.zero_bytes = o;
    l.tile_off = o;  o = align_up(o + sizeof(int) * (size_t)(batch + 1), 256);
    l.gains = o;     o = align_up(o + sizeof(float) * (size_t)batch, 256);
    l.stats = o;     o = align_up(o + sizeof(float) * 3 * kBins * (size_t)batch, 256);
    l.partials = o;
    o = align_up(o + sizeof(double2) * kBins * max_units(batch, total_frames), 256);
    l.total = o;
    return l;
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
// int16 sample pair -> fp64.  Two variants (ASRK_SPEC_CONV, measured in profiles/r2_spectrogram.md):
//   1  sign-extend + I2F.F64.S32 (conversion on the XU pipe), window multiply = DMUL
//   0  no conversion instruction: the pair is biased to unsigned with one XOR and each half is
//      byte-permuted into mantissa bits 32..47 of 2^20 (d = 2^20 + 2^15 + x exactly, low word zero);
//      x w = d w - (2^20 + 2^15) w is ONE DFMA with the constant tabulated next to the window.
#ifndef ASRK_SPEC_CONV
#define ASRK_SPEC_CONV 1
#endif
#ifndef ASRK_SPEC_TWGEN
#define ASRK_SPEC_TWGEN 0
#endif
#ifndef ASRK_SPEC_PGEN
#define ASRK_SPEC_PGEN 0
#endif
#if ASRK_SPEC_CONV == 0
constexpr double kMagic = 1081344.0;                 // 2^20 + 2^15
#else
constexpr double kMagic = 0.0;
#endif

// log(|X| * mag + 1) from p4 = 4 |X|^2, natural log (wav_util.py:76,107,111), fp32: lg2.approx
// returns exponent + log2(mantissa) with an absolute error of 2^-22 on the mantissa part, i.e. the
// result is as good as its own float32 rounding (ulp 1e-6 at log|X| ~ 15; tolerance 1e-4 max(|ref|,1)).
__device__ __forceinline__ float log_mag(float p4, float half_mag) {
    float r, l2;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p4));
    const float v = fmaf(r, half_mag, 1.0f);
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(v));
    return l2 * 0.69314718056f;
}

__device__ __forceinline__ void stg4_hint(float* p, float4 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w), "l"(policy)
                 : "memory");
}
// Synchronous staging of the PCM of one sub-tile (unaligned utterances, and the noise mix:
// fl32(signal + fl32(K * noise)) is formed on the way in).
template <bool F32>
__device__ __forceinline__ void load_pcm_tile(const Params& p, const Meta& m, int frame0, uint32_t* dst, int hth) {
    const long long t0 = (long long)frame0 * kHop;   // utterance-local first sample of the sub-tile
    if (!F32) {
        const short* src = reinterpret_cast<const short*>(p.samples) + m.sbase;
        constexpr int kChunks = kSubHopRows * 20;   // 16-byte chunks of 8 samples
        for (int c = hth; c < kChunks; c += kTeamThreads) {
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
        constexpr int kChunks = kSubHopRows * 40;   // 16-byte chunks of 4 samples
        for (int c = hth; c < kChunks; c += kTeamThreads) {
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
            *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// Asynchronous staging of one sub-tile (no arithmetic on the way): 16-byte LDGSTS
// copies into the padded hop rows, zero-filled past the end of the utterance.
// Needs the utterance start to be 16-byte aligned.
template <bool F32>
__device__ __forceinline__ void issue_pcm_tile_async(const Params& p, const Meta& m, int frame0, uint32_t* dst,
                                                     int hth) {
    const long long t0 = (long long)frame0 * kHop;
    constexpr int kPerChunk = F32 ? 4 : 8;             // samples per 16 bytes
    constexpr int kChunksPerHop = kHop / kPerChunk;    // 40 / 20
    constexpr int kChunks = kSubHopRows * kChunksPerHop;
    constexpr int kBytes = F32 ? 4 : 2;
    constexpr int kRowWords = F32 ? kHopWordsF32 : kHopWordsI16;
    const char* src = reinterpret_cast<const char*>(p.samples) + m.sbase * kBytes;
    for (int c = hth; c < kChunks; c += kTeamThreads) {
        const long long us = t0 + (long long)c * kPerChunk;
        const long long rem = (m.nsamp - us) * kBytes;
        const int nb = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
        cp_async16(dst + (c / kChunksPerHop) * kRowWords + (c % kChunksPerHop) * 4,
                   nb ? src + us * kBytes : src, nb);
    }
}

// unit -> utterance: the last b with tile_off[b] <= unit that owns units
__device__ __forceinline__ int find_utterance(const int* tile_off, int batch, int tile) {
    int lo = 0, hi = batch - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_off[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    while (tile >= tile_off[lo + 1]) ++lo;          // skip utterances without units
    return lo;
}

// Unit -> utterance and frame range (binary search in shared memory) by the team's first thread, one
// unit ahead; the utterance's constants follow as 8-byte cp.async copies (L2 hits) that nobody waits
// for before the team's next "PCM is in place" barrier.
__device__ __forceinline__ void fill_meta(const Params& p, const int* tile_off, int tile, Meta& m) {
    const int b = find_utterance(tile_off, p.batch, tile);
    cp_async8(&m.fo, p.frame_offsets + b, 8);
    cp_async8(&m.fo1, p.frame_offsets + b + 1, 8);
    cp_async8(&m.sbase, p.sample_offsets + b, 8);
    cp_async8(&m.nsamp, p.sample_counts + b, 8);
    if (p.out_row_offsets) cp_async8(&m.row0, p.out_row_offsets + b, 8);
    if (p.noise && p.gain) cp_async4(&m.gain, p.gain + b, 4);
    else m.gain = 0.0f;
    m.valid = 1;
    m.tile = tile;
    m.b = b;
    m.ntiles_b = tile_off[b + 1] - tile_off[b];
    m.f0 = (tile - tile_off[b]) * kUnit;
}

// mean and 1/std of utterance b from its units' partial sums, by the 160 threads of one team
// (sklearn.preprocessing.scale, wav_util.py:79: std with ddof = 0, std < 10 eps -> 1)
__device__ __forceinline__ void finalize_stats(const Params& p, const int* tile_off, int b, int tt) {
    const int t_lo = tile_off[b], t_hi = tile_off[b + 1];
    const long long nfr = p.frame_offsets[b + 1] - p.frame_offsets[b];
    for (int k = tt; k < kBins; k += kTeamThreads) {
        double a1 = 0.0, a2 = 0.0;
#pragma unroll 4
        for (int q = t_lo; q < t_hi; ++q) {               // fixed order: reproducible
            const double2 v = __ldcg(p.partials + (size_t)q * kBins + k);
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
        __stcg(st + k, mh);
        __stcg(st + kBins + k, (float)(mean - (double)mh));
        __stcg(st + 2 * kBins + k, (float)(1.0 / sd));
    }
}

template <bool F32, int kTeams, bool kSepOut>
struct Cfg {
    static constexpr int kPcmWords = kSubHopRows * (F32 ? kHopWordsF32 : kHopWordsI16);
    static constexpr int kExchBytes = 200 * kSub * 16;                       // 51 200
    static constexpr int kOutBytes = kSepOut ? kSub * kOutStride * 4 : 0;    // 13 056
    static constexpr int kTeamBytes = kExchBytes + kPcmWords * 4 + kOutBytes;
    static constexpr int kThreads = kTeams * kTeamThreads;
    static constexpr size_t smem_bytes() {
        return (size_t)(200 * 32 + 3200 + 3200) + sizeof(int) * (kMaxBatch + 1) + (size_t)kTeams * kTeamBytes + 16;
    }
};

__device__ __forceinline__ void team_bar(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(kTeamThreads) : "memory");
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
template <bool F32, int kTeams, bool kSepOut>
// (registers are allocated per warp in units of 512: 480 threads cannot have more than 128 each)
__global__ void __launch_bounds__((Cfg<F32, kTeams, kSepOut>::kThreads), 1) spectrogram_kernel(Params p) {
    using C = Cfg<F32, kTeams, kSepOut>;
    constexpr int kPcmWords = C::kPcmWords;
    constexpr int kThreads = C::kThreads;
    constexpr int kRowWords = F32 ? kHopWordsF32 : kHopWordsI16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tabW = reinterpret_cast<double*>(smem_raw);                       // [200] x (w0, w1, c0, c1)
    cplx16* tabTall = reinterpret_cast<cplx16*>(tabW + 800);                  // [10][20]
    cplx16* tabP = tabTall + 200;                                             // [200]
    int* tile_off = reinterpret_cast<int*>(tabP + 200);                       // [kMaxBatch + 1]
    unsigned char* team_base = reinterpret_cast<unsigned char*>(tile_off + kMaxBatch + 1);
    __shared__ Meta meta[kTeams][2];
    __shared__ int s_fin[kTeams];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < 400; i += kThreads) {
        const double w = g_tab[i];
        const int m = i >> 1, e = i & 1;
        tabW[4 * m + e] = w;
        tabW[4 * m + 2 + e] = -kMagic * w;
    }
    for (int i = tid; i < 800; i += kThreads) reinterpret_cast<double*>(tabTall)[i] = g_tab[400 + i];
    if (warp == 0) {
        // exclusive scan of ceil(n_frames / kUnit) over the utterances
        int carry = 0;
        for (int base = 0; base < p.batch; base += 32) {
            const int i = base + lane;
            int v = 0;
            if (i < p.batch) {
                long long n = p.frame_offsets[i + 1] - p.frame_offsets[i];
                if (n < 0) n = 0;
                v = (int)((n + kUnit - 1) / kUnit);
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
    const bool zin = want_stats && p.zscore_in_kernel;     // statistics finished inside the kernel
    if (blockIdx.x == 0 && want_stats)
        for (int i = tid; i <= p.batch; i += kThreads) p.tile_off_g[i] = tile_off[i];

    const int team = warp / kTeamWarps, q = warp - team * kTeamWarps;
    const int tt = tid - team * kTeamThreads;
    cplx16* exch = reinterpret_cast<cplx16*>(team_base + (size_t)team * C::kTeamBytes);   // [200][16]
    uint32_t* pcm = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(exch) + C::kExchBytes);
    // out tile [16][204]: its own buffer, or aliased onto the exchange (then two more team barriers guard it)
    float* ot = kSepOut ? reinterpret_cast<float*>(pcm + kPcmWords) : reinterpret_cast<float*>(exch);

    // lane -> (frame of the sub-tile, role): the two half-warps of warp q own roles q, q + 5
    const int f = lane & 15;
    const int r = q + 5 * (lane >> 4);
    const bool j0 = (r == 0);
    const int k1a = lane_k1a(r), k1b = lane_k1b(r);
    const int kb_hi = j0 ? -110 : r;          // bin of slot s >= 6 is kb_hi + 20 s
    const double2* tabW2 = reinterpret_cast<const double2*>(tabW);
    const cplx16* tabT = tabTall + r * 20;
    const bool mix = (p.noise != nullptr);
    // copy-out geometry: lane -> (float4 column 10 q + lane / 3, row phase lane % 3)
    const bool co_act = lane < 30;
    const int co_c4 = 10 * q + (co_act ? lane / 3 : 0);
    const int co_ph = co_act ? lane - 3 * (lane / 3) : kSub;   // idle lanes: no rows

    // unit metadata: claimed and looked up by the team's first thread, one unit ahead
    int next_tile = total_tiles;
    if (tt == 0) next_tile = atomicAdd(p.counters, 1);
    auto prepare_meta = [&](int slot) {
        if (tt == 0) {
            Meta& mn = meta[team][slot];
            const int tile = next_tile;
            if (tile < total_tiles) {
                next_tile = atomicAdd(p.counters, 1);
                fill_meta(p, tile_off, tile, mn);
            } else {
                mn.valid = 0;
            }
        }
    };
    auto stage = [&](const Meta& m, int frame0) {
        const bool aligned = (((reinterpret_cast<uintptr_t>(p.samples) + m.sbase * (F32 ? 4 : 2)) & 15) == 0);
        if (!mix && aligned) issue_pcm_tile_async<F32>(p, m, frame0, pcm, tt);
        else load_pcm_tile<F32>(p, m, frame0, pcm, tt);
        cp_async_commit();
    };
    // publication of a finished unit, one unit late (the fence then finds every store of the unit
    // acknowledged): the team's first thread counts the unit on its utterance; whoever completes the
    // utterance makes its statistics (whole team)
    int pub_b = -1, pub_nt = 0;      // (thread 0) utterance / unit count of the unit to publish
    auto publish_count = [&]() {     // thread 0, after a team barrier that follows the unit's last store
        if (tt == 0) {
            int fin = -1;
            if (pub_b >= 0) {
                __threadfence();
                const int old = atomicAdd(p.done + pub_b, 1);
                if (old + 1 == pub_nt) { __threadfence(); fin = pub_b; }
                pub_b = -1;
            }
            s_fin[team] = fin;
        }
    };
    prepare_meta(0);
    if (tt == 0) { cp_async_commit(); cp_async_wait<0>(); }
    team_bar(team);
    if (meta[team][0].valid) stage(meta[team][0], meta[team][0].f0);
    // the un-normalised rows are read again by the z-score: keep them in L2
    const uint64_t keep = want_stats ? l2_policy_evict_last() : l2_policy_evict_first();
    for (int u = 0;; ++u) {
        const Meta& m = meta[team][u & 1];
        if (!m.valid) break;
        const Meta& mnext = meta[team][(u + 1) & 1];
        cp_async_wait<0>();
        team_bar(team);                         // P: the first sub-tile's PCM and the unit's constants are in place
        const long long m_nfr = m.fo1 - m.fo;
        const int m_nf = (m_nfr - m.f0) < kUnit ? (int)(m_nfr - m.f0) : kUnit;
        const long long m_row0 = p.out_row_offsets ? m.row0 : m.fo;
        const float hm = 0.5f * ((p.mode == ASRK_SPEC_ASRT) ? (1.0f / (float)m.nsamp) : 1.0f);
        const int nsub = (m_nf + kSub - 1) / kSub;
        // z-score sums of this lane's four bins over the unit, shifted by the unit's first row
        float4 cS = make_float4(0.f, 0.f, 0.f, 0.f), sS = cS, qS = cS;
        int fin = -1;
        for (int sub = 0; sub < nsub; ++sub) {
            if (sub > 0) {
                cp_async_wait<0>();
                team_bar(team);                 // P: the sub-tile's PCM is in place; the previous copy-out is over
            } else {
                prepare_meta((u + 1) & 1);      // (flags visible to the team after the next barrier)
                if (zin) publish_count();
                // a one-sub-tile unit stages its successor's PCM right behind pass 1: the constants must be there
                if (nsub == 1 && tt == 0) { cp_async_commit(); cp_async_wait<0>(); }
            }
            // ---------------- pass 1: window, DFT20 of residue r, twiddle ----------------
            {
                cplx z[20], y[20];
#pragma unroll
                for (int n1 = 0; n1 < 20; ++n1) {
                    const int hq = n1 / 8;                // hop row offset of sample 2*(10 n1 + r)
                    const int wq = 10 * (n1 % 8) + r;     // sample-pair index inside the hop
                    const double2 w2 = tabW2[2 * (10 * n1 + r)];
                    if (!F32) {
                        const uint32_t w = pcm[(f + hq) * kRowWords + wq];
#if ASRK_SPEC_CONV == 0
                        const double2 c2 = tabW2[2 * (10 * n1 + r) + 1];
                        const uint32_t wb = w ^ 0x80008000u;
                        const double d0 = __hiloint2double((int)__byte_perm(wb, 0x41300000u, 0x7610), 0);
                        const double d1 = __hiloint2double((int)__byte_perm(wb, 0x41300000u, 0x7632), 0);
                        // wav_util.py:71 data_line * w:  x w = (M + x) w - M w, one DFMA per sample
                        z[n1] = cplx{fma(d0, w2.x, c2.x), fma(d1, w2.y, c2.y)};
#else
                        z[n1] = cplx{(double)(int)(short)(w & 0xffffu) * w2.x, (double)((int)w >> 16) * w2.y};
#endif
                    } else {
                        const float2 v = *reinterpret_cast<const float2*>(
                            reinterpret_cast<const float*>(pcm) + (f + hq) * kRowWords + 2 * wq);
                        z[n1] = cplx{(double)v.x * w2.x, (double)v.y * w2.y};
                    }
                }
                dft20(z, y);
                exch[r * kSub + f] = cplx16{y[0].x, y[0].y};
#if ASRK_SPEC_TWGEN
                // twiddles W200^(r k1) by recurrence from W200^r: 72 fp64 instructions instead of 18 more
                // 16-byte table broadcasts (the kernel is shared-memory-wavefront bound before it is fp64 bound)
                const cplx16 t1l = tabT[1];
                const cplx t1{t1l.x, t1l.y};
                cplx tk = t1;
#pragma unroll
                for (int k1 = 1; k1 < 20; ++k1) {
                    const cplx v = cmul(y[k1], tk);
                    exch[(k1 * 10 + r) * kSub + f] = cplx16{v.x, v.y};
                    if (k1 < 19) tk = cmul(tk, t1);
                }
#else
#pragma unroll
                for (int k1 = 1; k1 < 20; ++k1) {
                    const cplx16 t = tabT[k1];
                    const cplx v = cmul(y[k1], cplx{t.x, t.y});
                    exch[(k1 * 10 + r) * kSub + f] = cplx16{v.x, v.y};
                }
#endif
            }
            team_bar(team);                     // A: pass-1 stores -> pass-2 loads; the PCM has been read
            if (zin && sub == 0) fin = s_fin[team];
            // the next sub-tile's PCM arrives behind the arithmetic
            if (sub + 1 < nsub) stage(m, m.f0 + (sub + 1) * kSub);
            else if (mnext.valid) stage(mnext, mnext.f0);
            if (fin >= 0 && sub == 0) finalize_stats(p, tile_off, fin, tt);
            // ---------------- pass 2: DFT10 of rows j and 20-j, split, log ----------------
            {
                cplx ia[10], ib[10], za[10], zb[10];
#pragma unroll
                for (int n2 = 0; n2 < 10; ++n2) {
                    const cplx16 v = exch[(k1a * 10 + n2) * kSub + f];
                    ia[n2] = cplx{v.x, v.y};
                }
#pragma unroll
                for (int n2 = 0; n2 < 10; ++n2) {
                    const cplx16 v = exch[(k1b * 10 + n2) * kSub + f];
                    ib[n2] = cplx{v.x, v.y};
                }
                if (!kSepOut) team_bar(team);   // B: every lane holds its rows; the exchange becomes the out tile
                dft10(ia, za);   // za[k2] = Z[k1a + 20 k2]
                dft10(ib, zb);   // zb[k2] = Z[k1b + 20 k2]
                float* orow = ot + f * kOutStride;
                if (q == 0) {
                    // the warp that owns role 0 (rows 0 and 10 are their own mirrors): lane-uniform code
                    // with operand selects, eleven slots
                    auto loadP = [&](int s) {
                        const cplx16 t = tabP[(s < 6 ? r : kb_hi) + 20 * (s < 10 ? s : (j0 ? 10 : 0))];
                        return cplx{t.x, t.y};
                    };
                    auto emit = [&](int s, double pk, double pm) {
                        const int k = (s < 6 ? r : kb_hi) + 20 * s;
                        const float vk = log_mag((float)pk, hm), vm = log_mag((float)pm, hm);
                        if (s < 10 || j0) {
                            orow[200 - k] = vm;         // role 0, slot 0 writes bin "200" into the row padding
                            orow[k] = vk;               // role 0, slot 5: bin 100 from pk, as the last store
                        }
                    };
                    split_lane(j0, za, zb, loadP, emit);
                } else {
#if ASRK_SPEC_PGEN
                    const cplx16 p0l = tabP[r];
                    const cplx p0{p0l.x, p0l.y};
#endif
#pragma unroll
                    for (int s = 0; s < 10; ++s) {
                        const int k = r + 20 * s;
#if ASRK_SPEC_PGEN
                        // P[r + 20 s] = P[r] W20^s with W20^s a compile-time constant (constant-bank operands)
                        constexpr double kC20[10] = {1.0, 0.95105651629515357212, 0.80901699437494742410, 0.58778525229247312917,
                                                     0.30901699437494742410, 0.0, -0.30901699437494742410,
                                                     -0.58778525229247312917, -0.80901699437494742410, -0.95105651629515357212};
                        constexpr double kS20[10] = {0.0, -0.30901699437494742410, -0.58778525229247312917, -0.80901699437494742410,
                                                     -0.95105651629515357212, -1.0, -0.95105651629515357212,
                                                     -0.80901699437494742410, -0.58778525229247312917, -0.30901699437494742410};
                        const cplx t = (s == 0) ? p0 : cmul(p0, cplx{kC20[s], kS20[s]});
#else
                        const cplx16 t = tabP[k];
#endif
                        double pk, pm;
                        split_pair(za[s], zb[9 - s], cplx{t.x, t.y}, pk, pm);
                        orow[200 - k] = log_mag((float)pm, hm);
                        orow[k] = log_mag((float)pk, hm);
                    }
                }
            }
            team_bar(team);                     // C: the out tile is complete
            // ---------------- copy-out (16-byte accesses) + column sums ----------------
            {
                const int rows = (m_nf - sub * kSub) < kSub ? (m_nf - sub * kSub) : kSub;
                float* obase = p.out + (size_t)(m_row0 + m.f0 + sub * kSub) * kBins + 4 * co_c4;
                const float* tbase = ot + 4 * co_c4;
                if (sub == 0 && co_act) cS = *reinterpret_cast<const float4*>(tbase);
#pragma unroll 2
                for (int row = co_ph; row < rows; row += 3) {
                    const float4 y = *reinterpret_cast<const float4*>(tbase + row * kOutStride);
                    stg4_hint(obase + (size_t)row * kBins, y, keep);
                    float d = y.x - cS.x; sS.x += d; qS.x = fmaf(d, d, qS.x);
                    d = y.y - cS.y; sS.y += d; qS.y = fmaf(d, d, qS.y);
                    d = y.z - cS.z; sS.z += d; qS.z = fmaf(d, d, qS.z);
                    d = y.w - cS.w; sS.w += d; qS.w = fmaf(d, d, qS.w);
                }
            }
        }
        if (want_stats) {
            // the three row phases of a column group sit in adjacent lanes: fixed-order sums, then the
            // un-shifted sums of the unit in fp64:  sum y = s + n c,  sum y^2 = q + 2 c s + n c^2
            float sv[4] = {sS.x, sS.y, sS.z, sS.w}, qv[4] = {qS.x, qS.y, qS.z, qS.w};
            const float cv[4] = {cS.x, cS.y, cS.z, cS.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float s1 = __shfl_down_sync(0xffffffffu, sv[e], 1), s2 = __shfl_down_sync(0xffffffffu, sv[e], 2);
                const float q1 = __shfl_down_sync(0xffffffffu, qv[e], 1), q2 = __shfl_down_sync(0xffffffffu, qv[e], 2);
                sv[e] = (sv[e] + s1) + s2;
                qv[e] = (qv[e] + q1) + q2;
            }
            if (co_act && co_ph == 0) {
                const double n = (double)m_nf;
                double2* dst = p.partials + (size_t)m.tile * kBins + 4 * co_c4;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double c = (double)cv[e], s1 = (double)sv[e], q1 = (double)qv[e];
                    __stcg(dst + e, make_double2(fma(n, c, s1), fma(c, fma(n, c, 2.0 * s1), q1)));
                }
            }
            if (tt == 0) { pub_b = m.b; pub_nt = m.ntiles_b; }
        }
    }
    if (zin) {
        // drain: the team's last unit is published here
        team_bar(team);
        publish_count();
        team_bar(team);
        const int fin = s_fin[team];
        if (fin >= 0) finalize_stats(p, tile_off, fin, tt);
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
#pragma unroll 4
        for (int q = t_lo; q < t_hi; ++q) {               // fixed order: reproducible
            const double2 v = __ldcg(p.partials + (size_t)q * kBins + k);
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
__global__ void __launch_bounds__(128) normalize_kernel(Params p) {
    // newest utterances first: their rows are the ones the main kernel's evict-last stores still hold in L2
    const int b = (int)gridDim.y - 1 - (int)blockIdx.y;
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

template <bool F32, int kTeams, bool kSepOut>
static void launch_main(const Params& p, int grid, cudaStream_t stream) {
    using C = Cfg<F32, kTeams, kSepOut>;
    const size_t smem = C::smem_bytes();
    cudaFuncSetAttribute(spectrogram_kernel<F32, kTeams, kSepOut>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spectrogram_kernel<F32, kTeams, kSepOut><<<grid, C::kThreads, smem, stream>>>(p), asrk::note_launch();
}

// Experiment switches (read once): ASRK_SPEC_TEAMS = 2 | 3 teams per CTA, ASRK_SPEC_ZSCORE = kernel | separate
// (kernel: mean / 1/std finished by the transform kernel; separate: by the statistics kernel of round 1).
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
static int cfg_teams() {
    static int v = env_int("ASRK_SPEC_TEAMS", 3);
    return v == 2 ? 2 : 3;
}
static int cfg_sepout() {      // int16, three teams: out tile in its own buffer (one team barrier less per sub-tile)
    static int v = env_int("ASRK_SPEC_SEPOUT", 0);
    return v;
}
static int cfg_zscore_in_kernel() {
    static int v = [] {
        const char* e = getenv("ASRK_SPEC_ZSCORE");
        return (e && e[0] == 's') ? 0 : 1;
    }();
    return v;
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

    // the z-score runs inside the main kernel when both phases are asked for in one call (the normal
    // case); a harness that times the phases one by one gets the separate kernels
    const bool both = (phases & ASRK_PHASE_SPEC_MAIN) && (phases & ASRK_PHASE_SPEC_NORMALIZE);
    const int teams = cfg_teams();
    // mean / 1/std are finished inside the transform kernel when both phases come in one call (the normal
    // case); a harness that times the phases one by one gets the separate statistics kernel
    const int zin = (mode == ASRK_SPEC_FBANK && both && cfg_zscore_in_kernel()) ? 1 : 0;

    // the kernel locates units through a prefix array in shared memory: at most
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
        p.zscore_in_kernel = zin;
        p.out = out;
        p.counters = reinterpret_cast<int*>(ws + l.counters);
        p.done = reinterpret_cast<int*>(ws + l.done);
        p.tile_off_g = reinterpret_cast<int*>(ws + l.tile_off);
        p.partials = reinterpret_cast<double2*>(ws + l.partials);
        p.stats = reinterpret_cast<float*>(ws + l.stats) + (size_t)b0 * 3 * kBins;
        if (phases & ASRK_PHASE_SPEC_MAIN) {
            // counters (| done when the statistics are finished in the kernel) start from zero
            if (cudaMemsetAsync(ws + l.counters, 0, zin ? l.zero_bytes : sizeof(int) * 4, stream) != cudaSuccess)
                return ASRK_E_CUDA;
            if (sample_dtype == ASRK_DTYPE_I16) {
                if (teams == 2) launch_main<false, 2, true>(p, grid, stream);
                else if (cfg_sepout()) launch_main<false, 3, true>(p, grid, stream);
                else launch_main<false, 3, false>(p, grid, stream);
            } else {
                if (teams == 2) launch_main<true, 2, true>(p, grid, stream);
                else launch_main<true, 3, false>(p, grid, stream);
            }
        }
        if (mode == ASRK_SPEC_FBANK && (phases & ASRK_PHASE_SPEC_NORMALIZE)) {
            if (!zin) stats_kernel<<<nb, 256, 0, stream>>>(p), asrk::note_launch();
            normalize_kernel<<<dim3(32, nb), 128, 0, stream>>>(p), asrk::note_launch();   // small CTAs: they fit next to the CTC kernel
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
