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
//   * ONE persistent kernel, one CTA per SM, THREE independent teams of five warps.
//     A team transforms 16 consecutive frames at a time: lane = (frame, role half), the
//     two half-warps of warp q own roles q and q + 5 of the 20 x 10 Cooley-Tukey split
//     of the 200-point complex DFT (400-point real DFT = 200-point complex DFT + real-
//     input split; both factors twiddle-free prime-factor codelets), so window values
//     and twiddles are half-warp broadcasts from shared memory and the two passes
//     exchange through the team's own conflict-free 50 KB fp64 buffer with team-wide
//     named barriers only.  15 transform warps sit 4/4/4/3 on the four schedulers (a
//     single 10-role team sits 3/3/2/2 and leaves a sixth of the fp64 pipe idle), and
//     while one team loads, exchanges or copies out, the other two keep the pipe busy.
//   * a team claims units of 32 consecutive frames of one utterance from an atomic
//     counter (in utterance order), stages their PCM with 16-byte cp.async while the
//     previous unit's last sub-tile is still in flight (zero-filled past the end; the
//     noise mix fl32(s + fl32(K n)) is applied on the way in), and copies finished
//     16 x 200 log-magnitude tiles to global memory column-wise: warp q owns bins
//     40 q .. 40 q + 39, so every store is a full 32-byte sector and every bin's
//     z-score sums live in exactly one lane.
//   * z-score: the per-bin sums of every unit are accumulated during the copy-out
//     (fp32, shifted by the unit's first row so that nothing cancels) and written
//     un-shifted in fp64 per unit (fixed order: reproducible, no float atomics).  A tiny kernel turns the tile partials of every utterance into mean and
//     1/std, and a purely streaming kernel with the whole chip's memory parallelism
//     normalises in place (the rows are largely still in L2).  [Measured alternatives:
//     an in-kernel z-score by the CTA that retires an utterance's last tile -- one
//     CTA's 448 threads cannot keep enough L2 requests in flight, 40 us per utterance;
//     per-tile release fence + counter in the helpers -- the membar costs 19 us.]
//
// Round 2: the codelets lost 12 % of their fp64 instructions (asrk_fft.cuh: 32-instruction DFT5, 14-instruction
// split).  A rewrite of this kernel around an "instruction diet" (one-DFMA / I2F sample conversion, 16-byte
// exchange and copy-out accesses, role-0 selects confined to one warp, per-sub-tile PCM staging, z-score inside
// the kernel through TMA bulk copies -- first by a dedicated warp, then spread over the teams) executed 32 %
// fewer instructions, was parity-green, and was SLOWER in every context (A/B in profiles/r2_spectrogram.md):
// the kernel is bound by shared-memory wavefronts and dependent-issue latency, not by issue slots, so this
// layout was kept.
#include <math.h>

#include "asrk_common.cuh"
#include "asrk_fft.cuh"

namespace asrk {
namespace spec {

constexpr int kTeams = 3;
constexpr int kTeamWarps = 5;
constexpr int kTeamThreads = kTeamWarps * 32;        // 160
constexpr int kThreads = kTeams * kTeamThreads;      // 480
constexpr int kSub = 16;                             // frames per sub-tile (= lanes of a half-warp)
constexpr int kHop = 160, kFrameLen = 400, kBins = 200;
constexpr int kOutStride = 201;                      // padded row of the out tile
// hop rows stay 16-byte aligned (for 16-byte async copies) and are padded by 16
// bytes: frame f reads word 84 f + c -> the 16 frames of a half-warp hit 8 distinct
// banks (2-way conflict on the 20 sample loads of a thread per sub-tile).
constexpr int kHopWordsI16 = 84;                     // 80 words of int16 pairs + 4 pad
constexpr int kHopWordsF32 = 164;                    // 160 words + 4 pad
constexpr int kTabDoubles = 1200;                    // window[400] | tw[r][k1] | P[k]
constexpr int kMaxBatch = 2047;                      // utterances per launch (unit prefix in smem)
constexpr int kUttCache = 512;                       // utterances whose constants are cached in smem

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
    const size_t max_tiles = (size_t)(total_frames / kSub) + (size_t)batch + 1;
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
template <bool F32, int kHopRows>
__device__ __forceinline__ void load_pcm_tile(const Params& p, const Meta& m, uint32_t* dst, int hth) {
    constexpr int kHelperThreads = kTeamThreads;
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
template <bool F32, int kHopRows>
__device__ __forceinline__ void issue_pcm_tile_async(const Params& p, const Meta& m, uint32_t* dst, int hth) {
    constexpr int kHelperThreads = kTeamThreads;
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
// per-utterance constants, staged once per CTA when the batch is small enough, so that a
// claim costs a binary search and a few shared-memory reads instead of a global round trip
struct UttCache {
    long long sbase[kUttCache];
    long long nsamp[kUttCache];
    long long row0[kUttCache];
    int nfr[kUttCache];
    float gain[kUttCache];
};

__device__ void fill_meta(const Params& p, const int* tile_off, const UttCache* uc, int tile, int kTile, Meta& m) {
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
    if (uc != nullptr) {
        m.nfr = uc->nfr[b];
        m.sbase = uc->sbase[b];
        m.nsamp = uc->nsamp[b];
        m.row0 = uc->row0[b];
        m.gain = uc->gain[b];
    } else {
        const long long fo = p.frame_offsets[b];
        m.nfr = p.frame_offsets[b + 1] - fo;
        m.sbase = p.sample_offsets[b];
        m.nsamp = p.sample_counts[b];
        m.row0 = p.out_row_offsets ? p.out_row_offsets[b] : fo;
        m.gain = (p.noise && p.gain) ? p.gain[b] : 0.0f;
    }
    const long long rem = m.nfr - m.f0;
    m.nf = rem < kTile ? (int)rem : kTile;
    m.half_mag = 0.5f * ((p.mode == ASRK_SPEC_ASRT) ? (1.0f / (float)m.nsamp) : 1.0f);
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
template <bool F32>
struct Cfg {
    // frames per claimed unit: 32 (two sub-tiles) for int16; 16 for float32 so that the
    // staging buffers of three teams still fit next to the exchange buffers
    static constexpr int kUnit = F32 ? 16 : 32;
    static constexpr int kHopRows = kUnit + 2;                   // (kUnit-1)*160+400 samples
    static constexpr int kPcmWords = kHopRows * (F32 ? kHopWordsF32 : kHopWordsI16);
};

__device__ __forceinline__ void team_bar(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(kTeamThreads) : "memory");
}

template <bool F32>
__global__ void __launch_bounds__(kThreads, 1) spectrogram_kernel(Params p) {
    constexpr int kUnit = Cfg<F32>::kUnit;
    constexpr int kHopRows = Cfg<F32>::kHopRows;
    constexpr int kPcmWords = Cfg<F32>::kPcmWords;
    constexpr int kExchBytes = 200 * kSub * (int)sizeof(cplx);                // 51 200
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);                        // [1200]
    int* tile_off = reinterpret_cast<int*>(tab + kTabDoubles);                // [kMaxBatch + 1]
    UttCache* ucache = reinterpret_cast<UttCache*>(tile_off + kMaxBatch + 1);
    unsigned char* team_base = reinterpret_cast<unsigned char*>(ucache + 1);
    __shared__ Meta meta[kTeams][2];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int team = warp / kTeamWarps, q = warp - team * kTeamWarps;
    const int tt = tid - team * kTeamThreads;
    cplx* exch = reinterpret_cast<cplx*>(team_base + (size_t)team * (kExchBytes + kPcmWords * 4));   // [200][16]
    uint32_t* pcm = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(exch) + kExchBytes);
    float* ot = reinterpret_cast<float*>(exch);                               // out tile [16][201] aliases the exchange

    for (int i = tid; i < kTabDoubles; i += kThreads) tab[i] = g_tab[i];
    const UttCache* uc = (p.batch <= kUttCache) ? ucache : nullptr;
    if (uc != nullptr) {
        for (int b = tid; b < p.batch; b += kThreads) {
            const long long fo = p.frame_offsets[b];
            ucache->nfr[b] = (int)(p.frame_offsets[b + 1] - fo);
            ucache->sbase[b] = p.sample_offsets[b];
            ucache->nsamp[b] = p.sample_counts[b];
            ucache->row0[b] = p.out_row_offsets ? p.out_row_offsets[b] : fo;
            ucache->gain[b] = (p.noise && p.gain) ? p.gain[b] : 0.0f;
        }
    }
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
    if (blockIdx.x == 0 && want_stats)
        for (int i = tid; i <= p.batch; i += kThreads) p.tile_off_g[i] = tile_off[i];

    // lane -> (frame of the sub-tile, role): the two half-warps of warp q own roles q, q + 5
    const int f = lane & 15;
    const int r = q + 5 * (lane >> 4);
    const bool j0 = (r == 0);
    const int k1a = lane_k1a(r), k1b = lane_k1b(r);
    const int kb_hi = j0 ? -110 : r;          // bin of slot s >= 6 is kb_hi + 20 s
    const double2* tabW2 = reinterpret_cast<const double2*>(tab);
    const cplx* tabT = reinterpret_cast<const cplx*>(tab + 400) + r * 20;
    const cplx* tabP = reinterpret_cast<const cplx*>(tab + 800);
    const bool mix = (p.noise != nullptr);
    // the un-normalised rows are read again by the z-score kernel: keep them in L2
    const uint64_t keep = want_stats ? l2_policy_evict_last() : l2_policy_evict_first();

    // unit metadata: claimed and looked up by the team's first thread, one unit ahead
    int next_tile = total_tiles;
    if (tt == 0) next_tile = atomicAdd(p.counters, 1);
    auto prepare_meta = [&](int slot) {
        if (tt == 0) {
            Meta& mn = meta[team][slot];
            const int tile = next_tile;
            if (tile < total_tiles) {
                next_tile = atomicAdd(p.counters, 1);
                fill_meta(p, tile_off, uc, tile, kUnit, mn);
            } else {
                mn.valid = 0;
            }
        }
    };
    auto stage = [&](const Meta& m) {
        const bool aligned = (((reinterpret_cast<uintptr_t>(p.samples) + m.sbase * (F32 ? 4 : 2)) & 15) == 0);
        if (!mix && aligned) issue_pcm_tile_async<F32, kHopRows>(p, m, pcm, tt);
        else load_pcm_tile<F32, kHopRows>(p, m, pcm, tt);
        cp_async_commit();
    };

    prepare_meta(0);
    team_bar(team);
    if (meta[team][0].valid) stage(meta[team][0]);
    for (int u = 0;; ++u) {
        const Meta& m = meta[team][u & 1];
        if (!m.valid) break;
        cp_async_wait<0>();
        team_bar(team);                         // the unit's PCM is in place; everyone is done with unit u - 1
        prepare_meta((u + 1) & 1);              // (visible to the team after the next barrier)
        const Meta& mnext = meta[team][(u + 1) & 1];
        const int nsub = (m.nf + kSub - 1) / kSub;
        // z-score sums of this lane's bins over the unit: bin 40 q + lane, and 40 q + 32 + lane (lane < 8)
        float cA = 0.f, sA = 0.f, qA = 0.f, cB = 0.f, sB = 0.f, qB = 0.f;
        for (int sub = 0; sub < nsub; ++sub) {
            // ---------------- pass 1: window, DFT20 of residue r, twiddle ----------------
            {
                cplx z[20], y[20];
                const int F = sub * kSub + f;            // frame inside the unit
#pragma unroll
                for (int n1 = 0; n1 < 20; ++n1) {
                    const int hq = n1 / 8;                // hop row offset of sample 2*(10 n1 + r)
                    const int wq = 10 * (n1 % 8) + r;     // int16-pair index inside the hop
                    double x0, x1;
                    if (!F32) {
                        const uint32_t w = pcm[(F + hq) * kHopWordsI16 + wq];
                        x0 = i16_to_f64((int)(short)(w & 0xffffu));
                        x1 = i16_to_f64((int)(short)(w >> 16));
                    } else {
                        const float2 v = *reinterpret_cast<const float2*>(
                            reinterpret_cast<const float*>(pcm) + (F + hq) * kHopWordsF32 + 2 * wq);
                        x0 = (double)v.x;
                        x1 = (double)v.y;
                    }
                    const double2 w2 = tabW2[10 * n1 + r];
                    z[n1] = cplx{x0 * w2.x, x1 * w2.y};   // wav_util.py:71 data_line * w
                }
                dft20(z, y);
                exch[r * kSub + f] = y[0];
#pragma unroll
                for (int k1 = 1; k1 < 20; ++k1) exch[(k1 * 10 + r) * kSub + f] = cmul(y[k1], tabT[k1]);
            }
            team_bar(team);                     // ExchA: pass-1 stores -> pass-2 loads; the PCM has been read
            if (sub == nsub - 1 && mnext.valid) stage(mnext);    // next unit's PCM arrives behind the arithmetic
            // ---------------- pass 2: DFT10 of rows j and 20-j, split, log ----------------
            {
                cplx ia[10], ib[10], za[10], zb[10];
#pragma unroll
                for (int n2 = 0; n2 < 10; ++n2) ia[n2] = exch[(k1a * 10 + n2) * kSub + f];
#pragma unroll
                for (int n2 = 0; n2 < 10; ++n2) ib[n2] = exch[(k1b * 10 + n2) * kSub + f];
                team_bar(team);                 // ExchB: every lane holds its rows; the exchange becomes the out tile
                dft10(ia, za);   // za[k2] = Z[k1a + 20 k2]
                dft10(ib, zb);   // zb[k2] = Z[k1b + 20 k2]
                float* orow = ot + f * kOutStride;
                const float hm = m.half_mag;
                auto loadP = [&](int s) { return tabP[(s < 6 ? r : kb_hi) + 20 * (s < 10 ? s : (j0 ? 10 : 0))]; };
                auto emit = [&](int s, double pk, double pm) {
                    const int k = (s < 6 ? r : kb_hi) + 20 * s;
                    const float vk = log_mag((float)pk, hm), vm = log_mag((float)pm, hm);
                    if (s < 10 || j0) {
                        orow[200 - k] = vm;         // role 0, slot 0 writes bin "200" into the row padding
                        orow[k] = vk;               // role 0, slot 5: bin 100 from pk, as the last store
                    }
                };
                split_lane(j0, za, zb, loadP, emit);
            }
            team_bar(team);                     // the out tile is complete
            // ---------------- copy-out by columns + column sums ----------------
            {
                const int rows = (m.nf - sub * kSub) < kSub ? (m.nf - sub * kSub) : kSub;
                float* obase = p.out + (size_t)(m.row0 + m.f0 + sub * kSub) * kBins;
                const int colA = 40 * q + lane, colB = 40 * q + 32 + lane;
                const bool hasB = lane < 8;
#pragma unroll 4
                for (int row = 0; row < rows; ++row) {
                    const float yA = ot[row * kOutStride + colA];
                    const float yB = hasB ? ot[row * kOutStride + colB] : 0.f;
                    stg_hint(obase + (size_t)row * kBins + colA, yA, keep);
                    if (hasB) stg_hint(obase + (size_t)row * kBins + colB, yB, keep);
                    if (sub == 0 && row == 0) { cA = yA; cB = yB; }
                    float d = yA - cA;
                    sA += d;
                    qA = fmaf(d, d, qA);
                    d = yB - cB;
                    sB += d;
                    qB = fmaf(d, d, qB);
                }
            }
            team_bar(team);                     // the out tile has been read: the exchange is free again
        }
        if (want_stats) {
            // un-shifted sums of the unit in fp64:  sum y = s + n c,  sum y^2 = q + 2 c s + n c^2
            const double n = (double)m.nf;
            {
                const double c = (double)cA, s1 = (double)sA, q1 = (double)qA;
                p.partials[(size_t)m.tile * kBins + 40 * q + lane] =
                    make_double2(fma(n, c, s1), fma(c, fma(n, c, 2.0 * s1), q1));
            }
            if (lane < 8) {
                const double c = (double)cB, s1 = (double)sB, q1 = (double)qB;
                p.partials[(size_t)m.tile * kBins + 40 * q + 32 + lane] =
                    make_double2(fma(n, c, s1), fma(c, fma(n, c, 2.0 * s1), q1));
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
// 32 registers per thread on purpose: in the step this kernel runs NEXT TO the fused CTC kernel, whose two CTAs per
// SM leave ~8 K registers and ~11 KB of shared memory; at 48 registers one 128-thread CTA fitted there (8 KB in
// flight per SM, ~1.2 TB/s chip-wide, and 2/3 of the z-score was still to do when the CTC kernel had finished), at
// 32 two do.
__global__ void __launch_bounds__(128, 16) normalize_kernel(Params p) {
    // newest utterances first: their rows are the ones the main kernel's evict-last stores still hold in L2
    const int b = (int)gridDim.y - 1 - (int)blockIdx.y;
    __shared__ __align__(16) float s_stat[3 * kBins];
    for (int k = threadIdx.x; k < 3 * kBins; k += blockDim.x) s_stat[k] = p.stats[(size_t)b * 3 * kBins + k];
    __syncthreads();
    const long long fo = p.frame_offsets[b];
    const int nfr = (int)(p.frame_offsets[b + 1] - fo);
    const long long row0 = p.out_row_offsets ? p.out_row_offsets[b] : fo;
    float4* base = reinterpret_cast<float4*>(p.out + (size_t)row0 * kBins);
    const int n4 = nfr * (kBins / 4);
    const int stride = (int)(gridDim.x * blockDim.x);
    constexpr int kU = 4;
    const uint64_t drop = l2_policy_evict_first();
    for (int i0 = (int)(blockIdx.x * blockDim.x + threadIdx.x); i0 < n4; i0 += kU * stride) {
        float4 v[kU];
#pragma unroll
        for (int e = 0; e < kU; ++e) {
            const int i = i0 + e * stride;
            if (i < n4) v[e] = ldg_hint(base + i, drop);
        }
#pragma unroll
        for (int e = 0; e < kU; ++e) {
            const int i = i0 + e * stride;
            if (i < n4) {
                const int k = (i % (kBins / 4)) * 4;
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
    return sizeof(double) * kTabDoubles + sizeof(int) * (kMaxBatch + 1) + sizeof(UttCache) +
           (size_t)kTeams * (200 * kSub * sizeof(cplx) + sizeof(uint32_t) * Cfg<F32>::kPcmWords) + 16;
}

template <bool F32>
static void launch_main(const Params& p, int grid, cudaStream_t stream) {
    const size_t smem = main_smem_bytes<F32>();
    cudaFuncSetAttribute(spectrogram_kernel<F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spectrogram_kernel<F32><<<grid, kThreads, smem, stream>>>(p), asrk::note_launch();
}

}  // namespace spec
}  // namespace asrk

using namespace asrk;
using namespace asrk::spec;

extern "C" size_t asrk_spectrogram_workspace_bytes(int batch, long long total_frames) {
    if (batch < 0 || total_frames < 0) return 0;
    return ws_layout(batch, total_frames).total;
}

extern "C" int asrk_spectrogram_zscore_handles(void* workspace, size_t workspace_bytes, int batch, long long total_frames,
                                               float** stats, int** ticket) {
    if (!workspace || !stats || !ticket || batch < 1 || total_frames < 0) return ASRK_E_BADARG;
    const WsLayout l = ws_layout(batch, total_frames);
    if (workspace_bytes < l.total) return ASRK_E_WORKSPACE;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    *stats = reinterpret_cast<float*>(ws + l.stats);
    *ticket = reinterpret_cast<int*>(ws + l.counters) + 2;     // two ints, zeroed with the tile counter by the MAIN phase
    return ASRK_OK;
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
    if (!(phases & (ASRK_PHASE_SPEC_MAIN | ASRK_PHASE_SPEC_NORMALIZE | ASRK_PHASE_SPEC_STATS))) return launch_status();

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
        if (mode == ASRK_SPEC_FBANK && !(phases & ASRK_PHASE_SPEC_NORMALIZE) && (phases & ASRK_PHASE_SPEC_STATS)) {
            // statistics only: the rows are normalised by the fused CTC kernel's co-work (asrk_ctc_loss_grad_zscore_run)
            stats_kernel<<<nb, 256, 0, stream>>>(p), asrk::note_launch();
        }
        if (mode == ASRK_SPEC_FBANK && (phases & ASRK_PHASE_SPEC_NORMALIZE)) {
            // (Asking for the largest shared-memory carve-out so that these CTAs can share an SM with the fused CTC
            // kernel's was measured: the z-score alone went from 41 to 54 us -- less L1 -- and the two HBM-bound kernels
            // still overlapped by 18 us only, tools/time_tail.py.  Default carve-out kept.)
            stats_kernel<<<nb, 256, 0, stream>>>(p), asrk::note_launch();
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
