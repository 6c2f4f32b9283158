// Mix gain of the reference's noise augmentation, on the device.
//
// Reference: /root/reference/util/noise.py:48-52
//     energe_s = np.sum(signal * signal) / len(signal)
//     energe_n = np.sum(noise * noise) / len(noise)
//     K = np.sqrt(energe_s / energe_n) * (10 ** (-dB / 20))
// with float32 inputs.  The mixed signal fl32(signal + fl32(K * noise))
// (noise.py:108) feeds a z-scored log spectrogram, where a one-ulp difference in
// K is visible in low-energy bins, so K is reproduced bit for bit: the products
// are rounded to float32 and summed in the exact order of numpy's pairwise
// summation (8 interleaved accumulators over blocks of <= 128 elements, blocks
// combined along the recursion n -> (n/2 rounded down to a multiple of 8, rest)),
// then float32 divide / sqrt / multiply.
//
// One CTA per utterance.  A warp takes sixteen consecutive leaf blocks (<= 2048 samples) per pass: it copies that
// range of both arrays into shared memory with coalesced 16-byte cp.async (16 KB in flight per warp), then TWO lanes
// own one leaf -- lane h holds accumulators 4h .. 4h+3 of numpy's eight -- and the warp folds its 16 leaf sums
// through four levels of the recursion tree with shuffles (siblings are neighbours); what is left of the tree lives
// in a small shared-memory heap, folded bottom-up.
// (Round 1 gave a leaf to eight lanes with 4-byte loads in a dependent loop and kept the whole tree in 144 KB of
// shared memory: one 1024-thread CTA per SM, 1.7 waves, 1.8 TB/s = 181 us per C4 batch.  Round 2's first version
// read the leaves straight from global memory, 16 bytes per lane: every load instruction touched 16 cache lines
// (leaves are ~312 bytes apart) and the kernel was bound by L1 tag look-ups: 116 us.)
#include <math.h>

#include "asrk_common.cuh"

namespace asrk {
namespace noise {

constexpr int kThreads = 128;
constexpr int kMaxDepth = 15;                 // up to 128 * 2^15 = 4,194,304 samples
constexpr int kWarpLevels = 4;                // 16 leaves per warp and pass
constexpr int kHeap = 1 << (kMaxDepth - kWarpLevels + 1);   // heap slots above the warps' sub-trees (index 1 = root)
constexpr int kBlock = 128;                   // numpy PW_BLOCKSIZE
constexpr int kStage = 16 * kBlock;           // floats per warp and array: its sixteen leaves

// sum of squares of a[0 .. len) in numpy's leaf order (len <= 128), for the lane pair (h = 0 / 1) that owns the leaf;
// the result is valid in lane h == 0.  `a` points into the warp's staged copy of its sixteen leaves (shared memory).
__device__ __forceinline__ float leaf_sum(const float* a, int len, int h) {
    const int ngrp = len >> 3;                 // full groups of eight (<= 16); none: numpy's plain loop from 0
    float r = 0.f;
    if (ngrp > 0) {
        // (leaf starts are multiples of 8 samples from the staged range's start: 16-byte aligned)
        const float4* a4 = reinterpret_cast<const float4*>(a) + h;
        float4 v = a4[0];
        float r0 = __fmul_rn(v.x, v.x), r1 = __fmul_rn(v.y, v.y), r2 = __fmul_rn(v.z, v.z), r3 = __fmul_rn(v.w, v.w);
#pragma unroll 4
        for (int g = 1; g < ngrp; ++g) {
            v = a4[2 * g];
            r0 = __fadd_rn(r0, __fmul_rn(v.x, v.x));
            r1 = __fadd_rn(r1, __fmul_rn(v.y, v.y));
            r2 = __fadd_rn(r2, __fmul_rn(v.z, v.z));
            r3 = __fadd_rn(r3, __fmul_rn(v.w, v.w));
        }
        r = __fadd_rn(__fadd_rn(r0, r1), __fadd_rn(r2, r3));
    }
    // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)); the remainder one by one.  (Every lane of the warp executes the shuffle.)
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    if (h == 0)
        for (int i = ngrp << 3; i < len; ++i) r = __fadd_rn(r, __fmul_rn(a[i], a[i]));
    return r;
}

__global__ void __launch_bounds__(kThreads, 4) snr2k_kernel(const float* signal, const float* noise,
                                                          const long long* sample_offsets,
                                                          const long long* sample_counts,
                                                          const int* snr_db, float* gain_out) {
    __shared__ float heap[2][kHeap];
    __shared__ unsigned char present[kHeap];
    const int b = blockIdx.x;
    const long long s0 = sample_offsets[b];
    const long long n = sample_counts[b];
    const int tid = threadIdx.x;
    if (n <= 0) {
        if (tid == 0) gain_out[b] = 0.f;
        return;
    }
    int D = 0;
    {   // depth so that every leaf has <= 128 elements (the larger half is n - n2)
        long long len = n;
        while (len > kBlock) {
            long long n2 = len / 2;
            n2 -= n2 % 8;
            len = len - n2;
            ++D;
        }
    }
    if (D > kMaxDepth) {   // longer than the shared-memory heap supports: flag it
        if (tid == 0) gain_out[b] = __int_as_float(0x7fc00000);
        return;
    }
    const int wl = D < kWarpLevels ? D : kWarpLevels;      // levels folded inside a warp
    const int Dh = D - wl;                                  // depth of the heap's bottom level
    for (int i = tid; i < (2 << Dh); i += kThreads) present[i] = 0;
    __syncthreads();
    const float* sig = signal + s0;
    const float* noi = noise + s0;
    const bool vec = ((reinterpret_cast<uintptr_t>(sig) | reinterpret_cast<uintptr_t>(noi)) & 15) == 0;
    const int lane = tid & 31;
    extern __shared__ __align__(16) float stage_all[];                        // [warps][2][kStage]
    float* st_s = stage_all + (size_t)(tid >> 5) * 2 * kStage;
    float* st_n = st_s + kStage;
    const int q = lane >> 1, h = lane & 1;      // leaf of the warp, half of the leaf
    const int n_slots = 1 << D;
    for (int base = 0; base < n_slots; base += kThreads / 2) {
        const int idx = base + (tid >> 5) * 16 + q;
        bool valid = idx < n_slots;
        long long start = 0, len = n;
        if (valid) {
            for (int level = 0; level < D; ++level) {
                if (len <= kBlock) {
                    // leaf above the bottom level: only the all-zero suffix names it
                    if ((idx & ((1 << (D - level)) - 1)) != 0) valid = false;
                    break;
                }
                long long n2 = len / 2;
                n2 -= n2 % 8;
                const int bit = (idx >> (D - 1 - level)) & 1;
                if (bit) { start += n2; len -= n2; } else { len = n2; }
            }
        }
        // both lanes of a pair share `valid`; pairs of one warp may differ, so keep the shuffles converged by running
        // the leaf for everyone (an invalid pair reads nothing: len 0)
        // the warp's valid leaves are consecutive in memory: one coalesced copy of [first start, last end) of both
        // arrays into the warp's staging rows (16-byte cp.async; 4-byte when the utterance start is not aligned)
        const long long lo = valid ? start : 0x7fffffffffffffffLL, hi = valid ? start + len : -1;
        long long w0 = lo, w1 = hi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const long long a0 = __shfl_xor_sync(0xffffffffu, w0, o), a1 = __shfl_xor_sync(0xffffffffu, w1, o);
            w0 = a0 < w0 ? a0 : w0;
            w1 = a1 > w1 ? a1 : w1;
        }
        if (w1 > w0) {
            const int cnt = (int)(w1 - w0);                                  // <= 16 * 128 floats
            if (vec) {
                const int n4 = cnt >> 2;                                     // (starts and w0 are multiples of 8)
                for (int i = lane; i < n4; i += 32) {
                    cp_async16(st_s + 4 * i, sig + w0 + 4 * i, 16);
                    cp_async16(st_n + 4 * i, noi + w0 + 4 * i, 16);
                }
                for (int i = (n4 << 2) + lane; i < cnt; i += 32) {
                    cp_async4(st_s + i, sig + w0 + i, 4);
                    cp_async4(st_n + i, noi + w0 + i, 4);
                }
            } else {
                for (int i = lane; i < cnt; i += 32) {
                    cp_async4(st_s + i, sig + w0 + i, 4);
                    cp_async4(st_n + i, noi + w0 + i, 4);
                }
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        const int off = valid ? (int)(start - w0) : 0;
        float es = leaf_sum(st_s + off, valid ? (int)len : 0, h);
        float en = leaf_sum(st_n + off, valid ? (int)len : 0, h);
        __syncwarp();                                                        // the staging rows are free again
        // fold the warp's 16 leaves: the left sibling takes left + right when the right one exists (a leaf above the
        // bottom level sits in the leftmost slot of its sub-tree, the other slots are absent)
        bool here = valid;
        for (int s = 1; s < (1 << wl); s <<= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, es, 2 * s);
            const float on = __shfl_xor_sync(0xffffffffu, en, 2 * s);
            const bool oh = __shfl_xor_sync(0xffffffffu, (int)here, 2 * s) != 0;
            if ((q & s) == 0 && oh) {
                es = __fadd_rn(es, os);
                en = __fadd_rn(en, on);
            }
        }
        if (here && h == 0 && (q & ((1 << wl) - 1)) == 0) {
            const int node = (1 << Dh) + (idx >> wl);
            heap[0][node] = es;
            heap[1][node] = en;
            present[node] = 1;
        }
    }
    __syncthreads();
    for (int level = Dh - 1; level >= 0; --level) {
        const int first = 1 << level;
        for (int i = first + tid; i < 2 * first; i += kThreads) {
            if (present[2 * i]) {
                const bool both = present[2 * i + 1] != 0;
                heap[0][i] = both ? __fadd_rn(heap[0][2 * i], heap[0][2 * i + 1]) : heap[0][2 * i];
                heap[1][i] = both ? __fadd_rn(heap[1][2 * i], heap[1][2 * i + 1]) : heap[1][2 * i];
                present[i] = 1;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        const float fn = (float)n;
        const float es = __fdiv_rn(heap[0][1], fn);
        const float en = __fdiv_rn(heap[1][1], fn);
        const float ratio = __fsqrt_rn(__fdiv_rn(es, en));
        const float factor = (float)pow(10.0, -(double)snr_db[b] / 20.0);
        gain_out[b] = __fmul_rn(ratio, factor);
    }
}

}  // namespace noise
}  // namespace asrk

extern "C" int asrk_snr2k_run(const float* signal, const float* noise, const long long* sample_offsets,
                              const long long* sample_counts, const int* snr_db, int batch,
                              float* gain_out, asrk_stream_t stream_) {
    using namespace asrk::noise;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (batch < 0) return ASRK_E_BADARG;
    if (batch == 0) return ASRK_OK;
    if (!signal || !noise || !sample_offsets || !sample_counts || !snr_db || !gain_out)
        return ASRK_E_BADARG;
    const size_t smem = sizeof(float) * 2 * kStage * (kThreads / 32);
    cudaFuncSetAttribute(snr2k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    snr2k_kernel<<<batch, kThreads, smem, stream>>>(signal, noise, sample_offsets, sample_counts,
                                                       snr_db, gain_out), asrk::note_launch();
    return asrk::launch_status();
}
