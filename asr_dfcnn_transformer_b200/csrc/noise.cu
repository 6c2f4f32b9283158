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
// One CTA per utterance.  Eight lanes own the eight interleaved accumulators of
// one leaf block (32-byte coalesced sectors); leaf sums land in a shared-memory
// heap that mirrors the recursion tree and are folded bottom-up.
#include <math.h>

#include "asrk_common.cuh"

namespace asrk {
namespace noise {

constexpr int kThreads = 1024;
constexpr int kMaxDepth = 13;                 // up to 128 * 2^13 = 1,048,576 samples
constexpr int kHeap = 1 << (kMaxDepth + 1);   // heap slots (index 1 = root)
constexpr int kBlock = 128;                   // numpy PW_BLOCKSIZE

__device__ __forceinline__ float leaf_sum8(const float* a, long long len, int j) {
    // lanes j = 0..7 of a group: r[j] = a[j] + a[8+j] + ... ; combined as
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); remainder added one by one (lane 0).
    float res;
    if (len < 8) {
        res = 0.f;
        if (j == 0)
            for (long long i = 0; i < len; ++i) res = __fadd_rn(res, __fmul_rn(a[i], a[i]));
        return res;
    }
    const long long nfull = len - (len % 8);
    float r = __fmul_rn(a[j], a[j]);
    for (long long i = 8; i < nfull; i += 8) {
        const float v = a[i + j];
        r = __fadd_rn(r, __fmul_rn(v, v));
    }
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    if (j == 0)
        for (long long i = nfull; i < len; ++i) r = __fadd_rn(r, __fmul_rn(a[i], a[i]));
    return r;
}

__global__ void __launch_bounds__(kThreads) snr2k_kernel(const float* signal, const float* noise,
                                                          const long long* sample_offsets,
                                                          const long long* sample_counts,
                                                          const int* snr_db, float* gain_out) {
    extern __shared__ float heap[];            // [2][kHeap] values
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
    for (int i = tid; i < kHeap; i += kThreads) present[i] = 0;
    __syncthreads();
    const int group = tid >> 3, j = tid & 7;
    const int n_slots = 1 << D;
    for (int base = 0; base < n_slots; base += kThreads / 8) {
        const int idx = base + group;
        bool valid = idx < n_slots;
        long long start = 0, len = n;
        int node = 1;
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
                node = 2 * node + bit;
            }
        }
        // all 8 lanes of a group share `valid`; groups of one warp may differ, so
        // keep the shuffles inside leaf_sum8 converged by running it for everyone
        const float es = leaf_sum8(signal + s0 + (valid ? start : 0), valid ? len : 8, j);
        const float en = leaf_sum8(noise + s0 + (valid ? start : 0), valid ? len : 8, j);
        if (valid && j == 0) {
            heap[node] = es;
            heap[kHeap + node] = en;
            present[node] = 1;
        }
    }
    __syncthreads();
    for (int level = D - 1; level >= 0; --level) {
        const int first = 1 << level;
        for (int i = first + tid; i < 2 * first; i += kThreads) {
            if (present[2 * i] && present[2 * i + 1]) {
                heap[i] = __fadd_rn(heap[2 * i], heap[2 * i + 1]);
                heap[kHeap + i] = __fadd_rn(heap[kHeap + 2 * i], heap[kHeap + 2 * i + 1]);
                present[i] = 1;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        const float fn = (float)n;
        const float es = __fdiv_rn(heap[1], fn);
        const float en = __fdiv_rn(heap[kHeap + 1], fn);
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
    const size_t smem = sizeof(float) * 2 * kHeap;
    cudaFuncSetAttribute(snr2k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    snr2k_kernel<<<batch, kThreads, smem, stream>>>(signal, noise, sample_offsets, sample_counts,
                                                       snr_db, gain_out), asrk::note_launch();
    return asrk::launch_status();
}
