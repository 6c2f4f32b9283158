// Steps right after the hot path:
//   * low-frame-rate stacking of the features   -- util/utils.py:7-31 build_LFR_features
//     (used by lm_and_am/data_loader2.py:132 and end2end/data_loader.py:286,307 with m=4, n=3)
//   * label error of the greedy decode          -- tf.edit_distance(decoded, sparse_labels)
//     (normalize=True) as in lm_and_am/model/acoustic_model2.py:72-73
// Both are memory / latency trivial next to the feature and CTC kernels; they exist so that
// the data does not have to leave the device between the kernels of the path.
#include "asrk_common.cuh"

namespace asrk {
namespace post {

// out row i of utterance b = concat_k in[min(i n + k, T_b - 1)], k < m: the tail is padded by
// repeating the LAST frame (utils.py:26-30).  One thread per 16 bytes of output, grid-stride.
__global__ void __launch_bounds__(256) lfr_kernel(const float* in, const long long* in_offsets, float* out,
                                                  const long long* out_offsets, int batch, int dim, int m, int n) {
    const int d4 = dim >> 2;                       // float4 per input row
    const long long row4 = (long long)m * d4;      // float4 per output row
    const long long total4 = out_offsets[batch] * row4;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const long long orow = i / row4;
        const int within = (int)(i - orow * row4);
        const int k = within / d4, c = within - k * d4;
        // utterance of the output row: binary search in out_offsets
        int lo = 0, hi = batch - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (out_offsets[mid] <= orow) lo = mid; else hi = mid - 1;
        }
        const long long T = in_offsets[lo + 1] - in_offsets[lo];
        long long src = (orow - out_offsets[lo]) * n + k;
        if (src > T - 1) src = T - 1;
        out4[i] = in4[(in_offsets[lo] + src) * d4 + c];
    }
}

// Levenshtein distance between hyp[b][0..hyp_len[b]) and truth[b][0..truth_len[b]), one warp per
// utterance: lane j owns truth positions 2 j and 2 j + 1 (truth_len <= 64), the hypothesis is
// walked token by token, and the in-row dependency D[i][j] = min_k<=j (tmp[k] + j - k) is a
// warp prefix-minimum of tmp[k] - k.  normalize: distance / truth_len (tf.edit_distance: an
// empty truth gives +inf for a non-empty hypothesis and 0 for an empty one).
__global__ void __launch_bounds__(128) edit_distance_kernel(const int* hyp, int hyp_stride, const int* hyp_len,
                                                            const int* truth, int truth_stride,
                                                            const int* truth_len, int batch, int normalize,
                                                            float* out) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const int H = hyp_len[b], L = truth_len[b];
    const int* h = hyp + (size_t)b * hyp_stride;
    const int* t = truth + (size_t)b * truth_stride;
    const int j0 = 2 * lane, j1 = 2 * lane + 1;          // truth positions (0-based), columns j+1 of D
    const int t0 = (j0 < L) ? t[j0] : -1, t1 = (j1 < L) ? t[j1] : -1;
    const int kBig = 1 << 28;
    // D[0][j+1] = j + 1
    int d0 = (j0 < L) ? j0 + 1 : kBig, d1 = (j1 < L) ? j1 + 1 : kBig;
    for (int i = 0; i < H; ++i) {
        const int c = h[i];
        // diagonal neighbour of column j0+1 is column j0 of the previous row: the left lane's d1 (lane 0: D[i][0] = i)
        int diag0 = __shfl_up_sync(0xffffffffu, d1, 1);
        if (lane == 0) diag0 = i;
        const int diag1 = d0;
        int tmp0 = min(d0 + 1, diag0 + (c == t0 ? 0 : 1));
        int tmp1 = min(d1 + 1, diag1 + (c == t1 ? 0 : 1));
        if (j0 >= L) tmp0 = kBig;
        if (j1 >= L) tmp1 = kBig;
        // insertion chain: D[i+1][j+1] = j + 1 + min( min_{k<=j} (tmp[k] - (k + 1)), D[i+1][0] - 0 )
        int v0 = tmp0 - (j0 + 1), v1 = tmp1 - (j1 + 1);
        int pm = min(v0, v1);                              // per-lane minimum, then an inclusive warp prefix-min
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int nb = __shfl_up_sync(0xffffffffu, pm, o);
            if (lane >= o) pm = min(pm, nb);
        }
        int before = __shfl_up_sync(0xffffffffu, pm, 1);   // prefix-min over the lanes to the left
        if (lane == 0) before = kBig;
        before = min(before, i + 1);                       // column 0 of the new row: D[i+1][0] = i + 1
        const int p0 = min(before, v0);
        const int p1 = min(p0, v1);
        d0 = (j0 < L) ? p0 + (j0 + 1) : kBig;
        d1 = (j1 < L) ? p1 + (j1 + 1) : kBig;
    }
    // D[H][L]
    int res;
    if (L == 0) {
        res = H;
    } else {
        const int src = (L - 1) >> 1;
        const int a = __shfl_sync(0xffffffffu, d0, src), c = __shfl_sync(0xffffffffu, d1, src);
        res = ((L - 1) & 1) ? c : a;
    }
    if (lane == 0) {
        float r = (float)res;
        if (normalize) r = (L > 0) ? r / (float)L : (H > 0 ? __int_as_float(0x7f800000) : 0.f);
        out[b] = r;
    }
}

}  // namespace post
}  // namespace asrk

using namespace asrk;

extern "C" int asrk_lfr_run(const float* in, const long long* in_offsets, float* out, const long long* out_offsets,
                            int batch, int dim, int m, int n, long long total_out_rows, asrk_stream_t stream_) {
    if (batch < 0 || dim < 1 || m < 1 || n < 1 || total_out_rows < 0) return ASRK_E_BADARG;
    if (batch == 0 || total_out_rows == 0) return ASRK_OK;
    if (!in || !in_offsets || !out || !out_offsets) return ASRK_E_BADARG;
    if (dim % 4 != 0) return ASRK_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return ASRK_E_ALIGN;
    const long long total4 = total_out_rows * m * (dim / 4);
    long long blocks = (total4 + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    post::lfr_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
        in, in_offsets, out, out_offsets, batch, dim, m, n), asrk::note_launch();
    return launch_status();
}

extern "C" int asrk_edit_distance_run(const int* hyp, int hyp_stride, const int* hyp_len, const int* truth,
                                      int truth_stride, const int* truth_len, int batch, int normalize,
                                      float* out, asrk_stream_t stream_) {
    if (batch < 0 || hyp_stride < 0 || truth_stride < 0) return ASRK_E_BADARG;
    if (batch == 0) return ASRK_OK;
    if (!hyp_len || !truth_len || !out || (hyp_stride > 0 && !hyp) || (truth_stride > 0 && !truth)) return ASRK_E_BADARG;
    if (truth_stride > 64) return ASRK_E_SHAPE;          // two truth positions per lane (data_loader.py: 64 labels)
    post::edit_distance_kernel<<<(batch + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
        hyp, hyp_stride, hyp_len, truth, truth_stride, truth_len, batch, normalize, out), asrk::note_launch();
    return launch_status();
}
