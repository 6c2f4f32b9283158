// Parts 2 and 3 of the hot path: CTC loss forward-backward with the gradient
// w.r.t. the logits, and greedy CTC decode.
//
// Reference call sites: K.ctc_batch_cost (lm_and_am/model/cnn_ctc.py:149-152),
// tf.nn.ctc_loss_v2 (lm_and_am/model/acoustic_model2.py:79-80),
// tf.nn.ctc_greedy_decoder (acoustic_model2.py:69, consumed at lm_and_am/test.py:
// 48-52), util/utils.py:57-66 decode_ctc.  The arithmetic is TensorFlow's
// CTCLossOp / CTCGreedyDecoderOp (third-party, restated in oracle/ctc_ref.py).
//
// Launches per call (DESIGN.md "CTC kernels"):
//   prep    : one CTA per utterance -- effective label list (by length / drop zeros),
//             feasibility, chains of repeated labels, and the flag that tells the generic
//             kernels whether any utterance needs them.  When prep and fused are asked for in
//             one call, the fused kernel does this for its own utterance (no launch).
//   fused   : one CTA per utterance whose lattice fits a warp and 60 KB of shared memory
//             (every AISHELL-shaped utterance): rows streamed through per-warp buffers filled
//             by TMA bulk copies on per-warp mbarriers (max / first arg-max / log-sum-exp /
//             gather), alpha and beta on two warps concurrently in the LINEAR domain in fp64
//             with an exact power-of-two rescaling per column, greedy collapse on a third, then
//             the rows again from L2 for softmax minus occupancy, written once.  Logits cross
//             HBM once in, the gradient once out.  A caller that bounds the batch
//             (ASRK_CTC_SMALL_ONLY) gets exactly this one launch.  With ASRK_CTC_INPUT_PROB the
//             input is Keras' softmax output p: log(p + eps) and the gradient w.r.t. p are formed
//             inside.  Optionally (asrk_ctc_loss_grad_zscore_run) the kernel also carries the
//             z-score pass of the feature path as co-work: TMA-fed chunks through the shared
//             memory of CTAs that have finished their utterances.
//   generic path for long lattices (T in the thousands, hundreds of labels), skipped by a
//   device flag when every utterance was fused:
//   rows    : one WARP per (t,b) row, grid-stride -- max, first arg-max, log-sum-exp and the
//             gather of the log-probabilities the lattice needs.  HBM-bound.
//   lattice : grid (B, 2): the alpha sweep and the beta sweep of an utterance run in two CTAs
//             at the same time; log2 domain, one (blank,label) state pair per thread so a step
//             needs ONE neighbour value; running values in fp64 with fp32 transcendentals on the
//             differences, columns stored as float32 relative to a per-column level in double.
//   grad    : one CTA per (utterance, 64 frames), a warp per row -- softmax from the row and the
//             stored log-sum-exp, minus the lattice occupancies 2^(ahat + bhat + levels - log2 p)
//             scattered per class into a shared-memory row first, so that the gradient row is
//             written once with 16-byte stores; zero rows for t >= len.  Repeated labels are
//             summed along precomputed chains in a fixed order (no atomics): bit-reproducible.
//   collapse: greedy decode from the stored arg-max path (warp ballot + prefix count).
#include <math.h>

#include "asrk_common.cuh"

namespace asrk {
namespace ctc {

constexpr int kRowWarps = 8;                 // warps per CTA in the row kernels
constexpr float kNegInf = -INFINITY;

struct Params {
    const float* logits;
    long long stride_t, stride_b;
    int T, B, V;
    const int* labels;
    int label_stride;
    const int* label_len;
    const int* input_len;
    int blank;
    int label_mode;
    const float* grad_scale;
    float* loss;
    float* grad;
    long long gstride_t, gstride_b;
    int* row_status;
    int* tokens;
    int token_stride;
    int* token_len;
    float* neg_sum_logits;
    int merge_repeated;
    // workspace
    int* eff_labels;     // [B][Ls]
    int* eff_len;        // [B]
    int* chain_next;     // [B][Ls]  next position with the same label, -1 at the end
    int* chain_first;    // [B][Ls]  1 if no earlier position has the same label
    float* lse;          // [B][T]
    float* rowmax;       // [B][T]
    int* argmax;         // [B][T]
    float* lpl;          // [B][T][Ls+1]   slot 0 = blank, slot 1+j = label j (log2 of the softmax probability)
    float* occ;          // [B][T][2 Ls+1] log2 alpha relative to the column's level, lattice order
    double* coff;        // [B][T]        alpha levels C_t
    float* beta;         // [B][T][2 Ls+1] log2 beta relative to the column's level
    double* coffb;       // [B][T]        beta levels D_t
    double* logp;        // [B]
    int* need_generic;   // [1] set by prep when some utterance does not fit the fused kernel
    int Ls;              // label_stride
    int fused;           // 1: utterances with a small lattice are done by fused_small_kernel
    int prep_fused;      // 1: fused_small_kernel prepares its own utterance (no prep_kernel launch)
    int small_only;      // 1: the caller bounded the batch (ASRK_CTC_SMALL_ONLY): no generic kernels follow
    int prob;            // 1: `logits` holds PROBABILITIES p (Keras' softmax output); the op's input is log(p + eps)
    float eps;           //    (K.ctc_batch_cost: eps = 1e-7) and `grad` is the gradient w.r.t. p
    // co-work of the fused kernel: the z-score pass of the feature path (feat == nullptr: none)
    struct ZWork {
        float* feat;                      // rows of 200 float32, un-normalised log-spectrogram on entry
        const float* stats;               // [Bz][3][200]: mean (hi, lo), 1/std
        const long long* frame_offsets;   // [Bz + 1]
        const long long* row_offsets;     // [Bz] first output row of every utterance, or nullptr (= frame_offsets)
        int* ticket;                      // [2] next chunk, next utterance (zero on entry)
        int batch;
        int rows;                         // feature rows per chunk
        long long total_frames;
    } z;
};

// Utterances whose lattice fits one warp (L <= 31 labels -> 32 state pairs) and whose
// per-frame scalars + gathered log-probabilities + alpha + beta fit the shared-memory
// budget take the fused CTA-per-utterance kernel.
constexpr int kSmallSmemFloats = 15 * 1024;   // 60 KB per CTA (+ one row buffer per warp)
__host__ __device__ __forceinline__ bool small_lattice(int L, int T) {
    return L <= 31 && (long long)T * (5 * L + 14) <= kSmallSmemFloats;
}

struct WsLayout {
    size_t eff_labels, eff_len, chain_next, chain_first, lse, rowmax, argmax, lpl, occ, coff, beta, coffb, logp, flag, total;
};

static WsLayout ws_layout(int T, int B, int Ls) {
    WsLayout l;
    size_t o = 0;
    const size_t BT = (size_t)B * (size_t)T;
    l.eff_labels = o;  o = align_up(o + sizeof(int) * (size_t)B * Ls, 256);
    l.eff_len = o;     o = align_up(o + sizeof(int) * (size_t)B, 256);
    l.chain_next = o;  o = align_up(o + sizeof(int) * (size_t)B * Ls, 256);
    l.chain_first = o; o = align_up(o + sizeof(int) * (size_t)B * Ls, 256);
    l.lse = o;         o = align_up(o + sizeof(float) * BT, 256);
    l.rowmax = o;      o = align_up(o + sizeof(float) * BT, 256);
    l.argmax = o;      o = align_up(o + sizeof(int) * BT, 256);
    l.lpl = o;         o = align_up(o + sizeof(float) * BT * (size_t)(Ls + 1), 256);
    l.occ = o;         o = align_up(o + sizeof(float) * BT * (size_t)(2 * Ls + 1), 256);
    l.coff = o;        o = align_up(o + sizeof(double) * BT, 256);
    l.beta = o;        o = align_up(o + sizeof(float) * BT * (size_t)(2 * Ls + 1), 256);
    l.coffb = o;       o = align_up(o + sizeof(double) * BT, 256);
    l.logp = o;        o = align_up(o + sizeof(double) * (size_t)B, 256);
    l.flag = o;        o = align_up(o + sizeof(int), 256);
    l.total = o;
    return l;
}

// ---------------------------------------------------------------------------
// prep
// ---------------------------------------------------------------------------
// One CTA per utterance, one thread per label slot (the CTA has >= label_stride threads);
// `sh` holds label_stride + 40 ints.  Every thread returns (row status, effective length).
struct PrepOut {
    int status, L;
};
__device__ __forceinline__ PrepOut prep_body(const Params& p, int b, int* sh) {
    const int j = threadIdx.x;
    const int Ls = p.label_stride;
    int* seff = sh;                        // [Ls] effective labels
    int* scr = sh + (Ls > 0 ? Ls : 1);     // [33] scan scratch, [34] status, [35] L
    const int* lab = p.labels + (size_t)b * p.label_stride;
    int* eff = p.eff_labels + (size_t)b * p.Ls;
    int* nxt = p.chain_next + (size_t)b * p.Ls;
    int* fst = p.chain_first + (size_t)b * p.Ls;
    const int tl = p.input_len[b];
    if (j == 0) scr[34] = (tl < 1 || tl > p.T) ? ASRK_ROW_BAD_LENGTH : ASRK_ROW_OK;
    int Lby = 0;
    if (p.label_mode == ASRK_LABELS_BY_LENGTH) {
        Lby = p.label_len[b];
        if (Lby < 0 || Lby > Ls) Lby = -1;
    }
    __syncthreads();
    if (Lby < 0) {                         // uniform over the CTA
        if (j == 0) { p.eff_len[b] = 0; p.row_status[b] = ASRK_ROW_BAD_LENGTH; }
        return PrepOut{ASRK_ROW_BAD_LENGTH, 0};
    }
    // keep flag: by length (Keras) or every non-zero entry (dense_to_sparse, acoustic_model2.py:71)
    const int v = (j < Ls) ? lab[j] : 0;
    const int keep = (j < Ls) && ((p.label_mode == ASRK_LABELS_BY_LENGTH) ? (j < Lby) : (v != 0));
    // block-wide exclusive scan of keep
    const int lane = j & 31, warp = j >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int within = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) scr[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        int c = (lane < (int)(blockDim.x >> 5)) ? scr[lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        scr[lane] = incl - c;
        if (lane == 31) scr[32] = incl;
    }
    __syncthreads();
    const int L = scr[32];
    if (keep) {
        int c = v;
        if (c < 0 || c >= p.V) { atomicExch(&scr[34], ASRK_ROW_BAD_LENGTH); c = 0; }
        seff[scr[warp] + within] = c;
    }
    __syncthreads();
    int rep = 0;
    if (j < L) {
        const int c = seff[j];
        eff[j] = c;
        rep = (j > 0 && seff[j - 1] == c) ? 1 : 0;
        int n = -1;
        for (int k = j + 1; k < L; ++k)
            if (seff[k] == c) { n = k; break; }
        nxt[j] = n;
        int first = 1;
        for (int k = 0; k < j; ++k)
            if (seff[k] == c) { first = 0; break; }
        fst[j] = first;
    }
    const int repeats = __syncthreads_count(rep);
    if (j == 0) {
        int status = scr[34];
        if (status == ASRK_ROW_OK && tl < L + repeats) status = ASRK_ROW_NOT_ENOUGH_TIME;
        const int Le = (status == ASRK_ROW_BAD_LENGTH) ? 0 : L;
        if (status != ASRK_ROW_BAD_LENGTH && !small_lattice(L, tl < p.T ? tl : p.T)) {
            // a bounded batch (no generic kernels behind this one) must not hold such a row
            if (p.small_only) status = ASRK_ROW_NOT_SMALL;
            else atomicOr(p.need_generic, 1);
        }
        p.eff_len[b] = Le;
        p.row_status[b] = status;
        scr[34] = status;
        scr[35] = Le;
    }
    __syncthreads();
    return PrepOut{scr[34], scr[35]};
}

__global__ void __launch_bounds__(1024) prep_kernel(Params p) {
    extern __shared__ int sh[];
    prep_body(p, blockIdx.x, sh);
}

// ---------------------------------------------------------------------------
// rows: per-row max / argmax / log-sum-exp and gather of the lattice log-probs
// ---------------------------------------------------------------------------
template <int NV4>
struct RowRegs {
    float4 v[NV4];
};

// load one row (V floats, V % 4 == 0, 16-byte aligned) into registers; missing
// tail elements are -inf
template <int NV4>
__device__ __forceinline__ void load_row(const float* row, int V4, int lane, RowRegs<NV4>& r) {
    const float4* x4 = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
        const int i = lane + 32 * k;
        if (i < V4) r.v[k] = __ldg(x4 + i);
        else r.v[k] = make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
    }
}

// the same from a row staged in shared memory
template <int NV4>
__device__ __forceinline__ void lds_row(const float4* rb, int V4, int lane, RowRegs<NV4>& r) {
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
        const int i = lane + 32 * k;
        if (i < V4) r.v[k] = rb[i];
        else r.v[k] = make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
    }
}
// start the copy of one row into the warp's buffer: ONE TMA bulk copy (cp.async.bulk) issued by
// lane 0 and completed on the warp's mbarrier, with an L2 eviction priority (first read: keep the
// row for the gradient pass; second read: drop it)
__device__ __forceinline__ void issue_row(float4* rb, const float* row, int V4, int lane, uint64_t policy,
                                          uint64_t* bar) {
    if (lane == 0) tma_load_1d(rb, row, (unsigned)V4 * 16u, bar, policy);
}
// The bulk copy writes through the async proxy, which is not ordered after generic-proxy accesses of the same
// shared memory: before the buffer is handed back to the TMA, every lane has consumed what it loaded (a register
// dependence on all of its LDS results), issues the cross-proxy fence the architecture specifies for this
// (fence.proxy.async.shared::cta), and the warp converges.  Round 1 relied on the register dependence alone; with
// the shorter lattice phase of round 2 a handful of gradient rows per 10^4 were again overwritten under the reader
// (tests/test_gpu_ctc.py::test_full_size_c2_batch_properties) until the fence went in.
template <int NV4>
__device__ __forceinline__ void release_row(const float4 (&v)[NV4]) {
    unsigned dep = 0;
#pragma unroll
    for (int k = 0; k < NV4; ++k) dep |= __float_as_uint(v[k].w);
    asm volatile("" ::"r"(dep) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
}

// first maximum over the row: strict '>' scan order == lowest index among equals
template <int NV4>
__device__ __forceinline__ void row_argmax(const RowRegs<NV4>& r, int lane, float& m, int& am) {
    m = kNegInf;
    am = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
        const int base = 4 * (lane + 32 * k);
        // NaN never wins a strict '>' comparison
        if (r.v[k].x > m) { m = r.v[k].x; am = base; }
        if (r.v[k].y > m) { m = r.v[k].y; am = base + 1; }
        if (r.v[k].z > m) { m = r.v[k].z; am = base + 2; }
        if (r.v[k].w > m) { m = r.v[k].w; am = base + 3; }
    }
    // per-lane candidates are in increasing index order inside the lane only, so the
    // cross-lane reduction must compare (value desc, index asc)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    if (am == 0x7fffffff) am = 0;   // all -inf / NaN row: TF's scan leaves index 0
}

// one-MUFU exp2 / log2 (flush-to-zero: a softmax term below 2^-126 is zero anyway)
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_fast(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2e = 1.4426950408889634f;

template <int NV4>
__device__ __forceinline__ float row_sumexp(const RowRegs<NV4>& r, float m) {
    const float m2 = -m * kLog2e;            // exp(x - m) = 2^(x log2e - m log2e): one FFMA + one MUFU per term
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
        s += ex2_fast(fmaf(r.v[k].x, kLog2e, m2));
        s += ex2_fast(fmaf(r.v[k].y, kLog2e, m2));
        s += ex2_fast(fmaf(r.v[k].z, kLog2e, m2));
        s += ex2_fast(fmaf(r.v[k].w, kLog2e, m2));
    }
    return warp_sum(s);
}

// PROB input: the op's input is x = log(p + eps), so softmax(x) = (p + eps) / sum(p + eps): no exponentials.
// Sum of (p + eps) over the row (tail lanes of the register image are not part of the row).
template <int NV4>
__device__ __forceinline__ float row_sum_eps(const RowRegs<NV4>& r, int V4, int lane, float eps) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
        if (lane + 32 * k < V4) s += ((r.v[k].x + eps) + (r.v[k].y + eps)) + ((r.v[k].z + eps) + (r.v[k].w + eps));
    }
    return warp_sum(s);
}

// generic (unaligned / V % 4 != 0 / very large V) row pass: two sweeps
__device__ __forceinline__ void row_stats_generic(const float* row, int V, int lane, float& m, int& am,
                                                  float& s) {
    m = kNegInf;
    am = 0x7fffffff;
    for (int i = lane; i < V; i += 32) {
        const float x = row[i];
        if (x > m) { m = x; am = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    if (am == 0x7fffffff) am = 0;
    s = 0.f;
    for (int i = lane; i < V; i += 32) s += __expf(row[i] - m);
    s = warp_sum(s);
}

template <int NV4, bool WANT_LSE>
__device__ __forceinline__ void rows_body(const Params& p, long long row, int lane) {
    const int t = (int)(row / p.B), b = (int)(row % p.B);
    int tl = p.input_len[b];
    if (tl > p.T) tl = p.T;
    if (t >= tl) return;
    if (WANT_LSE && p.fused && small_lattice(p.eff_len[b], tl)) return;   // fused_small_kernel
    const float* x = p.logits + (size_t)t * p.stride_t + (size_t)b * p.stride_b;
    float m, s = 0.f;
    int am;
    const bool prob = WANT_LSE && p.prob;
    if constexpr (NV4 > 0) {
        RowRegs<NV4> r;
        load_row(x, p.V >> 2, lane, r);
        row_argmax(r, lane, m, am);
        if (WANT_LSE) s = prob ? row_sum_eps(r, p.V >> 2, lane, p.eps) : row_sumexp(r, m);
    } else {
        row_stats_generic(x, p.V, lane, m, am, s);
        if (prob) {
            s = 0.f;
            for (int i = lane; i < p.V; i += 32) s += x[i] + p.eps;
            s = warp_sum(s);
        }
    }
    const size_t bt = (size_t)b * p.T + t;
    if (lane == 0) {
        p.rowmax[bt] = m;
        p.argmax[bt] = am;
    }
    if (WANT_LSE) {
        // log of the softmax denominator of the op's input: log sum exp(x), or log sum (p + eps)
        const float lse = prob ? __logf(s) : m + __logf(s);
        if (lane == 0) p.lse[bt] = lse;
        const int L = p.eff_len[b];
        const int* eff = p.eff_labels + (size_t)b * p.Ls;
        float* dst = p.lpl + bt * (size_t)(p.Ls + 1);
        for (int j = lane; j <= L; j += 32) {
            const int c = (j == 0) ? p.blank : eff[j - 1];
            dst[j] = ((prob ? __logf(x[c] + p.eps) : x[c]) - lse) * kLog2e;         // log2 y_t(l'_j)
        }
    }
}

// one warp per (t,b) row, grid-stride (a no-op launch costs a launch, not a sweep)
template <int NV4, bool WANT_LSE>
__global__ void __launch_bounds__(kRowWarps * 32) rows_kernel(Params p) {
    if (WANT_LSE && p.fused && *p.need_generic == 0) return;   // every utterance took the fused kernel
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)p.T * p.B;
    for (long long row = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * kRowWarps)
        rows_body<NV4, WANT_LSE>(p, row, lane);
}

// ---------------------------------------------------------------------------
// lattice
// ---------------------------------------------------------------------------
// log(e^a + e^b [+ e^c]) with the running values in double and the transcendentals
// in fp32: the exp arguments are differences <= 0 and the log argument is in
// [1,3], so the absolute error per step is ~3e-7 whatever the magnitude of the
// running sums (which reach T*log V ~ 1e4 for T in the thousands) -- plain fp32
// log-space would lose 1e-3 there.
__device__ __forceinline__ double lse2(double a, double b) {
    const double m = fmax(a, b);
    if (m == (double)kNegInf) return m;
    const float s = __expf((float)(a - m)) + __expf((float)(b - m));
    return m + (double)__logf(s);
}
__device__ __forceinline__ double lse3(double a, double b, double c) {
    const double m = fmax(fmax(a, b), c);
    if (m == (double)kNegInf) return m;
    const float s = __expf((float)(a - m)) + __expf((float)(b - m)) + __expf((float)(c - m));
    return m + (double)__logf(s);
}

// block-wide max over all threads (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();            // red[] free
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    float r = (lane < nw) ? red[lane] : kNegInf;
    return warp_max(r);
}

// log2(2^a + 2^b [+ 2^c]) in fp32, branch-free, one MUFU per exp / log: the exp
// arguments are <= 0 (flush below -126 is exact enough: 2^-126 of the column's mass), the
// log argument is in [1, 3] where lg2.approx is good to 2^-22 absolute; -inf operands are
// zeros and all -inf stays -inf (m + lg2(0)).
__device__ __forceinline__ float lse2_log2(float a, float b) {
    const float m = fmaxf(a, b);
    const float ms = (m == kNegInf) ? 0.f : m;
    return m + lg2_fast(ex2_fast(a - ms) + ex2_fast(b - ms));
}
__device__ __forceinline__ float lse3_log2(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    const float ms = (m == kNegInf) ? 0.f : m;
    return m + lg2_fast(ex2_fast(a - ms) + ex2_fast(b - ms) + ex2_fast(c - ms));
}
// log2(2^a + 2^b [+ 2^c]) with the running values in DOUBLE and the transcendentals in fp32: the exp arguments are
// differences <= 0 and the log argument is in [1, 3], so the absolute error per step is ~3e-7 whatever the magnitude
// of the running values -- fp32 running values drift: at T = 1998 the states that carry the occupancy sit ~1000
// binades below the column maximum, where a float has 6e-5 of resolution per step.
#ifndef ASRK_LSE_V1
// The step of the generic lattice is bound by the conversion / special-function unit (16 lanes per clock and SM):
// the largest term of the sum is exactly 1, so only the OTHER terms are converted and exponentiated.
__device__ __forceinline__ double lse2_log2d(double a, double b) {
    const double m = fmax(a, b), n = fmin(a, b);
    const double ms = (m == (double)kNegInf) ? 0.0 : m;
    return m + (double)lg2_fast(1.0f + ex2_fast((float)(n - ms)));
}
__device__ __forceinline__ double lse3_log2d(double a, double b, double c) {
    const double hi = fmax(a, b), lo = fmin(a, b);
    const double m = fmax(hi, c), x = fmin(hi, c);        // m the largest; x and lo the other two
    const double ms = (m == (double)kNegInf) ? 0.0 : m;
    return m + (double)lg2_fast((1.0f + ex2_fast((float)(x - ms))) + ex2_fast((float)(lo - ms)));
}
#else
__device__ __forceinline__ double lse2_log2d(double a, double b) {
    const double m = fmax(a, b);
    const double ms = (m == (double)kNegInf) ? 0.0 : m;
    return m + (double)lg2_fast(ex2_fast((float)(a - ms)) + ex2_fast((float)(b - ms)));
}
__device__ __forceinline__ double lse3_log2d(double a, double b, double c) {
    const double m = fmax(fmax(a, b), c);
    const double ms = (m == (double)kNegInf) ? 0.0 : m;
    return m + (double)lg2_fast(ex2_fast((float)(a - ms)) + ex2_fast((float)(b - ms)) + ex2_fast((float)(c - ms)));
}
#endif

// warp maximum of arbitrary-sign floats with ONE REDUX: floats order like the unsigned
// patterns  bits ^ (sign ? 0xffffffff : 0x80000000)
__device__ __forceinline__ float warp_max_redux(float v) {
    const unsigned b = __float_as_uint(v);
    const unsigned key = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    const unsigned m = __reduce_max_sync(0xffffffffu, key);
    return __uint_as_float(m ^ ((m >> 31) ? 0x80000000u : 0xffffffffu));
}

// ---------------------------------------------------------------------------
// fused CTA-per-utterance kernel (small lattices: the AISHELL-shaped case)
// ---------------------------------------------------------------------------
// One CTA owns one utterance from the first read of its logits to the last write of
// its gradient:
//   A  every warp streams rows (one 5.7 KB row per warp, held in registers):
//      max / first arg-max / log-sum-exp, and the few log-probabilities the lattice
//      needs go to shared memory;
//   B  warp 0 runs alpha and warp 1 runs beta AT THE SAME TIME (lane = state pair,
//      one shuffle per step, no block barrier inside the recursion) while warp 2
//      collapses the arg-max path into the greedy decode;
//   C  every warp streams the rows again (L2 hits: the CTA read them microseconds
//      ago) and writes softmax minus occupancy once, 16 bytes per lane.
// The logits cross HBM once in and the gradient once out; alpha, beta and the
// gathered log-probabilities never leave shared memory.
#ifdef ASRK_CTC_TIMING
// SM cycle counter: %globaltimer takes microseconds to read on this part (two adjacent reads were 1.7-5.5 us apart)
#define ASRK_TICK(i) do { if (threadIdx.x == 0) s_tick[i] = (unsigned long long)clock64(); } while (0)
#else
#define ASRK_TICK(i) do { } while (0)
#endif
// The z-score pass of the FEATURE path as co-work of this kernel (asrk_ctc_loss_grad_zscore_run).  The transform
// kernel owns every SM while it runs, so the HBM-bound rest of a step -- 202 MB of z-score traffic, 184 MB of CTC
// traffic -- runs behind it; as two kernels they overlapped by 18 us of 41 + 57 (the z-score CTAs only get the
// registers the CTC CTAs leave), and the CTC kernel alone leaves HBM idle in its lattice phase, on the 40 SMs that
// hold one CTA instead of two, and behind every utterance shorter than the longest.  Here every CTA that has
// finished its utterance -- and the CTAs launched beyond the batch to fill the free slots -- takes tickets for
// chunks of `rows` feature rows (newest first: still in L2) and normalises them in place with the arithmetic of
// spec::normalize_kernel.  What a CTA needs for that is bytes in flight, not threads (a first version with per-thread
// loads, 16 KB in flight per CTA, took 131 us for the merged kernel): a chunk arrives as ONE TMA bulk copy into the
// shared memory the utterance no longer needs, the next three tickets' chunks are in flight while this one is
// normalised (4 x 25.6 KB per CTA for V = 1424; with two buffers a CTA was bound by one bulk copy's latency at
// 26 GB/s), and the rows leave through plain 16-byte evict-first stores.
// Thread (r, q) owns column group q (4 bins) of rows r, r + 5, ...: its statistics live in registers per utterance.
constexpr int kZBins = 200;
constexpr int kZCache = 512;            // utterance starts cached in shared memory (larger batches: global look-ups)
constexpr int kZBufs = 4;               // chunk buffers per CTA: three bulk copies in flight while one chunk is normalised
__device__ __forceinline__ void zscore_cowork(const Params::ZWork& z, float* sm, int budget) {
    __shared__ uint64_t s_zbar[kZBufs];
    __shared__ int s_tk[kZBufs];
    const int tid = threadIdx.x;
    const int rows = z.rows;                                   // rows per chunk (host: what the dynamic smem holds kZBufs times)
    long long* s_fo = reinterpret_cast<long long*>(sm);        // [kZCache]
    float4* buf0 = reinterpret_cast<float4*>(s_fo + kZCache);
    const size_t buf_f4 = (size_t)rows * (kZBins / 4);
    const bool cached = (z.batch + 1 <= kZCache);
    const long long* fo = cached ? s_fo : z.frame_offsets;
    const long long nchunks = (z.total_frames + rows - 1) / rows;
    const uint64_t drop = l2_policy_evict_first();
    __syncthreads();                                           // everybody has left the utterance's shared memory
    if (tid == 0) {
        for (int j = 0; j < kZBufs; ++j) mbar_init(&s_zbar[j], 1);
        mbar_fence_init();
    }
    if (cached)
        for (int i = tid; i <= z.batch; i += blockDim.x) s_fo[i] = z.frame_offsets[i];
    // ticket k of this CTA -> buffer k % kZBufs: chunk c covers frames [total - (c+1) rows, total - c rows)
    // `budget`: tickets this call may take (< 0: until the queue is empty)
    auto issue = [&](int k) {
        const int c = (budget >= 0 && k >= budget) ? 0x7fffffff : atomicAdd(z.ticket, 1);
        s_tk[k % kZBufs] = c;
        if (c < nchunks) {
            const long long f1 = z.total_frames - (long long)c * rows;
            const long long f0 = f1 > rows ? f1 - rows : 0;
            tma_load_1d(buf0 + (size_t)(k % kZBufs) * buf_f4, z.feat + (size_t)f0 * kZBins,
                        (unsigned)(f1 - f0) * kZBins * 4u, &s_zbar[k % kZBufs], drop);
        }
    };
    __syncthreads();
    if (tid == 0)
        for (int k = 0; k < kZBufs - 1; ++k) issue(k);
    __syncthreads();
    const int q = tid % (kZBins / 4), r0 = tid / (kZBins / 4);      // column group, first row (threads >= 250 idle)
    for (int k = 0;; ++k) {
        // (the buffer of ticket k + kZBufs - 1 is the one iteration k - 1 released; s_tk[k] was published at least one
        // barrier ago)
        if (tid == 0) issue(k + kZBufs - 1);
        const int c = s_tk[k % kZBufs];
        if (c >= nchunks) break;             // (tickets only grow: nothing is in flight for the later ones either)
        const long long f1 = z.total_frames - (long long)c * rows;
        const long long f0 = f1 > rows ? f1 - rows : 0;
        // utterance of the chunk's first frame: largest b with fo[b] <= f0
        int b = 0;
        {
            int lo = 0, hi = z.batch - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (fo[mid] <= f0) lo = mid; else hi = mid - 1;
            }
            b = lo;
        }
        const float4* src = buf0 + (size_t)(k % kZBufs) * buf_f4;
        float4* dst = reinterpret_cast<float4*>(z.feat + (size_t)f0 * kZBins);
        const unsigned parity = (unsigned)(k / kZBufs) & 1u;
        bool arrived = false;
        long long cur = f0;
        while (cur < f1) {
            const long long next = fo[b + 1];
            const long long s1 = next < f1 ? next : f1;
            if (s1 > cur && r0 < 5) {
                const float* st = z.stats + (size_t)b * 3 * kZBins + 4 * q;
                const float4 mh = __ldg(reinterpret_cast<const float4*>(st));
                const float4 ml = __ldg(reinterpret_cast<const float4*>(st + kZBins));
                const float4 iv = __ldg(reinterpret_cast<const float4*>(st + 2 * kZBins));
                if (!arrived) { mbar_wait(&s_zbar[k % kZBufs], parity); arrived = true; }
#pragma unroll 4
                for (int r = (int)(cur - f0) + r0; r < (int)(s1 - f0); r += 5) {
                    const float4 v = src[r * (kZBins / 4) + q];
                    float4 o;
                    o.x = ((v.x - mh.x) - ml.x) * iv.x;
                    o.y = ((v.y - mh.y) - ml.y) * iv.y;
                    o.z = ((v.z - mh.z) - ml.z) * iv.z;
                    o.w = ((v.w - mh.w) - ml.w) * iv.w;
                    stg_evict_first(dst + r * (kZBins / 4) + q, o);
                }
            }
            if (s1 > cur) cur = s1;
            if (cur < f1) ++b;
        }
        if (!arrived) mbar_wait(&s_zbar[k % kZBufs], parity);               // (idle threads: the phase must be observed)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // reads of this buffer before its next bulk write
        __syncthreads();
    }
}

template <int NV4, bool PROB>
__device__ __forceinline__ void fused_small_body(const Params& p, float* sm, int b);

template <int NV4, bool PROB>
__global__ void __maxnreg__(112) fused_small_kernel(Params p) {
    extern __shared__ __align__(16) float sm[];
    if (p.z.feat == nullptr) {
        fused_small_body<NV4, PROB>(p, sm, (int)blockIdx.x);
        return;
    }
    // with co-work: utterances are CLAIMED (ticket z.ticket[1]), CTC work first -- a CTA that looped on z-score chunks
    // while utterances were still waiting for a slot (fewer than two CTAs resident per SM at launch: observed, the
    // kernel then took 130-230 us instead of 84) would put the z-score in front of the CTC critical path
    __shared__ int s_utt;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_utt = atomicAdd(p.z.ticket + 1, 1);
        __syncthreads();
        const int b = s_utt;
        if (b >= p.B) break;
        fused_small_body<NV4, PROB>(p, sm, b);
    }
    zscore_cowork(p.z, sm, -1);
}

template <int NV4, bool PROB>
__device__ __forceinline__ void fused_small_body(const Params& p, float* sm, const int b) {
    __shared__ double s_fin;
#ifdef ASRK_CTC_TIMING
    __shared__ unsigned long long s_tick[12];
#endif
    ASRK_TICK(0);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int T = p.input_len[b];
    const int V4 = p.V >> 2;
    // one row buffer per warp behind the lattice region, filled by TMA bulk copies on the warp's
    // mbarrier: the next row of the warp arrives while the current one is reduced, so HBM / L2
    // latency is off the chain.  The first rows are requested before anything else is known about
    // the utterance (the label preparation below runs under their latency).
    float4* rb = reinterpret_cast<float4*>(sm + kSmallSmemFloats) + (size_t)warp * V4;
    const float* xb = p.logits + (size_t)b * p.stride_b;
    const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
    __shared__ uint64_t s_bar[kRowWarps];
    uint64_t* bar = s_bar + warp;
    unsigned parity = 0;
    if (lane == 0) mbar_init(bar, 1);
    mbar_fence_init();
    __syncwarp();
    const bool prefetched = (T >= 1 && T <= p.T && warp < T);
    if (prefetched) issue_row(rb, xb + (size_t)warp * p.stride_t, V4, lane, keep, bar);
    int status, L;
    if (p.prep_fused) {
        __shared__ int s_prep[kRowWarps * 32 + 40];
        const PrepOut po = prep_body(p, b, s_prep);
        status = po.status;
        L = po.L;
    } else {
        status = p.row_status[b];
        L = p.eff_len[b];
    }
    const double ninf = (double)kNegInf;
    if (status == ASRK_ROW_BAD_LENGTH || status == ASRK_ROW_NOT_SMALL) {
        if (prefetched) mbar_wait(bar, 0);      // no bulk copy may be in flight when the CTA exits
        if (tid == 0) { p.loss[b] = __int_as_float(0x7fc00000); p.logp[b] = ninf; }
        if (p.tokens && tid == 0) { p.token_len[b] = 0; if (p.neg_sum_logits) p.neg_sum_logits[b] = 0.f; }
        if (p.grad) {   // keep the gradient defined: zeros
            for (int t = warp; t < p.T; t += kRowWarps) {
                float4* g4 = reinterpret_cast<float4*>(p.grad + (size_t)t * p.gstride_t + (size_t)b * p.gstride_b);
                for (int i = lane; i < (p.V >> 2); i += 32) stg_stream(g4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
        return;
    }
    if (!small_lattice(L, T)) {                 // generic path
        if (prefetched) mbar_wait(bar, 0);
        return;
    }
    const int W = L + 1;                        // row width of the gathered log-probs
    const int Ub = 2 * L + 1;                   // lattice states
    double* sC = reinterpret_cast<double*>(sm); // [T] level of alpha column t: everything subtracted up to t
    double* sD = sC + T;                        // [T] level of beta column t
    float* slse = reinterpret_cast<float*>(sD + T);  // [T]
    float* smax = slse + T;                     // [T]
    int* samax = reinterpret_cast<int*>(smax + T);   // [T]
    float* sK = smax + 2 * T;                   // [T] C_t + D_t - log2 p
    float* slp = sK + T;                        // [T][W] y_t(l'_j), linear
    float* sal = slp + T * W;                   // [T][Ub]
    float* sbe = sal + T * Ub;                  // [T][Ub]
    const int* eff = p.eff_labels + (size_t)b * p.Ls;

    // ---- A: row statistics + gather -----------------------------------------
    const int gc = (lane == 0) ? p.blank : ((lane < W) ? eff[lane - 1] : 0);   // class gathered by this lane
    for (int t = warp; t < T; t += kRowWarps) {
        mbar_wait(bar, parity);
        parity ^= 1;
        RowRegs<NV4> r;
        lds_row(rb, V4, lane, r);
        float xg = reinterpret_cast<const float*>(rb)[gc];
        asm volatile("" : "+f"(xg));
        release_row(r.v);
        if (t + kRowWarps < T) issue_row(rb, xb + (size_t)(t + kRowWarps) * p.stride_t, V4, lane, keep, bar);
        if (PROB) {
            // the op's input is log(p + eps): y = (p + eps) / sum(p + eps), no exponentials; slse holds 1 / sum
            const float ssum = row_sum_eps(r, V4, lane, p.eps);
            if (lane == 0) { slse[t] = __fdividef(1.0f, ssum); smax[t] = 0.f; samax[t] = 0; }
            if (lane < W) slp[t * W + lane] = (xg + p.eps) * __fdividef(1.0f, ssum);   // y_t(l'_j)
        } else {
            float m;
            int am;
            row_argmax(r, lane, m, am);
            const float ssum = row_sumexp(r, m);
            const float lse = m + __logf(ssum);
            if (lane == 0) { slse[t] = lse; smax[t] = m; samax[t] = am; }
            if (lane < W) slp[t * W + lane] = ex2_fast((xg - lse) * 1.4426950408889634f);   // y_t(l'_j)
        }
    }
    __syncthreads();
    ASRK_TICK(1);

    // ---- B: alpha || beta || greedy collapse ----------------------------------
    // LINEAR-domain recursions in fp64 (the fp64 pipe is idle in this kernel): a step is one shuffle, two or
    // three DADD and one DMUL on the dependency chain (~50 cycles) instead of a log-sum-exp with three MUFU
    // round trips (~280 cycles in round 1: 13 us of a CTA's 48 us with six of its eight warps waiting).
    // Every column is rescaled by an exact power of two, the top exponent of the PREVIOUS column (one REDUX on
    // the high words; it meets the chain only at the final multiply, through y 2^-e), so a stored column's
    // largest value stays in [y_min, 6): nothing drifts with T and no low-mass state is crushed the way a
    // float32 linear recursion crushed them.  (Taking the exponent two columns back would take the reduction
    // off the chain entirely, but that control loop is only marginally stable: the column maxima random-walk.)
    // The accumulated exponents are integers: the levels are exact.
    //   alpha_t(u) = A_t(u) 2^(C_t),  beta_t(u) (excludes y_t, TF) = B_t(u) 2^(D_t)
    // Stored for the gradient pass: the HIGH WORD of A and of B (sign, 11-bit exponent, 20 mantissa bits): no
    // conversion instruction in the recursion loop, the whole fp64 range, 2^-20 relative precision.  (A lattice that
    // fits this kernel -- T <= 200 -- spans ~100 binades per column; the generic kernel's T = 1998 columns span more
    // than fp64 has and stay in the log domain.)
    const int i = lane;
    const int lab_i = (i < L) ? eff[i] : -1;
    const int lab_im1 = (i >= 1 && i <= L) ? eff[i - 1] : -2;
    const bool skip = (i >= 1 && i < L && lab_i != lab_im1);
    const bool has_blank = (i <= L);
    auto pow2 = [](int e) { return __hiloint2double((1023 + e) << 20, 0); };
    auto top_exponent = [](double a, double b) {              // exponent of the largest value of the column (0 if all zero)
        const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)max(__double2hiint(a), __double2hiint(b)));
        return hi ? (int)(hi >> 20) - 1023 : 0;
    };
    if (warp == 0) {
        // alpha: pair (blank 2i, label 2i+1)
        const bool has_lab = (i < L);
        double a_b = 0.0, a_l = 0.0;
        if (i == 0) {
            a_b = (double)slp[0];
            if (L >= 1) a_l = (double)slp[1];
        }
        float yb_n = 0.f, yl_n = 0.f;            // y of the next frame, loaded ahead of the chain
        if (T > 1) { yb_n = slp[W]; yl_n = has_lab ? slp[W + 1 + i] : 0.f; }
        int e1 = 0;                              // top exponent of column t-1
        int lvl = 0;                             // C_t
        for (int t = 0; t < T; ++t) {
            if (t > 0) {
                const double sc = pow2(-e1);
                const double yb = (double)yb_n * sc, yl = (double)yl_n * sc;
                if (t + 1 < T) { yb_n = slp[(t + 1) * W]; yl_n = has_lab ? slp[(t + 1) * W + 1 + i] : 0.f; }
                double p1 = __shfl_up_sync(0xffffffffu, a_l, 1);
                if (i == 0) p1 = 0.0;
                const double nb = yb * (a_b + p1);
                const double nl = yl * ((a_l + a_b) + (skip ? p1 : 0.0));
                a_b = has_blank ? nb : 0.0;
                a_l = nl;                          // (yl = 0 where the label state does not exist)
                lvl += e1;
            }
            float* o = sal + t * Ub;
            if (has_blank) o[2 * i] = __int_as_float(__double2hiint(a_b));
            if (has_lab) o[2 * i + 1] = __int_as_float(__double2hiint(a_l));
            sC[t] = (double)lvl;                   // every lane, same value
            e1 = top_exponent(a_b, a_l);
        }
        // mass of the two terminal states at T-1 (relative to the last column's level)
        const double fb = __shfl_sync(0xffffffffu, a_b, L);
        const double fl = (L >= 1) ? __shfl_sync(0xffffffffu, a_l, L - 1) : 0.0;
        if (lane == 0) s_fin = log2(fb + fl);      // -inf when no alignment exists
#ifdef ASRK_CTC_TIMING
        if (lane == 0) s_tick[8] = (unsigned long long)clock64();
#endif
    } else if (warp == 1) {
        if (p.grad != nullptr) {
            // beta: pair (label 2i-1, blank 2i); excludes y_t
            const bool has_lab = (i >= 1 && i <= L);
            double b_l = 0.0, b_b = 0.0;
            if (i == L) {
                b_b = 1.0;
                if (L >= 1) b_l = 1.0;
            }
            float yb_n = 0.f, yl_n = 0.f;
            if (T > 1) { yb_n = slp[(T - 1) * W]; yl_n = has_lab ? slp[(T - 1) * W + i] : 0.f; }
            int e1 = 0;
            int lvl = 0;                         // D_t
            for (int t = T - 1; t >= 0; --t) {
                if (t < T - 1) {
                    const double sc = pow2(-e1);
                    const double yb = (double)yb_n * sc, yl = (double)yl_n * sc;
                    if (t >= 1) { yb_n = slp[t * W]; yl_n = has_lab ? slp[t * W + i] : 0.f; }
                    const double e_b = b_b * yb;
                    const double e_l = b_l * yl;
                    double n1 = __shfl_down_sync(0xffffffffu, e_l, 1);
                    if (i == 31) n1 = 0.0;
                    const double nbb = e_b + n1;
                    const double nbl = (e_l + e_b) + (skip ? n1 : 0.0);
                    b_b = has_blank ? nbb : 0.0;
                    b_l = has_lab ? nbl : 0.0;
                    lvl += e1;
                }
                float* o = sbe + t * Ub;
                if (has_blank) o[2 * i] = __int_as_float(__double2hiint(b_b));
                if (has_lab) o[2 * i - 1] = __int_as_float(__double2hiint(b_l));
                sD[t] = (double)lvl;               // every lane, same value: no divergent branch in the chain
                e1 = top_exponent(b_b, b_l);
            }
        }
#ifdef ASRK_CTC_TIMING
        if (lane == 0) s_tick[9] = (unsigned long long)clock64();
#endif
    } else if (warp == 2 && p.tokens != nullptr) {
        // greedy decode of this utterance: ballot + prefix count over 32-frame chunks
        int* out = p.tokens + (size_t)b * p.token_stride;
        int count = 0, carry_prev = -1;
        double nsl = 0.0;
        for (int base = 0; base < T; base += 32) {
            const int t = base + lane;
            const bool in = t < T;
            const int c = in ? samax[t] : -1;
            int prev = __shfl_up_sync(0xffffffffu, c, 1);
            if (lane == 0) prev = carry_prev;
            const bool keep = in && c != p.blank && !(p.merge_repeated && c == prev);
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            const int rank = __popc(mask & ((1u << lane) - 1u));
            if (keep && count + rank < p.token_stride) out[count + rank] = c;
            count += __popc(mask);
            carry_prev = __shfl_sync(0xffffffffu, c, 31);
            if (in) nsl += (double)smax[t];
        }
        nsl = warp_sum(nsl);
        if (lane == 0) {
            p.token_len[b] = count < p.token_stride ? count : p.token_stride;
            if (p.neg_sum_logits) p.neg_sum_logits[b] = (float)(-nsl);
        }
#ifdef ASRK_CTC_TIMING
        if (lane == 0) s_tick[10] = (unsigned long long)clock64();
#endif
    }
    __syncthreads();
    ASRK_TICK(2);
    ASRK_TICK(5);
    ASRK_TICK(7);
    // log2 p = C_{T-1} + log2(relative mass of the two terminal alphas at T-1)
    const double logp2 = sC[T - 1] + s_fin;
    const double logp = logp2 * 0.6931471805599453;
    if (tid == 0) {
        p.logp[b] = logp;
        p.loss[b] = (float)(-logp);
        if (logp == ninf && status == ASRK_ROW_OK) p.row_status[b] = ASRK_ROW_INFEASIBLE;
    }
    if (p.grad == nullptr) return;

    // ---- C: gradient rows -----------------------------------------------------
    const bool fix = (logp != ninf) && (status == ASRK_ROW_OK);   // TF: no valid path -> dy = y
    // occupancy(t, u) = alpha_t(u) beta_t(u) / p = 2^(ahat_t(u) + bhat_t(u) + K_t)
    ASRK_TICK(6);
    // occupancy(t, u) = alpha_t(u) beta_t(u) / p = A B 2^(C_t + D_t - log2 p), formed in double: the integer part of
    // the exponent is exact, the fraction of log2 p one factor per utterance.  sD[t] becomes that factor.
    if (fix) {
        const double lpf = floor(logp2);
        const double fr = exp2(-(logp2 - lpf));
        const int lpi = (int)lpf;
        for (int t = tid; t < T; t += kRowWarps * 32) {
            int K = (int)(sC[t] + sD[t]) - lpi;
            K = K < -1022 ? -1022 : (K > 1023 ? 1023 : K);
            sD[t] = __hiloint2double((1023 + K) << 20, 0) * fr;
        }
    }
    __syncthreads();
    ASRK_TICK(3);
    const float scale = p.grad_scale ? p.grad_scale[b] : 1.0f;
    // lane j < L owns label position j: the positions that carry the same label (one
    // MATCH), so repeated labels are summed in a fixed order without touching memory
    const int my_lab = (lane < L) ? eff[lane] : -1 - lane;
    const unsigned same = __match_any_sync(0xffffffffu, my_lab);
    const bool owner = (lane < L) && ((int)__ffs(same) - 1 == lane) && (my_lab != p.blank);
    const bool is_blank_lab = (lane < L) && (my_lab == p.blank);
    if (!PROB && warp < T) issue_row(rb, xb + (size_t)warp * p.stride_t, V4, lane, drop, bar);   // L2 hits: read a moment ago
    for (int t = warp; t < p.T; t += kRowWarps) {
        float* g = p.grad + (size_t)t * p.gstride_t + (size_t)b * p.gstride_b;
        float4* g4 = reinterpret_cast<float4*>(g);
        if (t >= T) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k = lane; k < V4; k += 32) stg_evict_first(g4 + k, z);
            continue;
        }
        if (PROB) {
            // dL/dp = dL/dx / (p + eps) with dL/dx = y - occupancy and y = (p + eps) / S:  scale / S for every
            // class (the row is not read again), minus scale occupancy / (p + eps) = (scale / S) occupancy / y
            // for the classes of the lattice
            const float base = scale * slse[t];
            const float4 bv = make_float4(base, base, base, base);
            for (int k = lane; k < V4; k += 32) stg_evict_first(g4 + k, bv);
            if (!fix) continue;
            __syncwarp();
            const float* al = sal + t * Ub;
            const float* be = sbe + t * Ub;
            const double Kt = sD[t];
            float ob = (lane <= L) ? (float)(__hiloint2double(__float_as_int(al[2 * lane]), 0) * __hiloint2double(__float_as_int(be[2 * lane]), 0) * Kt) : 0.f;
            if (is_blank_lab) ob += (float)(__hiloint2double(__float_as_int(al[2 * lane + 1]), 0) * __hiloint2double(__float_as_int(be[2 * lane + 1]), 0) * Kt);
            if (owner) {
                float o = 0.f;
                for (unsigned mset = same; mset; mset &= mset - 1) {
                    const int k = __ffs(mset) - 1;
                    o += (float)(__hiloint2double(__float_as_int(al[2 * k + 1]), 0) * __hiloint2double(__float_as_int(be[2 * k + 1]), 0) * Kt);
                }
                g[my_lab] = base * (1.0f - __fdividef(o, slp[t * W + 1 + lane]));
            }
            ob = warp_sum(ob);
            if (lane == 0) g[p.blank] = base * (1.0f - __fdividef(ob, slp[t * W]));
            continue;
        }
        const float nlse2 = -slse[t] * kLog2e;
        mbar_wait(bar, parity);
        parity ^= 1;
        float4 v[NV4];
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
            const int idx = lane + 32 * k;
            v[k] = (idx < V4) ? rb[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        release_row(v);
        if (t + kRowWarps < T) issue_row(rb, xb + (size_t)(t + kRowWarps) * p.stride_t, V4, lane, drop, bar);
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
            const int idx = lane + 32 * k;
            if (idx < V4) {
                float4 y;
                y.x = ex2_fast(fmaf(v[k].x, kLog2e, nlse2)) * scale;
                y.y = ex2_fast(fmaf(v[k].y, kLog2e, nlse2)) * scale;
                y.z = ex2_fast(fmaf(v[k].z, kLog2e, nlse2)) * scale;
                y.w = ex2_fast(fmaf(v[k].w, kLog2e, nlse2)) * scale;
                stg_evict_first(g4 + idx, y);
            }
        }
        if (!fix) continue;
        __syncwarp();
        const float* al = sal + t * Ub;
        const float* be = sbe + t * Ub;
        const double Kt = sD[t];
        float ob = (lane <= L) ? (float)(__hiloint2double(__float_as_int(al[2 * lane]), 0) * __hiloint2double(__float_as_int(be[2 * lane]), 0) * Kt) : 0.f;          // blank states (L <= 31)
        if (is_blank_lab) ob += (float)(__hiloint2double(__float_as_int(al[2 * lane + 1]), 0) * __hiloint2double(__float_as_int(be[2 * lane + 1]), 0) * Kt);           // a label equal to the blank index
        if (owner) {
            float o = 0.f;
            for (unsigned mset = same; mset; mset &= mset - 1) {
                const int k = __ffs(mset) - 1;
                o += (float)(__hiloint2double(__float_as_int(al[2 * k + 1]), 0) * __hiloint2double(__float_as_int(be[2 * k + 1]), 0) * Kt);
            }
            g[my_lab] = (slp[t * W + 1 + lane] - o) * scale;
        }
        ob = warp_sum(ob);
        if (lane == 0) g[p.blank] = (slp[t * W] - ob) * scale;
    }
#ifdef ASRK_CTC_TIMING
    __syncthreads();
    ASRK_TICK(4);
    if (tid == 0 && p.tokens) {       // debug build only: phase times (ns) over the first token slots
        int* o = p.tokens + (size_t)b * p.token_stride;
        unsigned sm_id;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_id));
        o[0] = (int)(s_tick[0] & 0x3fffffffull);
        for (int i = 1; i <= 4; ++i) o[i] = (int)(s_tick[i] - s_tick[0]);
        o[5] = (int)sm_id; o[6] = T; o[7] = L; o[8] = (int)(s_tick[5] - s_tick[2]); o[9] = (int)(s_tick[6] - s_tick[5]); o[10] = (int)(s_tick[7] - s_tick[5]); o[11] = (int)(s_tick[8] - s_tick[1]); o[12] = (int)(s_tick[9] - s_tick[1]); o[13] = (int)(s_tick[10] - s_tick[1]);
    }
#endif
}

// One CTA per utterance; thread i owns the state pair
//   alpha sweep: (blank 2i, label 2i+1)     beta sweep: (label 2i-1, blank 2i)
// so that each step needs exactly one neighbour value (the previous / next label state): a warp shuffle,
// and across warps one float per warp through a double-buffered shared array behind the step's only
// __syncthreads.  Log2 domain in fp32, every column stored relative to a LEVEL kept in double (the block-wide
// maximum of the previous column is subtracted each step; it rides on the same barrier: one word per warp,
// reduced by every warp with one REDUX).  A linear-domain fp64 recursion with one scale per column -- what the
// fused kernel uses for its short lattices -- is 4x shorter per step but cannot hold these columns: at T = 1998,
// L = 300 the states that carry the occupancy sit 2^-1020 below the column's largest value (measured on the
// C3 batch: tools/debug_c3.py), beyond fp64's normal range.
// The probabilities of the next frames and, in the beta sweep, the stored alpha values and levels are fetched
// kRing - 1 steps ahead with cp.async into a shared-memory ring, so no global-memory round trip sits inside a
// step (a register look-ahead does not work: six scoreboards per warp, waiting for the oldest load waits for the
// newest too).  Round 1: 1.1 us per step, all of it L2 latency; T = 1998 frames x 2 sweeps = 4.5 ms per C3 batch.
constexpr int kRing = 8;
__host__ __device__ inline int lattice_slot_floats(int P) { return 4 + P; }   // pad(2) yb(1) pad(1) | yl[P]
__device__ __forceinline__ unsigned float_key(float v) {      // order-preserving bit pattern
    const unsigned b = __float_as_uint(v);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned m) { return __uint_as_float(m ^ ((m >> 31) ? 0x80000000u : 0xffffffffu)); }
// grid (B, 2): CTA (b, 0) runs the alpha sweep of utterance b and CTA (b, 1) -- launched only when a gradient is
// wanted -- its beta sweep AT THE SAME TIME on another SM (the two recursions are independent; round 2's first
// version ran them back to back in one CTA, the beta sweep forming the occupancies from the stored alpha: 2 T
// dependent steps instead of T).  Both store their columns as float32 relative to their own per-column level
// (double); grad_kernel forms occupancy(t, u) = 2^(ahat + bhat + C_t + D_t - log2 p) when it stages a frame.
__global__ void lattice_kernel(Params p) {
    extern __shared__ double smd[];
    const int b = blockIdx.x;
    const bool beta = (blockIdx.y == 1);
    const int i = threadIdx.x;
    const int P = blockDim.x;
    const int lane = i & 31, warp = i >> 5, nwarp = P >> 5;
    double* edged = smd;                                      // [2][32] last / first running value of every warp
    unsigned* wtop = reinterpret_cast<unsigned*>(smd + 64);   // [2][32] key of every warp's column maximum
    double* fin2d = smd + 96;                                 // [2]
    float* ring = reinterpret_cast<float*>(smd + 98);         // [kRing][slot]: look-ahead frames, filled by cp.async
    const int SF = lattice_slot_floats(P);
    const int status = p.row_status[b];
    const int L = p.eff_len[b];
    const int T = p.input_len[b];
    const double ninf = (double)kNegInf;
    if (status == ASRK_ROW_BAD_LENGTH) {
        if (i == 0 && !beta) { p.loss[b] = __int_as_float(0x7fc00000); p.logp[b] = ninf; }
        return;
    }
    if (p.fused && small_lattice(L, T)) return;   // handled by fused_small_kernel
    const int S = p.Ls + 1;
    const int U = 2 * p.Ls + 1;
    const float* lpl = p.lpl + (size_t)b * p.T * S;
    const int* eff = p.eff_labels + (size_t)b * p.Ls;

    const int lab_i = (i < L) ? eff[i] : -1;                     // label index i   (state 2i+1)
    const int lab_im1 = (i >= 1 && i <= L) ? eff[i - 1] : -2;    // label index i-1 (state 2i-1)
    // the skip transition 2i-1 <-> 2i+1 exists when both labels exist and differ
    const bool skip = (i >= 1 && i < L && lab_i != lab_im1);
    const bool has_blank = (i <= L);
    const bool has_lab_a = (i < L);             // alpha pair's label state 2i+1 exists
    const bool has_lab_b = (i >= 1 && i <= L);  // beta pair's label state 2i-1 exists
    // maximum of the whole column from the per-warp keys written before the step's barrier (0 if all -inf)
    auto block_max = [&](int buf) {
        const unsigned w = (lane < nwarp) ? wtop[buf * 32 + lane] : 0u;
        const float m = key_float(__reduce_max_sync(0xffffffffu, w));
        return (m == kNegInf) ? 0.f : m;
    };
    auto slot = [&](int frame) { return ring + (size_t)((frame % kRing + kRing) % kRing) * SF; };

    if (!beta) {
        // ------------------------------ alpha ------------------------------
        float* occ = p.occ + (size_t)b * p.T * U;
        double* coff = p.coff + (size_t)b * p.T;
        // look-ahead: log2 y of frame f (blank, this thread's label) -> slot(f)
        auto fetch_a = [&](int f) {
            float* s = slot(f);
            const bool in = (f >= 1 && f < T);
            if (i == 0) cp_async4(s + 2, lpl + (size_t)(in ? f : 0) * S, in ? 4 : 0);
            cp_async4(s + 4 + i, lpl + (size_t)(in ? f : 0) * S + 1 + (has_lab_a ? i : 0), (in && has_lab_a) ? 4 : 0);
            cp_async_commit();
        };
        // running values: absolute log2 alpha in double (exchanged as such); STORED as float32 relative to the level C
        // (double, the same in every thread: it follows the block-wide maximum of the previous column)
        double a_b = ninf, a_l = ninf;
        if (i == 0) {
            a_b = (double)lpl[0];
            if (L >= 1) a_l = (double)lpl[1];
        }
        double lvl = 0.0;                             // C_t
        {
            const float rb = (float)a_b, rl = (float)a_l;
            if (has_blank) occ[2 * i] = rb;
            if (has_lab_a) occ[2 * i + 1] = rl;
            if (i == 0) coff[0] = 0.0;
            const float c = warp_max_redux(fmaxf(rb, rl));
            if (lane == 0) wtop[warp] = float_key(c);
            if (lane == 31) edged[warp] = a_l;
        }
        for (int k = 0; k < kRing - 1; ++k) fetch_a(1 + k);
        for (int t = 1; t < T; ++t) {
            const int buf = (t - 1) & 1;
            cp_async_wait<kRing - 2>();                // this thread's copies of frame t have landed
            __syncthreads();                           // ... everybody's; column t-1 is published (edges, maxima)
            fetch_a(t + kRing - 1);                    // (into the slot frame t-1 used: everyone has read it)
            const float* sl = slot(t);
            lvl += (double)block_max(buf);
            const double yb = (double)sl[2], yl = has_lab_a ? (double)sl[4 + i] : ninf;
            double p1 = __shfl_up_sync(0xffffffffu, a_l, 1);
            if (lane == 0) p1 = (warp > 0) ? edged[buf * 32 + warp - 1] : ninf;
            const double nb = yb + lse2_log2d(a_b, p1);
            const double nl = yl + lse3_log2d(a_l, a_b, skip ? p1 : ninf);
            a_b = has_blank ? nb : ninf;
            a_l = nl;
            const float rb = (float)(a_b - lvl), rl = (float)(a_l - lvl);
            float* o = occ + (size_t)t * U;
            if (has_blank) o[2 * i] = rb;
            if (has_lab_a) o[2 * i + 1] = rl;
            if (i == 0) coff[t] = lvl;
            const float c = warp_max_redux(fmaxf(rb, rl));
            if (lane == 0) wtop[(buf ^ 1) * 32 + warp] = float_key(c);
            if (lane == 31) edged[(buf ^ 1) * 32 + warp] = a_l;
        }
        cp_async_wait<0>();
        // log2 p = log2( alpha_{T-1}(2L) + alpha_{T-1}(2L-1) )
        __syncthreads();
        if (i == L) fin2d[0] = a_b;
        if (i == L - 1) fin2d[1] = a_l;
        if (L == 0 && i == 0) fin2d[1] = ninf;
        __syncthreads();
        const double logp2 = lse2_log2d(fin2d[0], fin2d[1]);
        const double logp = logp2 * 0.6931471805599453;
        if (i == 0) {
            p.logp[b] = logp;
            p.loss[b] = (float)(-logp);
            if (logp == ninf && status == ASRK_ROW_OK) p.row_status[b] = ASRK_ROW_INFEASIBLE;
        }
        return;
    }

    // ------------------------------ beta -------------------------------
    // beta excludes y_t (TensorFlow's convention); e(u) = beta_{t+1}(u) + log2 y_{t+1}(l'_u).
    // look-ahead for the step that produces column f: y of frame f + 1 -> slot(f)
    float* bet = p.beta + (size_t)b * p.T * U;
    double* coffb = p.coffb + (size_t)b * p.T;
    auto fetch_b = [&](int f) {
        float* s = slot(f);
        const bool iny = (f >= 0 && f + 1 < T);
        if (i == 0) cp_async4(s + 2, lpl + (size_t)(iny ? f + 1 : 0) * S, iny ? 4 : 0);
        cp_async4(s + 4 + i, lpl + (size_t)(iny ? f + 1 : 0) * S + (has_lab_b ? i : 0), (iny && has_lab_b) ? 4 : 0);   // label i-1 -> slot i
        cp_async_commit();
    };
    double b_l = ninf, b_b = ninf;
    if (i == L) {
        b_b = 0.0;
        if (L >= 1) b_l = 0.0;
    }
    double lvl = 0.0;                                 // D_t
    for (int k = 0; k < kRing - 1; ++k) fetch_b(T - 1 - k);
    for (int t = T - 1; t >= 0; --t) {
        const int buf = t & 1;
        cp_async_wait<kRing - 3>();                // frames t and t-1 have landed (t-1: the edge value below)
        __syncthreads();                           // ... everybody's; column t+1 is published
        fetch_b(t - (kRing - 1));                  // (into the slot frame t+1 used: everyone has read it)
        const float* sl = slot(t);
        if (t < T - 1) {
            lvl += (double)block_max(buf ^ 1);
            const double yb = (double)sl[2], yl = has_lab_b ? (double)sl[4 + i] : ninf;
            const double e_b = b_b + yb;
            const double e_l = b_l + yl;
            // the pair needs e of the NEXT label state (2i+1): lane i+1's e_l; across warps the first lane
            // of the next warp published b_l + log2 y_l (absolute, double) with its column
            double n1 = __shfl_down_sync(0xffffffffu, e_l, 1);
            if (lane == 31) n1 = (warp + 1 < nwarp) ? edged[(buf ^ 1) * 32 + warp + 1] : ninf;
            const double nbb = lse2_log2d(e_b, n1);
            const double nbl = lse3_log2d(e_l, e_b, skip ? n1 : ninf);
            b_b = has_blank ? nbb : ninf;
            b_l = has_lab_b ? nbl : ninf;
        }
        const float rb = (float)(b_b - lvl), rl = (float)(b_l - lvl);
        float* o = bet + (size_t)t * U;
        if (has_blank) o[2 * i] = rb;
        if (has_lab_b) o[2 * i - 1] = rl;
        if (i == 0) coffb[t] = lvl;
        const float c = warp_max_redux(fmaxf(rb, rl));
        if (lane == 0) wtop[buf * 32 + warp] = float_key(c);
        // publish column t for the previous warp's last lane: b_l(first lane) + log2 y_t(its label), what it
        // needs as "e of the next label state" in the next step
        if (lane == 0) edged[buf * 32 + warp] = has_lab_b ? b_l + (double)slot(t - 1)[4 + i] : ninf;   // y_t of label i-1 (own copy)
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// grad
// ---------------------------------------------------------------------------
// One CTA = one utterance x kGradFrames consecutive frames (a warp per frame, grid-stride inside the block of frames).
// The utterance's label list and repeat chains are staged in shared memory once per CTA and every frame's occupancy
// row once per warp, with independent coalesced loads: the per-row fix-up of the lattice classes then walks shared
// memory.  (Round 1 mapped warps to rows in (t, b) order and walked the label list, the chains and the occupancies
// in global memory: ~30 dependent L2 round trips per row, 1.0 ms per C3 batch at 26 % of the HBM bandwidth.)
constexpr int kGradFrames = 64;
// Every gradient row is written ONCE, 16 bytes per lane: the occupancies of the lattice classes are scattered into a
// per-warp shared-memory row of V floats (zero elsewhere) BEFORE the row is formed, and cleared again afterwards.
// (The first version streamed y * scale out and then patched the lattice classes with 4-byte stores: with ~300
// labels per utterance nearly every 32-byte sector of the row was written twice, and `ncu` counted 1.06 GB of DRAM
// writes and 1.27 GB of reads for a 0.73 GB gradient.)
#ifndef ASRK_GRAD_MIN_CTAS
#define ASRK_GRAD_MIN_CTAS 3      // 80 registers: three CTAs per SM (the whole row in registers took 117: two)
#endif
template <int NV4>
__global__ void __launch_bounds__(kRowWarps * 32, ASRK_GRAD_MIN_CTAS) grad_kernel(Params p) {
    if (p.fused && *p.need_generic == 0) return;
    extern __shared__ __align__(16) int gsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kGradFrames;
    const int tl = p.input_len[b];
    const int status = p.row_status[b];
    const int V = p.V;
    const int L = p.eff_len[b];
    if (p.fused && (status == ASRK_ROW_BAD_LENGTH || small_lattice(L, tl))) return;
    const int Ls = p.Ls;
    const int U = 2 * Ls + 1;
    const int Vfix = (NV4 > 0) ? V : 0;                     // (V is a multiple of 4 on the vector path)
    float* s_fix = reinterpret_cast<float*>(gsm) + (size_t)warp * Vfix;   // [V] occupancy (or occupancy / y) per class
    int* s_eff = gsm + kRowWarps * Vfix;   // [Ls]
    int* s_nxt = s_eff + Ls;               // [Ls]
    int* s_fst = s_eff + 2 * Ls;           // [Ls]
    // per warp: [Ls] occupancies of the label states | [Ls + 1] log2 y of the lattice classes (staged only where the
    // row pass does not produce y itself: probabilities in, or the scalar path)
    const bool need_lpl = (NV4 == 0) || p.prob;
    const int per_warp = Ls + (need_lpl ? Ls + 1 : 0);
    float* s_lab = reinterpret_cast<float*>(s_eff + 3 * Ls) + (size_t)warp * per_warp;
    float* s_lpl = s_lab + Ls;
    const bool live = (status == ASRK_ROW_OK) && (t0 < tl);
    if (live) {
        for (int j = threadIdx.x; j < L; j += blockDim.x) {
            s_eff[j] = p.eff_labels[(size_t)b * Ls + j];
            s_nxt[j] = p.chain_next[(size_t)b * Ls + j];
            s_fst[j] = p.chain_first[(size_t)b * Ls + j];
        }
    }
    for (int k = threadIdx.x; k < kRowWarps * Vfix; k += blockDim.x) reinterpret_cast<float*>(gsm)[k] = 0.f;
    __syncthreads();
    const float scale = p.grad_scale ? p.grad_scale[b] : 1.0f;
    for (int t = t0 + warp; t < t0 + kGradFrames && t < p.T; t += kRowWarps) {
        float* g = p.grad + (size_t)t * p.gstride_t + (size_t)b * p.gstride_b;
        if (t >= tl || status == ASRK_ROW_BAD_LENGTH) {
            if constexpr (NV4 > 0) {
                float4* g4 = reinterpret_cast<float4*>(g);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = lane; i < (V >> 2); i += 32) stg_stream(g4 + i, z);
            } else {
                for (int i = lane; i < V; i += 32) g[i] = 0.f;
            }
            continue;
        }
        const float* x = p.logits + (size_t)t * p.stride_t + (size_t)b * p.stride_b;
        const size_t bt = (size_t)b * p.T + t;
        // the row itself first (the longest latency), then the frame's lattice columns
        [[maybe_unused]] float4 v[(NV4 > 0 ? NV4 : 1)];
        if constexpr (NV4 > 0) {
            if (!p.prob) {
                const float4* x4 = reinterpret_cast<const float4*>(x);
#pragma unroll
                for (int k = 0; k < NV4; ++k) {
                    const int i = lane + 32 * k;
                    if (i < (V >> 2)) v[k] = ldg_stream(x4 + i);
                }
            }
        }
        const float lse = p.lse[bt];
        const bool fix = (status == ASRK_ROW_OK);   // TF: no valid path -> dy = y
        float ob = 0.f;                             // blank occupancy: even states (+ labels equal to the blank index)
        if (fix) {
            // occupancy(t, u) = alpha_t(u) beta_t(u) / p = 2^(ahat + bhat + (C_t + D_t - log2 p)): the two sweeps stored
            // their columns relative to their own levels
            const float* al = p.occ + bt * (size_t)U;
            const float* be = p.beta + bt * (size_t)U;
            const double kt = (p.coff[bt] + p.coffb[bt]) - p.logp[b] * 1.4426950408889634;
            for (int k = lane; k < 2 * L + 1; k += 32) {
                // (the sum in double: the three terms are ~1000 binades each and cancel to the occupancy's exponent)
                const float o = ex2_fast((float)(((double)__ldcs(al + k) + (double)__ldcs(be + k)) + kt));
                if (k & 1) s_lab[k >> 1] = o;      // label state 2j+1
                else ob += o;                      // blank state
            }
            if (need_lpl) {
                const float* lpl = p.lpl + bt * (size_t)(Ls + 1);
                for (int k = lane; k <= L; k += 32) s_lpl[k] = __ldcs(lpl + k);
            }
        }
        const float base = p.prob ? scale * __expf(-lse) : 0.f;   // gradient w.r.t. the probabilities: scale / S
        if constexpr (NV4 == 0) {
            // scalar path (V not a multiple of 4, unaligned rows or a very large vocabulary): row, then patches
            if (p.prob) for (int i = lane; i < V; i += 32) g[i] = base;
            else for (int i = lane; i < V; i += 32) g[i] = __expf(x[i] - lse) * scale;
        }
        __syncwarp();
        // the lattice occupancies per class: blank = sum over even states (+ labels equal to the blank index); every
        // distinct label = sum over its chain of positions, in a fixed order
        if (fix) {
            for (int j = lane; j < L; j += 32) {
                const int c = s_eff[j];
                if (c == p.blank) {
                    ob += s_lab[j];
                } else if (s_fst[j]) {
                    float o = 0.f;
                    for (int k = j; k >= 0; k = s_nxt[k]) o += s_lab[k];
                    if constexpr (NV4 > 0) {
                        s_fix[c] = p.prob ? __fdividef(o, ex2_fast(s_lpl[1 + j])) : o;
                    } else {
                        const float y = ex2_fast(s_lpl[1 + j]);
                        g[c] = p.prob ? base * (1.0f - o / y) : (y - o) * scale;
                    }
                }
            }
            ob = warp_sum(ob);
            if (lane == 0) {
                if constexpr (NV4 > 0) {
                    s_fix[p.blank] = p.prob ? __fdividef(ob, ex2_fast(s_lpl[0])) : ob;
                } else {
                    const float y = ex2_fast(s_lpl[0]);
                    g[p.blank] = p.prob ? base * (1.0f - ob / y) : (y - ob) * scale;
                }
            }
        }
        if constexpr (NV4 > 0) {
            __syncwarp();
            const float nlse2 = -lse * kLog2e;
            float4* g4 = reinterpret_cast<float4*>(g);
            const float4* f4 = reinterpret_cast<const float4*>(s_fix);
#pragma unroll
            for (int k = 0; k < NV4; ++k) {
                const int i = lane + 32 * k;
                if (i < (V >> 2)) {
                    const float4 f = f4[i];
                    float4 y;
                    if (p.prob) {
                        y.x = base * (1.0f - f.x);
                        y.y = base * (1.0f - f.y);
                        y.z = base * (1.0f - f.z);
                        y.w = base * (1.0f - f.w);
                    } else {
                        y.x = (ex2_fast(fmaf(v[k].x, kLog2e, nlse2)) - f.x) * scale;
                        y.y = (ex2_fast(fmaf(v[k].y, kLog2e, nlse2)) - f.y) * scale;
                        y.z = (ex2_fast(fmaf(v[k].z, kLog2e, nlse2)) - f.z) * scale;
                        y.w = (ex2_fast(fmaf(v[k].w, kLog2e, nlse2)) - f.w) * scale;
                    }
                    stg_stream(g4 + i, y);
                }
            }
            __syncwarp();
            if (fix) {                           // clear the classes this row touched
                for (int j = lane; j < L; j += 32) {
                    const int c = s_eff[j];
                    if (c != p.blank && s_fst[j]) s_fix[c] = 0.f;
                }
                if (lane == 0) s_fix[p.blank] = 0.f;
            }
        }
        __syncwarp();                        // the warp's staging rows are free again
    }
}

// ---------------------------------------------------------------------------
// greedy decode: collapse repeats / drop blanks with a warp ballot + prefix count
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) collapse_kernel(Params p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= p.B) return;
    int tl = p.input_len[b];
    if (tl < 0) tl = 0;
    if (tl > p.T) tl = p.T;
    if (p.fused && (p.row_status[b] == ASRK_ROW_BAD_LENGTH || small_lattice(p.eff_len[b], tl))) return;
    const int* am = p.argmax + (size_t)b * p.T;
    const float* mx = p.rowmax + (size_t)b * p.T;
    int* out = p.tokens + (size_t)b * p.token_stride;
    int count = 0;
    int carry_prev = -1;           // class of the last frame of the previous chunk
    double nsl = 0.0;
    for (int base = 0; base < tl; base += 32) {
        const int t = base + lane;
        const bool in = t < tl;
        const int c = in ? am[t] : -1;
        int prev = __shfl_up_sync(0xffffffffu, c, 1);
        if (lane == 0) prev = carry_prev;
        const bool keep = in && c != p.blank && !(p.merge_repeated && c == prev);
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        if (keep && count + rank < p.token_stride) out[count + rank] = c;
        count += __popc(mask);
        carry_prev = __shfl_sync(0xffffffffu, c, 31);
        if (in) nsl += (double)mx[t];
    }
    nsl = warp_sum(nsl);
    if (lane == 0) {
        p.token_len[b] = count < p.token_stride ? count : p.token_stride;
        if (p.neg_sum_logits) p.neg_sum_logits[b] = (float)(-nsl);
    }
}

__global__ void fill_bad_rows_kernel(float* loss, int* row_status, int* token_len, float* nsl, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    loss[b] = __int_as_float(0x7fc00000);
    row_status[b] = ASRK_ROW_BAD_LENGTH;
    if (token_len) token_len[b] = 0;
    if (nsl) nsl[b] = 0.f;
}

// [sum of the losses of the rows TF would accept, number of such rows] in float64, one
// CTA, fixed order (reproducible): the operand of the path's only collective.
__global__ void __launch_bounds__(256) loss_sum_kernel(const float* loss, const int* row_status, int B, double* out2,
                                                       int accumulate) {
    __shared__ double s_sum[256];
    __shared__ int s_cnt[256];
    double a = 0.0;
    int c = 0;
    for (int b = threadIdx.x; b < B; b += 256) {
        if (row_status == nullptr || row_status[b] == ASRK_ROW_OK) { a += (double)loss[b]; ++c; }
    }
    s_sum[threadIdx.x] = a;
    s_cnt[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (accumulate) { out2[0] += s_sum[0]; out2[1] += (double)s_cnt[0]; }
        else { out2[0] = s_sum[0]; out2[1] = (double)s_cnt[0]; }
    }
}

// CTAs per SM of the two zero-copy staging kernels.  PCIe needs ~100 KB in flight per direction; one CTA per SM (8
// warps x 3 KB) holds 3.5 MB.  With 8 and 4 CTAs per SM (round 2's first version) the two kernels filled every thread
// slot of the chip and ran one after the other instead of side by side: 59 GB/s for both directions together against
// 50 + 50 for two DMA copies (tools/pcie_ceiling.py).
#ifndef ASRK_STAGE_CTAS_PER_SM
#define ASRK_STAGE_CTAS_PER_SM 1
#endif

// Device -> host return of the gradient without its padding: the mirror of stage_logits_kernel.  Rows
// t < input_len[b] are written straight into (mapped, pinned) host memory by the SMs with 16-byte stores;
// rows t >= input_len[b] (all zeros by definition) never cross PCIe.
__global__ void __launch_bounds__(256) unstage_rows_kernel(const float* src, long long sst, long long ssb, float* dst,
                                                           long long dst_t, long long dst_b, const int* input_len,
                                                           int* stale_len, int T, int B, int V) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)T * B;
    const int V4 = V >> 2;
    for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
        const int t = (int)(row / B), b = (int)(row % B);
        float4* d4 = reinterpret_cast<float4*>(dst + (size_t)t * dst_t + (size_t)b * dst_b);
        if (t >= input_len[b]) {
            // a row the buffer's previous occupant wrote and this batch does not: back to zero
            if (stale_len != nullptr && t < stale_len[b])
                for (int i = lane; i < V4; i += 32) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)t * sst + (size_t)b * ssb);
        for (int i0 = lane; i0 < V4; i0 += 32 * 6) {
            float4 v[6];
#pragma unroll
            for (int e = 0; e < 6; ++e)
                if (i0 + 32 * e < V4) v[e] = ldg_stream(s4 + i0 + 32 * e);
#pragma unroll
            for (int e = 0; e < 6; ++e)
                if (i0 + 32 * e < V4) d4[i0 + 32 * e] = v[e];
        }
    }
}

// stale_len <- input_len once every row of the copy above has been decided (a separate launch: stream order)
__global__ void copy_lengths_kernel(const int* input_len, int* stale_len, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) stale_len[b] = input_len[b];
}

// Host -> device staging of the logits without their padding: rows t < input_len[b] only are pulled
// straight out of (mapped, pinned) host memory by the SMs with 16-byte loads and written to the device
// tensor; rows t >= input_len[b] are never read (the CTC kernels never look at them).  One warp per
// row, grid-stride; every warp keeps several 16-byte requests per lane in flight to cover the PCIe
// round trip.
__global__ void __launch_bounds__(256) stage_logits_kernel(const float* src, long long sst, long long ssb,
                                                           float* dst, long long dst_t, long long dst_b,
                                                           const int* input_len, int T, int B, int V) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)T * B;
    const int V4 = V >> 2;
    for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
        const int t = (int)(row / B), b = (int)(row % B);
        if (t >= input_len[b]) continue;
        const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)t * sst + (size_t)b * ssb);
        float4* d4 = reinterpret_cast<float4*>(dst + (size_t)t * dst_t + (size_t)b * dst_b);
        for (int i0 = lane; i0 < V4; i0 += 32 * 6) {
            float4 v[6];
#pragma unroll
            for (int e = 0; e < 6; ++e)
                if (i0 + 32 * e < V4) v[e] = ldg_stream(s4 + i0 + 32 * e);
#pragma unroll
            for (int e = 0; e < 6; ++e)
                if (i0 + 32 * e < V4) d4[i0 + 32 * e] = v[e];
        }
    }
}

static int pick_nv4(const Params& p, const void* a, long long st, long long sb, const void* g,
                    long long gt, long long gb) {
    // vector path: V % 4 == 0, 16-byte aligned bases and strides, V <= 2048
    if (p.V % 4 != 0 || p.V > 2048) return 0;
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (st % 4) || (sb % 4)) return 0;
    if (g && ((reinterpret_cast<uintptr_t>(g) & 15) || (gt % 4) || (gb % 4))) return 0;
    const int v4 = p.V / 4;
    if (v4 <= 32 * 4) return 4;
    if (v4 <= 32 * 8) return 8;
    if (v4 <= 32 * 12) return 12;
    return 16;
}

template <bool WANT_LSE>
static void launch_rows(const Params& p, int nv4, cudaStream_t stream) {
    const long long rows = (long long)p.T * p.B;
    long long want = (rows + kRowWarps - 1) / kRowWarps;
    const long long cap = (long long)sm_count() * 8;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    switch (nv4) {
        case 4: rows_kernel<4, WANT_LSE><<<grid, kRowWarps * 32, 0, stream>>>(p), asrk::note_launch(); break;
        case 8: rows_kernel<8, WANT_LSE><<<grid, kRowWarps * 32, 0, stream>>>(p), asrk::note_launch(); break;
        case 12: rows_kernel<12, WANT_LSE><<<grid, kRowWarps * 32, 0, stream>>>(p), asrk::note_launch(); break;
        case 16: rows_kernel<16, WANT_LSE><<<grid, kRowWarps * 32, 0, stream>>>(p), asrk::note_launch(); break;
        default: rows_kernel<0, WANT_LSE><<<grid, kRowWarps * 32, 0, stream>>>(p), asrk::note_launch(); break;
    }
}

static void launch_fused(const Params& p, int nv4, cudaStream_t stream) {
    const size_t smem = sizeof(float) * (kSmallSmemFloats + (size_t)kRowWarps * p.V);
    // with z-score co-work: CTAs beyond the batch fill the resident slots the batch leaves free (two CTAs per SM)
    int grid = p.B;
    if (p.z.feat != nullptr && grid < 2 * sm_count()) grid = 2 * sm_count();
    switch (nv4) {
#define ASRK_FUSED_CASE(N)                                                                                        \
    case N:                                                                                                       \
        if (p.prob) {                                                                                             \
            cudaFuncSetAttribute(fused_small_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
            fused_small_kernel<N, true><<<grid, kRowWarps * 32, smem, stream>>>(p), asrk::note_launch();                                \
        } else {                                                                                                  \
            cudaFuncSetAttribute(fused_small_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            fused_small_kernel<N, false><<<grid, kRowWarps * 32, smem, stream>>>(p), asrk::note_launch();                               \
        }                                                                                                         \
        break;
        ASRK_FUSED_CASE(4)
        ASRK_FUSED_CASE(8)
        ASRK_FUSED_CASE(12)
        ASRK_FUSED_CASE(16)
#undef ASRK_FUSED_CASE
        default: break;
    }
}

static void launch_grad(const Params& p, int nv4, cudaStream_t stream) {
    const dim3 grid((unsigned)((p.T + kGradFrames - 1) / kGradFrames), (unsigned)p.B);
    const bool need_lpl = (nv4 == 0) || p.prob;
    const size_t smem = sizeof(int) * 3 * (size_t)p.Ls + sizeof(float) * kRowWarps * (size_t)(p.Ls + (need_lpl ? p.Ls + 1 : 0)) +
                        (nv4 > 0 ? sizeof(float) * kRowWarps * (size_t)p.V : 0);
    switch (nv4) {
#define ASRK_GRAD_CASE(N)                                                                                   \
    case N:                                                                                                 \
        cudaFuncSetAttribute(grad_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        grad_kernel<N><<<grid, kRowWarps * 32, smem, stream>>>(p), asrk::note_launch();                      \
        break;
        ASRK_GRAD_CASE(4)
        ASRK_GRAD_CASE(8)
        ASRK_GRAD_CASE(12)
        ASRK_GRAD_CASE(16)
        default:
            cudaFuncSetAttribute(grad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            grad_kernel<0><<<grid, kRowWarps * 32, smem, stream>>>(p), asrk::note_launch();
            break;
#undef ASRK_GRAD_CASE
    }
}

}  // namespace ctc
}  // namespace asrk

using namespace asrk;
using namespace asrk::ctc;

extern "C" size_t asrk_ctc_workspace_bytes(int T, int B, int label_stride) {
    if (T <= 0 || B <= 0 || label_stride < 0) return 0;
    return ws_layout(T, B, label_stride < 1 ? 1 : label_stride).total;
}

extern "C" int asrk_ctc_fits_fused(int max_input_len, int max_label_len) {
    return (max_input_len >= 0 && max_label_len >= 0 && small_lattice(max_label_len, max_input_len)) ? 1 : 0;
}

extern "C" size_t asrk_ctc_decode_workspace_bytes(int T, int B) {
    if (T <= 0 || B <= 0) return 0;
    return ws_layout(T, B, 1).total;
}

static void bind_workspace(Params& p, void* workspace, const WsLayout& l) {
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    p.eff_labels = reinterpret_cast<int*>(ws + l.eff_labels);
    p.eff_len = reinterpret_cast<int*>(ws + l.eff_len);
    p.chain_next = reinterpret_cast<int*>(ws + l.chain_next);
    p.chain_first = reinterpret_cast<int*>(ws + l.chain_first);
    p.lse = reinterpret_cast<float*>(ws + l.lse);
    p.rowmax = reinterpret_cast<float*>(ws + l.rowmax);
    p.argmax = reinterpret_cast<int*>(ws + l.argmax);
    p.lpl = reinterpret_cast<float*>(ws + l.lpl);
    p.occ = reinterpret_cast<float*>(ws + l.occ);
    p.coff = reinterpret_cast<double*>(ws + l.coff);
    p.beta = reinterpret_cast<float*>(ws + l.beta);
    p.coffb = reinterpret_cast<double*>(ws + l.coffb);
    p.logp = reinterpret_cast<double*>(ws + l.logp);
    p.need_generic = reinterpret_cast<int*>(ws + l.flag);
}

static int run_phases_impl(const float* logits, long long stride_t, long long stride_b,
                           int T, int B, int V, const int* labels, int label_stride,
                           const int* label_len, const int* input_len, int blank,
                           int label_mode, const float* grad_scale, float* loss,
                           float* grad, long long gstride_t, long long gstride_b,
                           int* row_status, int* tokens, int token_stride,
                           int* token_len, float* neg_sum_logits, void* workspace,
                           size_t workspace_bytes, asrk_stream_t stream_, int phases, const Params::ZWork* zwork) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (T < 0 || B < 0 || V < 1 || label_stride < 0) return ASRK_E_BADARG;
    if (B == 0) return ASRK_OK;
    if (T == 0) {
        // no frames at all: every row is rejected like an input_len outside [1, T] (nothing is left undefined)
        if (!loss || !row_status) return ASRK_E_BADARG;
        fill_bad_rows_kernel<<<(B + 255) / 256, 256, 0, stream>>>(loss, row_status, token_len, neg_sum_logits, B), asrk::note_launch();
        return launch_status();
    }
    if (!logits || !labels || !input_len || !loss || !row_status || !workspace) return ASRK_E_BADARG;
    const int prob = (phases & ASRK_CTC_INPUT_PROB) ? 1 : 0;
    if (prob && tokens) return ASRK_E_BADARG;   // the greedy decode is defined on the op's input, not on p
    if (label_mode != ASRK_LABELS_BY_LENGTH && label_mode != ASRK_LABELS_DROP_ZEROS) return ASRK_E_BADARG;
    if (label_mode == ASRK_LABELS_BY_LENGTH && !label_len) return ASRK_E_BADARG;
    if (blank < 0 || blank >= V) return ASRK_E_BADARG;
    if (tokens && (!token_len || token_stride < 1)) return ASRK_E_BADARG;
    if (label_stride + 1 > 1024) return ASRK_E_SHAPE;   // one state pair per thread
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return ASRK_E_WORKSPACE;
    const int Ls = label_stride < 1 ? 1 : label_stride;
    const WsLayout l = ws_layout(T, B, Ls);
    if (workspace_bytes < l.total) return ASRK_E_WORKSPACE;

    Params p{};
    p.logits = logits; p.stride_t = stride_t; p.stride_b = stride_b;
    p.T = T; p.B = B; p.V = V;
    p.labels = labels; p.label_stride = label_stride;
    p.label_len = label_len; p.input_len = input_len;
    p.blank = blank; p.label_mode = label_mode;
    p.grad_scale = grad_scale; p.loss = loss;
    p.grad = grad; p.gstride_t = gstride_t; p.gstride_b = gstride_b;
    p.row_status = row_status;
    p.tokens = tokens; p.token_stride = token_stride; p.token_len = token_len;
    p.neg_sum_logits = neg_sum_logits; p.merge_repeated = 1;
    p.Ls = Ls;
    p.prob = prob;
    p.eps = 1e-7f;                                       // K.epsilon()
    bind_workspace(p, workspace, l);

    const int nv4 = pick_nv4(p, logits, stride_t, stride_b, grad, gstride_t, gstride_b);
    p.fused = (nv4 > 0) ? 1 : 0;
    // the fused kernel prepares its own utterance when both phases are asked for in one call
    p.prep_fused = (p.fused && (phases & ASRK_PHASE_CTC_PREP) && (phases & ASRK_PHASE_CTC_FUSED) &&
                    Ls <= kRowWarps * 32) ? 1 : 0;
    // a batch the caller bounded to small lattices needs none of the generic kernels
    p.small_only = (p.fused && (phases & ASRK_CTC_SMALL_ONLY)) ? 1 : 0;
    if (zwork != nullptr) {
        // the z-score rides on the fused kernel only: the caller must have bounded the batch to small lattices and
        // the vector path must apply; nothing has been launched yet, the caller falls back to the two-kernel path
        if (!p.small_only || !(phases & ASRK_PHASE_CTC_FUSED) || T == 0) return ASRK_E_SHAPE;
        p.z = *zwork;
        // kZBufs chunk buffers behind the utterance-start cache in the kernel's dynamic shared memory
        const size_t smem = sizeof(float) * (kSmallSmemFloats + (size_t)kRowWarps * V);
        long long rows = ((long long)smem - (long long)sizeof(long long) * kZCache) / kZBufs / (kZBins * 4);
        rows = rows > 32 ? 32 : rows - rows % 8;
        if (rows < 8) return ASRK_E_SHAPE;
        p.z.rows = (int)rows;
        // fresh tickets for THIS launch (a second call behind the same spectrogram call must not find them spent:
        // no utterance would be claimed and the outputs would silently keep their old contents)
        if (cudaMemsetAsync(p.z.ticket, 0, 2 * sizeof(int), stream) != cudaSuccess) return ASRK_E_CUDA;
    }
    if (phases & ASRK_PHASE_CTC_PREP) {
        if (!p.small_only && cudaMemsetAsync(p.need_generic, 0, sizeof(int), stream) != cudaSuccess)
            return ASRK_E_CUDA;
        if (!p.prep_fused) {
            const int pt = ((label_stride > 0 ? label_stride : 1) + 31) / 32 * 32;
            prep_kernel<<<B, pt, sizeof(int) * (Ls + 40), stream>>>(p), asrk::note_launch();
        }
    }
    // utterances with a small lattice: one fused CTA each; the row-parallel kernels
    // below skip them and handle the long ones
    if (p.fused && (phases & ASRK_PHASE_CTC_FUSED)) launch_fused(p, nv4, stream);
    if (p.small_only) return launch_status();
    if (phases & ASRK_PHASE_CTC_ROWS) launch_rows<true>(p, nv4, stream);
    int P = ((label_stride + 1) + 31) / 32 * 32;
    if (phases & ASRK_PHASE_CTC_LATTICE) {
        const size_t lsm = sizeof(double) * 98 + sizeof(float) * kRing * lattice_slot_floats(P);
        cudaFuncSetAttribute(lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm);
        lattice_kernel<<<dim3((unsigned)B, grad ? 2u : 1u), P, lsm, stream>>>(p), asrk::note_launch();
    }
    if (grad && (phases & ASRK_PHASE_CTC_GRAD)) launch_grad(p, nv4, stream);
    if (tokens && (phases & ASRK_PHASE_CTC_COLLAPSE)) collapse_kernel<<<(B + 3) / 4, 128, 0, stream>>>(p), asrk::note_launch();
    return launch_status();
}

extern "C" int asrk_ctc_loss_grad_run_phases(const float* logits, long long stride_t, long long stride_b,
                                             int T, int B, int V, const int* labels, int label_stride,
                                             const int* label_len, const int* input_len, int blank,
                                             int label_mode, const float* grad_scale, float* loss,
                                             float* grad, long long gstride_t, long long gstride_b,
                                             int* row_status, int* tokens, int token_stride,
                                             int* token_len, float* neg_sum_logits, void* workspace,
                                             size_t workspace_bytes, asrk_stream_t stream_, int phases) {
    return run_phases_impl(logits, stride_t, stride_b, T, B, V, labels, label_stride, label_len, input_len, blank,
                           label_mode, grad_scale, loss, grad, gstride_t, gstride_b, row_status, tokens, token_stride,
                           token_len, neg_sum_logits, workspace, workspace_bytes, stream_, phases, nullptr);
}

extern "C" int asrk_ctc_loss_grad_zscore_run(const float* logits, long long stride_t, long long stride_b,
                                             int T, int B, int V, const int* labels, int label_stride,
                                             const int* label_len, const int* input_len, int blank,
                                             int label_mode, const float* grad_scale, float* loss,
                                             float* grad, long long gstride_t, long long gstride_b,
                                             int* row_status, int* tokens, int token_stride,
                                             int* token_len, float* neg_sum_logits, void* workspace,
                                             size_t workspace_bytes, asrk_stream_t stream_, int flags,
                                             float* z_features, const float* z_stats, const long long* z_frame_offsets,
                                             const long long* z_row_offsets, int z_batch, long long z_total_frames,
                                             int* z_ticket) {
    if (!z_features || !z_stats || !z_frame_offsets || !z_ticket || z_batch < 1 || z_total_frames < 0) return ASRK_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(z_features) & 15) != 0) return ASRK_E_ALIGN;
    if (B < 1) return ASRK_E_SHAPE;              // no CTC kernel to ride on
    if (z_row_offsets != nullptr) return ASRK_E_SHAPE;   // chunks are bulk copies of consecutive rows: flat layout only
    Params::ZWork z;
    z.feat = z_features; z.stats = z_stats; z.frame_offsets = z_frame_offsets; z.row_offsets = z_row_offsets;
    z.ticket = z_ticket; z.batch = z_batch; z.total_frames = z_total_frames; z.rows = 0;
    return run_phases_impl(logits, stride_t, stride_b, T, B, V, labels, label_stride, label_len, input_len, blank,
                           label_mode, grad_scale, loss, grad, gstride_t, gstride_b, row_status, tokens, token_stride,
                           token_len, neg_sum_logits, workspace, workspace_bytes, stream_,
                           ASRK_PHASE_ALL | (flags & ~0xffff), &z);
}

extern "C" int asrk_ctc_loss_grad_run(const float* logits, long long stride_t, long long stride_b,
                                      int T, int B, int V, const int* labels, int label_stride,
                                      const int* label_len, const int* input_len, int blank,
                                      int label_mode, const float* grad_scale, float* loss, float* grad,
                                      long long gstride_t, long long gstride_b, int* row_status,
                                      int* tokens, int token_stride, int* token_len,
                                      float* neg_sum_logits, void* workspace, size_t workspace_bytes,
                                      asrk_stream_t stream_) {
    return asrk_ctc_loss_grad_run_phases(logits, stride_t, stride_b, T, B, V, labels, label_stride,
                                         label_len, input_len, blank, label_mode, grad_scale, loss, grad,
                                         gstride_t, gstride_b, row_status, tokens, token_stride,
                                         token_len, neg_sum_logits, workspace, workspace_bytes, stream_,
                                         ASRK_PHASE_ALL);
}

extern "C" int asrk_ctc_batch_cost_run(const float* y_pred, long long stride_t, long long stride_b, int T, int B, int V,
                                       const int* labels, int label_stride, const int* label_len,
                                       const int* input_len, const float* grad_scale, float* loss, float* grad,
                                       long long gstride_t, long long gstride_b, int* row_status, void* workspace,
                                       size_t workspace_bytes, asrk_stream_t stream_, int flags) {
    // K.ctc_batch_cost: blank = V - 1, labels masked by label_length, input = log(y_pred + 1e-7)
    return asrk_ctc_loss_grad_run_phases(y_pred, stride_t, stride_b, T, B, V, labels, label_stride, label_len,
                                         input_len, V - 1, ASRK_LABELS_BY_LENGTH, grad_scale, loss, grad,
                                         gstride_t, gstride_b, row_status, nullptr, 0, nullptr, nullptr, workspace,
                                         workspace_bytes, stream_,
                                         ASRK_PHASE_ALL | ASRK_CTC_INPUT_PROB | (flags & ASRK_CTC_SMALL_ONLY));
}

extern "C" int asrk_ctc_greedy_decode_run(const float* logits, long long stride_t, long long stride_b,
                                          int T, int B, int V, const int* input_len, int blank,
                                          int merge_repeated, int* tokens, int token_stride,
                                          int* token_len, float* neg_sum_logits, void* workspace,
                                          size_t workspace_bytes, asrk_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (T < 0 || B < 0 || V < 1) return ASRK_E_BADARG;
    if (B == 0) return ASRK_OK;
    if (!logits || !input_len || !tokens || !token_len || !workspace) return ASRK_E_BADARG;
    if (token_stride < 1) return ASRK_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return ASRK_E_WORKSPACE;
    const int Teff = T < 1 ? 1 : T;
    const WsLayout l = ws_layout(Teff, B, 1);
    if (workspace_bytes < l.total) return ASRK_E_WORKSPACE;
    Params p{};
    p.logits = logits; p.stride_t = stride_t; p.stride_b = stride_b;
    p.T = T; p.B = B; p.V = V;
    p.input_len = input_len; p.blank = blank;
    p.tokens = tokens; p.token_stride = token_stride; p.token_len = token_len;
    p.neg_sum_logits = neg_sum_logits; p.merge_repeated = merge_repeated ? 1 : 0;
    p.Ls = 1;
    bind_workspace(p, workspace, l);
    if (T > 0) {
        const int nv4 = pick_nv4(p, logits, stride_t, stride_b, nullptr, 0, 0);
        launch_rows<false>(p, nv4, stream);
    }
    collapse_kernel<<<(B + 3) / 4, 128, 0, stream>>>(p), asrk::note_launch();
    return launch_status();
}

extern "C" int asrk_ctc_loss_sum_run(const float* loss, const int* row_status, int B, double* out2,
                                     asrk_stream_t stream_) {
    if (B < 0 || !out2 || (B > 0 && !loss)) return ASRK_E_BADARG;
    loss_sum_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(loss, row_status, B, out2, 0), asrk::note_launch();
    return launch_status();
}

extern "C" int asrk_ctc_loss_sum_acc_run(const float* loss, const int* row_status, int B, double* out2,
                                         asrk_stream_t stream_) {
    if (B < 0 || !out2 || (B > 0 && !loss)) return ASRK_E_BADARG;
    loss_sum_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(loss, row_status, B, out2, 1), asrk::note_launch();
    return launch_status();
}

extern "C" int asrk_ctc_unstage_rows_run(const float* src, long long src_stride_t, long long src_stride_b,
                                         float* dst, long long dst_stride_t, long long dst_stride_b,
                                         const int* input_len, int* stale_len, int T, int B, int V,
                                         asrk_stream_t stream_) {
    if (T < 0 || B < 0 || V < 1) return ASRK_E_BADARG;
    if (T == 0 || B == 0) return ASRK_OK;
    if (!src || !dst || !input_len) return ASRK_E_BADARG;
    if (V % 4 != 0 || (src_stride_t % 4) || (src_stride_b % 4) || (dst_stride_t % 4) || (dst_stride_b % 4)) return ASRK_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return ASRK_E_ALIGN;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    unstage_rows_kernel<<<sm_count() * ASRK_STAGE_CTAS_PER_SM, 256, 0, stream>>>(
        src, src_stride_t, src_stride_b, dst, dst_stride_t, dst_stride_b, input_len, stale_len, T, B, V), asrk::note_launch();
    if (stale_len != nullptr)
        copy_lengths_kernel<<<(B + 255) / 256, 256, 0, stream>>>(input_len, stale_len, B), asrk::note_launch();
    return launch_status();
}

extern "C" int asrk_ctc_stage_logits_run(const float* src, long long src_stride_t, long long src_stride_b,
                                         float* dst, long long dst_stride_t, long long dst_stride_b,
                                         const int* input_len, int T, int B, int V, asrk_stream_t stream_) {
    if (T < 0 || B < 0 || V < 1) return ASRK_E_BADARG;
    if (T == 0 || B == 0) return ASRK_OK;
    if (!src || !dst || !input_len) return ASRK_E_BADARG;
    if (V % 4 != 0 || (src_stride_t % 4) || (src_stride_b % 4) || (dst_stride_t % 4) || (dst_stride_b % 4)) return ASRK_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return ASRK_E_ALIGN;
    stage_logits_kernel<<<sm_count() * ASRK_STAGE_CTAS_PER_SM, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
        src, src_stride_t, src_stride_b, dst, dst_stride_t, dst_stride_b, input_len, T, B, V), asrk::note_launch();
    return launch_status();
}
