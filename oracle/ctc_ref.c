/* Plain-C restatement of TensorFlow's CTC loss / gradient (CTCLossOp) and greedy
 * decoder (CTCGreedyDecoderOp) -- the CPU kernels the reference reaches through
 * K.ctc_batch_cost (lm_and_am/model/cnn_ctc.py:149-152), tf.nn.ctc_loss_v2
 * (lm_and_am/model/acoustic_model2.py:79-80) and tf.nn.ctc_greedy_decoder
 * (acoustic_model2.py:69).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used as the checker at sizes
 * where the numpy oracle is too slow and as the timed CPU baseline of bench.py.
 * PARITY UNPINNED against the reference itself (TensorFlow is not vendored and
 * cannot be installed); pinned against oracle/ctc_ref.py and torch float64.
 *
 * Follows the published algorithm of tensorflow/core/util/ctc/
 * ctc_loss_calculator.{h,cc} (TF 1.14): per utterance softmax over the classes,
 * l' = blank-interleaved labels, log-space forward (alpha, includes y_t) and
 * backward (beta, excludes y_t) variables over the window
 * [max(0, U - 2 (T - t)), min(U, 2 (t + 1))), log p = LSE_u alpha(u,0)+beta(u,0),
 * dy = y - exp(LSE_{u: l'_u = v}(alpha + beta) - log p).  Like TF, utterances are
 * sharded over threads (OpenMP here, Eigen's pool there); the arithmetic type is
 * float in TF (CTC_REAL=float) -- the double build is for checking.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef CTC_REAL
#define CTC_REAL float
#endif
#ifndef CTC_NAME
#define CTC_NAME(x) x##_f32
#endif
typedef CTC_REAL real;

static const real kLogZero = (real)-INFINITY;

static inline real lse(real a, real b) {
    /* ctc_loss_util.h LogSumExp */
    if (a == kLogZero) return b;
    if (b == kLogZero) return a;
    return a > b ? a + (real)log1p(exp((double)(b - a))) : b + (real)log1p(exp((double)(a - b)));
}

/* logits [T,B,V] time-major; labels int32 [B,Lmax]; returns 0.  loss[b]=+inf and
 * grad = y when no valid path exists.  status[b]: 0 ok, 1 infeasible. */
int CTC_NAME(ctc_loss_grad)(const float* logits, int T, int B, int V, const int* labels, int Lmax,
                            const int* label_len, const int* input_len, int blank, float* loss,
                            float* grad, int* status, int n_threads) {
    if (n_threads > 0) {
#ifdef _OPENMP
        extern void omp_set_num_threads(int);
        omp_set_num_threads(n_threads);
#endif
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const int Tb = input_len[b];
        const int L = label_len[b];
        const int U = 2 * L + 1;
        int* lp = (int*)malloc(sizeof(int) * U);
        real* y = (real*)malloc(sizeof(real) * (size_t)Tb * V);     /* softmax */
        real* ly = (real*)malloc(sizeof(real) * (size_t)Tb * U);    /* log y_t(l'_u) */
        real* alpha = (real*)malloc(sizeof(real) * (size_t)Tb * U);
        real* beta = (real*)malloc(sizeof(real) * (size_t)Tb * U);
        real* acc = (real*)malloc(sizeof(real) * V);
        for (int u = 0; u < U; ++u) lp[u] = (u & 1) ? labels[(size_t)b * Lmax + u / 2] : blank;
        for (int t = 0; t < Tb; ++t) {
            const float* x = logits + ((size_t)t * B + b) * V;
            real m = x[0];
            for (int v = 1; v < V; ++v) if (x[v] > m) m = x[v];
            real s = 0;
            for (int v = 0; v < V; ++v) { real e = (real)exp((double)(x[v] - m)); y[(size_t)t * V + v] = e; s += e; }
            for (int v = 0; v < V; ++v) y[(size_t)t * V + v] /= s;
            for (int u = 0; u < U; ++u) ly[(size_t)t * U + u] = (real)log((double)y[(size_t)t * V + lp[u]]);
        }
        for (size_t k = 0; k < (size_t)Tb * U; ++k) { alpha[k] = kLogZero; beta[k] = kLogZero; }
        /* forward */
        alpha[0] = ly[0];
        if (U > 1) alpha[1] = ly[1];
        for (int t = 1; t < Tb; ++t) {
            int lo = U - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > U) hi = U;
            for (int u = lo; u < hi; ++u) {
                real s = alpha[(size_t)(t - 1) * U + u];
                if (u >= 1) s = lse(s, alpha[(size_t)(t - 1) * U + u - 1]);
                if (u >= 2 && lp[u] != blank && lp[u] != lp[u - 2]) s = lse(s, alpha[(size_t)(t - 1) * U + u - 2]);
                alpha[(size_t)t * U + u] = (s == kLogZero) ? kLogZero : s + ly[(size_t)t * U + u];
            }
        }
        /* backward */
        beta[(size_t)(Tb - 1) * U + U - 1] = 0;
        if (U > 1) beta[(size_t)(Tb - 1) * U + U - 2] = 0;
        for (int t = Tb - 2; t >= 0; --t) {
            int lo = U - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > U) hi = U;
            for (int u = lo; u < hi; ++u) {
                const real* bn = beta + (size_t)(t + 1) * U;
                const real* ln = ly + (size_t)(t + 1) * U;
                real s = (bn[u] == kLogZero) ? kLogZero : bn[u] + ln[u];
                if (u + 1 < U && bn[u + 1] != kLogZero) s = lse(s, bn[u + 1] + ln[u + 1]);
                if (u + 2 < U && lp[u] != blank && lp[u] != lp[u + 2] && bn[u + 2] != kLogZero)
                    s = lse(s, bn[u + 2] + ln[u + 2]);
                beta[(size_t)t * U + u] = s;
            }
        }
        real log_p = kLogZero;
        for (int u = 0; u < U; ++u) {
            real a = alpha[u], be = beta[u];
            if (a != kLogZero && be != kLogZero) log_p = lse(log_p, a + be);
        }
        const int ok = (log_p != kLogZero);
        loss[b] = ok ? (float)(-log_p) : INFINITY;
        if (status) status[b] = ok ? 0 : 1;
        if (grad) {
            for (int t = 0; t < T; ++t) {
                float* g = grad + ((size_t)t * B + b) * V;
                if (t >= Tb) { memset(g, 0, sizeof(float) * V); continue; }
                if (ok) {
                    for (int v = 0; v < V; ++v) acc[v] = kLogZero;
                    for (int u = 0; u < U; ++u) {
                        real a = alpha[(size_t)t * U + u], be = beta[(size_t)t * U + u];
                        if (a != kLogZero && be != kLogZero) acc[lp[u]] = lse(acc[lp[u]], a + be);
                    }
                    for (int v = 0; v < V; ++v) {
                        real o = (acc[v] == kLogZero) ? 0 : (real)exp((double)(acc[v] - log_p));
                        g[v] = (float)(y[(size_t)t * V + v] - o);
                    }
                } else {
                    for (int v = 0; v < V; ++v) g[v] = (float)y[(size_t)t * V + v];
                }
            }
        }
        free(lp); free(y); free(ly); free(alpha); free(beta); free(acc);
    }
    return 0;
}

/* ctc_decoder_ops.cc CTCGreedyDecoderOp: first maximum per frame, merge repeats,
 * drop blank; tokens [B,T]; neg_sum_logits [B]. */
int CTC_NAME(ctc_greedy_decode)(const float* logits, int T, int B, int V, const int* input_len,
                                int blank, int merge_repeated, int* tokens, int* token_len,
                                float* neg_sum_logits, int n_threads) {
    if (n_threads > 0) {
#ifdef _OPENMP
        extern void omp_set_num_threads(int);
        omp_set_num_threads(n_threads);
#endif
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int prev = -1, n = 0;
        float lp = 0.f;
        for (int t = 0; t < input_len[b]; ++t) {
            const float* x = logits + ((size_t)t * B + b) * V;
            int am = 0;
            float m = x[0];
            for (int v = 1; v < V; ++v) if (x[v] > m) { m = x[v]; am = v; }
            lp += -m;
            if (am != blank && !(merge_repeated && am == prev)) tokens[(size_t)b * T + n++] = am;
            prev = am;
        }
        token_len[b] = n;
        if (neg_sum_logits) neg_sum_logits[b] = lp;
    }
    return 0;
}
