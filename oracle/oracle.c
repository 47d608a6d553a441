/*
 * TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of the ptgnn message-passing hot path.
 *
 * Second, independent restatement (next to oracle/ptgnn_oracle.py) used to check the integer
 * bookkeeping bit-exactly and the floating-point path in strict edge order.  Scalar, one thread.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built
 * from this file (oracle/_build/liboracle.so); the product never links it.
 *
 * Follows (paths relative to /root/reference/ptgnn):
 *   oracle_edge_plan        no reference counterpart -- canonical stable target-sorted CSR of
 *                           cat(adjacency lists) (gatedmessagepassing.py:46 + the grouping done inside
 *                           torch_scatter.scatter); contract of the CUDA plan builder.
 *   oracle_scatter_f32      neuralmodels/gnn/messagepassing/abstractmessagepassing.py:44-50 ->
 *                           torch_scatter.scatter (third-party, torch-scatter>=2.0.5, setup.py:23;
 *                           published semantics restated: SURVEY.md Appendix A).
 *   oracle_gated_forward_f32   neuralmodels/gnn/messagepassing/gatedmessagepassing.py:37-69 (+ nn.GRUCell)
 *   oracle_mlp_forward_f32     neuralmodels/gnn/messagepassing/mlpmessagepassing.py:68-117, mlp.py:50-80
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { RED_SUM = 0, RED_MEAN = 1, RED_MAX = 2, RED_MIN = 3 };
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_TANH = 2, ACT_RELU = 3 };

/* ---------------------------------------------------------------- edge plan (bit-exact contract) */
int oracle_edge_plan(int64_t num_nodes, int32_t num_types, const int64_t *const *src, const int64_t *const *tgt,
                     const int64_t *counts, int32_t *row_ptr /*[N+1]*/, int32_t *perm /*[E]*/, int32_t *pos /*[E]*/,
                     int32_t *src_sorted /*[E]*/, uint8_t *etype_sorted /*[E]*/) {
    int64_t E = 0;
    for (int t = 0; t < num_types; ++t) E += counts[t];
    if (E >= INT32_MAX || num_nodes >= INT32_MAX || num_types > 255) return -1;
    memset(row_ptr, 0, sizeof(int32_t) * (size_t)(num_nodes + 1));
    for (int t = 0; t < num_types; ++t)
        for (int64_t i = 0; i < counts[t]; ++i) {
            int64_t v = tgt[t][i], s = src[t][i];
            if (v < 0 || v >= num_nodes || s < 0 || s >= num_nodes) return -2;
            row_ptr[v + 1]++;
        }
    for (int64_t v = 0; v < num_nodes; ++v) row_ptr[v + 1] += row_ptr[v];
    int32_t *cursor = (int32_t *)malloc(sizeof(int32_t) * (size_t)(num_nodes + 1));
    if (!cursor) return -3;
    memcpy(cursor, row_ptr, sizeof(int32_t) * (size_t)(num_nodes + 1));
    int64_t e = 0;
    for (int t = 0; t < num_types; ++t)
        for (int64_t i = 0; i < counts[t]; ++i, ++e) { /* edge-id order => stable */
            int32_t j = cursor[tgt[t][i]]++;
            perm[j] = (int32_t)e;
            pos[e] = j;
            src_sorted[j] = (int32_t)src[t][i];
            etype_sorted[j] = (uint8_t)t;
        }
    free(cursor);
    return 0;
}

/* ---------------------------------------------------------------- torch_scatter.scatter, dim=0 */
int oracle_scatter_f32(const float *src /*[E,D]*/, const int64_t *index /*[E]*/, int64_t E, int64_t D, int64_t N,
                       int reduce, float *out /*[N,D]*/, int64_t *arg /*[N,D] or NULL*/) {
    if (reduce == RED_SUM || reduce == RED_MEAN) {
        memset(out, 0, sizeof(float) * (size_t)(N * D));
        for (int64_t e = 0; e < E; ++e) {
            if (index[e] < 0 || index[e] >= N) return -2;
            float *o = out + index[e] * D;
            const float *s = src + e * D;
            for (int64_t d = 0; d < D; ++d) o[d] += s[d];
        }
        if (reduce == RED_MEAN) {
            float *cnt = (float *)calloc((size_t)N, sizeof(float));
            if (!cnt) return -3;
            for (int64_t e = 0; e < E; ++e) cnt[index[e]] += 1.0f;
            for (int64_t v = 0; v < N; ++v) {
                float c = cnt[v] < 1.0f ? 1.0f : cnt[v];
                for (int64_t d = 0; d < D; ++d) out[v * D + d] /= c;
            }
            free(cnt);
        }
        return 0;
    }
    if (reduce != RED_MAX && reduce != RED_MIN) return -1;
    const float init = reduce == RED_MAX ? -FLT_MAX : FLT_MAX; /* numeric_limits::lowest()/max() */
    int64_t *a = arg ? arg : (int64_t *)malloc(sizeof(int64_t) * (size_t)(N * D));
    if (!a) return -3;
    for (int64_t i = 0; i < N * D; ++i) { out[i] = init; a[i] = E; }
    for (int64_t e = 0; e < E; ++e) {
        if (index[e] < 0 || index[e] >= N) { if (!arg) free(a); return -2; }
        float *o = out + index[e] * D;
        int64_t *oa = a + index[e] * D;
        const float *s = src + e * D;
        for (int64_t d = 0; d < D; ++d) {
            int better = reduce == RED_MAX ? (s[d] > o[d]) : (s[d] < o[d]); /* strict; NaN never wins */
            if (better) { o[d] = s[d]; oa[d] = e; }
        }
    }
    for (int64_t i = 0; i < N * D; ++i) if (a[i] == E) out[i] = 0.0f; /* untouched -> 0 */
    if (!arg) free(a);
    return 0;
}

/* ---------------------------------------------------------------- helpers */
static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
static float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
static float act(float x, int kind) {
    switch (kind) {
        case ACT_GELU: return gelu_erf(x);
        case ACT_TANH: return tanhf(x);
        case ACT_RELU: return x > 0.0f ? x : 0.0f;
        default: return x;
    }
}
/* y[D] = W[D,K] * x[K]   (nn.Linear without bias: y = x W^T) */
static void matvec(const float *W, const float *x, int64_t D, int64_t K, float *y) {
    for (int64_t d = 0; d < D; ++d) {
        float acc = 0.0f;
        const float *w = W + d * K;
        for (int64_t k = 0; k < K; ++k) acc += w[k] * x[k];
        y[d] = acc;
    }
}

/* messages [E,D] in edge-id order.  use_target: input = [h_src ; h_tgt] (Mlp) else [h_src] (Gated). */
static int build_messages(const float *h, int64_t H, int32_t T, const int64_t *const *src, const int64_t *const *tgt,
                          const int64_t *counts, const float *const *W /*T x [D,K]*/, int64_t D, int use_target,
                          float *msg, int64_t *targets) {
    int64_t K = use_target ? 2 * H : H;
    float *x = (float *)malloc(sizeof(float) * (size_t)K);
    if (!x) return -3;
    int64_t e = 0;
    for (int t = 0; t < T; ++t)
        for (int64_t i = 0; i < counts[t]; ++i, ++e) {
            memcpy(x, h + src[t][i] * H, sizeof(float) * (size_t)H);
            if (use_target) memcpy(x + H, h + tgt[t][i] * H, sizeof(float) * (size_t)H);
            matvec(W[t], x, D, K, msg + e * D);
            targets[e] = tgt[t][i];
        }
    free(x);
    return 0;
}

/* ---------------------------------------------------------------- GatedMessagePassingLayer.forward */
int oracle_gated_forward_f32(const float *h /*[N,H]*/, int64_t N, int64_t H, int64_t D, int32_t T,
                             const int64_t *const *src, const int64_t *const *tgt, const int64_t *counts,
                             const float *const *W /*T x [D,H]*/, const float *w_ih /*[3H,D]*/,
                             const float *w_hh /*[3H,H]*/, const float *b_ih /*[3H]*/, const float *b_hh /*[3H]*/,
                             int reduce, float *out /*[N,H]*/) {
    int64_t E = 0;
    for (int t = 0; t < T; ++t) E += counts[t];
    float *msg = (float *)malloc(sizeof(float) * (size_t)(E * D + 1));
    int64_t *targets = (int64_t *)malloc(sizeof(int64_t) * (size_t)(E + 1));
    float *agg = (float *)malloc(sizeof(float) * (size_t)(N * D + 1));
    float *gi = (float *)malloc(sizeof(float) * (size_t)(3 * H));
    float *gh = (float *)malloc(sizeof(float) * (size_t)(3 * H));
    if (!msg || !targets || !agg || !gi || !gh) return -3;
    int rc = build_messages(h, H, T, src, tgt, counts, W, D, 0, msg, targets);
    if (!rc) rc = oracle_scatter_f32(msg, targets, E, D, N, reduce, agg, NULL);
    if (!rc)
        for (int64_t v = 0; v < N; ++v) {
            matvec(w_ih, agg + v * D, 3 * H, D, gi);
            matvec(w_hh, h + v * H, 3 * H, H, gh);
            for (int64_t j = 0; j < H; ++j) {
                float r = sigmoidf_(gi[j] + b_ih[j] + gh[j] + b_hh[j]);
                float z = sigmoidf_(gi[H + j] + b_ih[H + j] + gh[H + j] + b_hh[H + j]);
                float n = tanhf(gi[2 * H + j] + b_ih[2 * H + j] + r * (gh[2 * H + j] + b_hh[2 * H + j]));
                out[v * H + j] = (1.0f - z) * n + z * h[v * H + j];
            }
        }
    free(msg); free(targets); free(agg); free(gi); free(gh);
    return rc;
}

/* ---------------------------------------------------------------- MlpMessagePassingLayer.forward
 * default message MLP (mlp_hidden_layers = 0 => one bias-free Linear), string aggregator. */
int oracle_mlp_forward_f32(const float *h /*[N,H]*/, int64_t N, int64_t H, int64_t D, int64_t Hout, int32_t T,
                           const int64_t *const *src, const int64_t *const *tgt, const int64_t *counts,
                           const float *const *W /*T x [D,K]*/, int use_target, int reduce, int msg_act,
                           const float *ln_w /*[D] or NULL*/, const float *ln_b, float ln_eps,
                           const float *dense_w /*[Hout,D] or NULL*/, const float *dense_b, int dense_act,
                           float *out /*[N,Hout] (or [N,D] without dense)*/) {
    int64_t E = 0;
    for (int t = 0; t < T; ++t) E += counts[t];
    float *msg = (float *)malloc(sizeof(float) * (size_t)(E * D + 1));
    int64_t *targets = (int64_t *)malloc(sizeof(int64_t) * (size_t)(E + 1));
    float *agg = (float *)malloc(sizeof(float) * (size_t)(N * D + 1));
    float *y = (float *)malloc(sizeof(float) * (size_t)(Hout + 1));
    if (!msg || !targets || !agg || !y) return -3;
    int rc = build_messages(h, H, T, src, tgt, counts, W, D, use_target, msg, targets);
    if (!rc) rc = oracle_scatter_f32(msg, targets, E, D, N, reduce, agg, NULL);
    if (!rc)
        for (int64_t v = 0; v < N; ++v) {
            float *a = agg + v * D;
            for (int64_t d = 0; d < D; ++d) a[d] = act(a[d], msg_act);
            if (ln_w) {
                float mean = 0.0f, var = 0.0f;
                for (int64_t d = 0; d < D; ++d) mean += a[d];
                mean /= (float)D;
                for (int64_t d = 0; d < D; ++d) var += (a[d] - mean) * (a[d] - mean);
                var /= (float)D;
                float rstd = 1.0f / sqrtf(var + ln_eps);
                for (int64_t d = 0; d < D; ++d) a[d] = (a[d] - mean) * rstd * ln_w[d] + ln_b[d];
            }
            if (dense_w) {
                matvec(dense_w, a, Hout, D, y);
                for (int64_t j = 0; j < Hout; ++j) out[v * Hout + j] = act(y[j] + (dense_b ? dense_b[j] : 0.0f), dense_act);
            } else {
                memcpy(out + v * D, a, sizeof(float) * (size_t)D);
            }
        }
    free(msg); free(targets); free(agg); free(y);
    return rc;
}
