/* CPU ORACLE (test infrastructure, NOT a product path).
 *
 * Canonical-arithmetic restatement of utils/tf_util.py:647-666 (pairwise_distance_mask) and
 * utils/tf_util.py:577-610 (pairwise_distance, knn) of the reference.
 *
 *   inner_ij = p_i . p_j                      tf.matmul(pc, pc^T)              :651-652
 *   inner2   = -2 * inner                                                      :653
 *   s_i      = sum(square(p_i))                                                :654
 *   a_ij     = -((s_i + inner2_ij) + s_j)                                      :656
 *   kth_i    = min(top_k(a_i, 20).values)      literal 20                      :660-663
 *   mask_ij  = (a_ij >= kth_i)                                                 :664-665
 *
 * TensorFlow itself cannot run here, so the fp32 evaluation order inside the
 * K=3 matmul is a RECONSTRUCTION, selectable with `arith`:
 *   arith 0 ("muladd", default): ((x x' + y y') + z z'), every op rounded separately -- the
 *            stock TF-1.12 CPU wheel (AVX, no FMA; Eigen GEBP accumulates k = 0,1,2 in order).
 *   arith 1 ("fma"): fma(z, z', fma(y, y', x x'))  -- cuBLAS/FFMA style, the TF GPU path.
 * s_i is (x^2 + y^2) + z^2 with separately rounded squares in both modes (Square, then Sum).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (see oracle/Makefile).  -ffp-contract=off is
 * REQUIRED: it stops gcc from fusing the muladd mode into FMAs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float inner3(const float *p, const float *q, int arith) {
    if (arith == 0) {
        float a = p[0] * q[0];
        float b = p[1] * q[1];
        float c = p[2] * q[2];
        float ab = a + b;
        return ab + c;
    } else {
        float acc = p[0] * q[0];
        acc = fmaf(p[1], q[1], acc);
        acc = fmaf(p[2], q[2], acc);
        return acc;
    }
}

static inline float sq3(const float *p) {
    float a = p[0] * p[0];
    float b = p[1] * p[1];
    float c = p[2] * p[2];
    float ab = a + b;
    return ab + c;
}

/* row i of a (length N) for one cloud */
static void a_row(const float *pc, const float *s, int N, int i, int arith, float *a) {
    const float *pi = pc + 3 * (size_t)i;
    for (int j = 0; j < N; ++j) {
        float inner = inner3(pi, pc + 3 * (size_t)j, arith);
        float inner2 = -2.0f * inner;
        float t = s[i] + inner2;
        float u = t + s[j];
        a[j] = -u;
    }
}

typedef struct { float v; int32_t j; } cand_t;

/* Streaming top-k in tf.nn.top_k order: descending value, ties -> lower index first.
 * j ascends, so a candidate equal to the current k-th value loses the tie (strict >). */
static void topk_row(const float *a, int N, int k, cand_t *top) {
    int n = 0;
    for (int j = 0; j < N; ++j) {
        float v = a[j];
        if (n == k && !(v > top[k - 1].v)) continue;
        int pos = (n < k) ? n : k - 1;
        while (pos > 0 && v > top[pos - 1].v) { top[pos] = top[pos - 1]; --pos; }
        top[pos].v = v; top[pos].j = j;
        if (n < k) ++n;
    }
}

/* Per row: sorted top-k indices (tf.nn.top_k order), the k-th largest a, and the number of
 * members of the thresholded set {j : a_ij >= kth_i}.
 *   pc    [B,N,3] fp32
 *   idx   [B,N,k] int32    (may be NULL)
 *   kth   [B,N]   fp32     (may be NULL)  value of a (i.e. minus the distance expression)
 *   count [B,N]   int32    (may be NULL)
 */
int epc_oracle_knn(const float *pc, int B, int N, int k, int arith, int32_t *idx, float *kth, int32_t *count) {
    if (k < 1 || k > N) return -1;
    float *s = (float *)malloc(sizeof(float) * (size_t)N);
    float *a = (float *)malloc(sizeof(float) * (size_t)N);
    cand_t *c = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
    if (!s || !a || !c) return -2;
    for (int b = 0; b < B; ++b) {
        const float *p = pc + (size_t)b * N * 3;
        for (int i = 0; i < N; ++i) s[i] = sq3(p + 3 * (size_t)i);
        for (int i = 0; i < N; ++i) {
            a_row(p, s, N, i, arith, a);
            topk_row(a, N, k, c);
            float th = c[k - 1].v;
            size_t row = (size_t)b * N + i;
            if (idx) for (int t = 0; t < k; ++t) idx[row * k + t] = c[t].j;
            if (kth) kth[row] = th;
            if (count) {
                int n = 0;
                for (int j = 0; j < N; ++j) n += (a[j] >= th);
                count[row] = n;
            }
        }
    }
    free(s); free(a); free(c);
    return 0;
}

/* Dense (B,N,N) outputs: a (may be NULL) and the 0/1 fp32 mask (may be NULL).  The caller passes
 * k = 20 to mirror the literal at utils/tf_util.py:660. */
int epc_oracle_mask(const float *pc, int B, int N, int k, int arith, float *a_out, float *mask_out) {
    if (k < 1 || k > N) return -1;
    float *s = (float *)malloc(sizeof(float) * (size_t)N);
    float *a = (float *)malloc(sizeof(float) * (size_t)N);
    cand_t *c = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
    if (!s || !a || !c) return -2;
    for (int b = 0; b < B; ++b) {
        const float *p = pc + (size_t)b * N * 3;
        for (int i = 0; i < N; ++i) s[i] = sq3(p + 3 * (size_t)i);
        for (int i = 0; i < N; ++i) {
            a_row(p, s, N, i, arith, a);
            size_t row = ((size_t)b * N + i) * (size_t)N;
            if (a_out) memcpy(a_out + row, a, sizeof(float) * (size_t)N);
            if (mask_out) {
                topk_row(a, N, k, c);
                float th = c[k - 1].v;
                for (int j = 0; j < N; ++j) mask_out[row + j] = (a[j] >= th) ? 1.0f : 0.0f;
            }
        }
    }
    free(s); free(a); free(c);
    return 0;
}
