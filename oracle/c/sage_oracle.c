/* GraphSAGE_T.forward -- plain C + OpenMP restatement of the frozen spec.  TEST INFRASTRUCTURE.
 *
 * Second, independent witness of oracle/sage_ref.py (which is PyTorch index_select/index_add_/addmm)
 * and the all-host-cores CPU arm of bench.py (`--impl reference`, cpu_baseline).  Never linked,
 * imported or executed by the product (nerrf_b200/).
 *
 * Spec source in the reference (prose only; ai/models/GraphSAGE-T.py is named in README.md:73 and
 * ROADMAP.md:127 but does not exist -- SURVEY.md section 0):
 *   node anomaly_score in [0,1]                         docs/content/docs/architecture.mdx:157
 *   "classify edges as normal/attack"                   docs/content/docs/architecture.mdx:49-53
 *   edge weight = causality confidence, 30-60 s window  docs/content/docs/architecture.mdx:40-42
 * Frozen spec v0:
 *   m_v  = sum_{e: dst(e)=v} w_e * h_src(e) / max(sum_e w_e, 1e-12)      (isolated node -> 0)
 *   h'_v = ReLU([h_v || m_v] @ W_l + b_l),  W_l in R^{2F_l x H}
 *   node_score_v = sigmoid(h_v . w_n + b_n)
 *
 * Association order (documented so that a third implementation can match it):
 *   aggregate  per row, edges in ascending CSR order, acc[f] += w_e * x[src][f] (one rounded multiply and one
 *              rounded add when built with -ffp-contract=off; a fused multiply-add otherwise), wsum += w_e,
 *              mean = acc / max(wsum, 1e-12)  (a division, like sage_ref.py)
 *   transform  z[h] = b[h], then for k = 0 .. 2F-1 ascending: z[h] += a[k] * W[k][h]
 *   head       dot over h ascending in fp32, sigmoid = 1 / (1 + expf(-t))
 * Rows are independent, so the OpenMP schedule does not change any result.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HC 32 /* output columns per register block */
#define RB 8 /* destination rows per block: the block's [RB, 2F] operand and [RB, H] result stay in L1 */

static inline int64_t rp_at(const void* rowptr, int is64, int64_t i) {
    return is64 ? ((const int64_t*)rowptr)[i] : (int64_t)((const int32_t*)rowptr)[i];
}

void nerrf_oracle_sage_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int nerrf_oracle_sage_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* out is [row_end - row_begin, H] (row r of out = node row_begin + r). */
int nerrf_oracle_sage_layer(const float* x, int64_t n_nodes, int F, const void* rowptr, int rowptr_is64,
                            const int32_t* col, const float* ew, const float* W, const float* b, int H, int relu,
                            int64_t row_begin, int64_t row_end, float* out) {
    if (F <= 0 || H <= 0 || row_begin < 0 || row_end > n_nodes || row_begin > row_end) return -1;
    const int K = 2 * F;
    const int64_t n_blocks = (row_end - row_begin + RB - 1) / RB;
    int bad = 0;
#pragma omp parallel reduction(| : bad)
    {
        float* a = (float*)malloc(sizeof(float) * (size_t)RB * K);
        float* z = (float*)malloc(sizeof(float) * (size_t)RB * H);
#pragma omp for schedule(dynamic, 64)
        for (int64_t blk = 0; blk < n_blocks; ++blk) {
            const int64_t r0 = row_begin + blk * RB;
            const int nr = (int)((row_end - r0) < RB ? (row_end - r0) : RB);
            for (int r = 0; r < nr; ++r) {
                const int64_t v = r0 + r;
                float* av = a + (size_t)r * K;
                float* acc = av + F;
                memcpy(av, x + (size_t)v * F, sizeof(float) * F);
                for (int f = 0; f < F; ++f) acc[f] = 0.f;
                float wsum = 0.f;
                const int64_t e0 = rp_at(rowptr, rowptr_is64, v), e1 = rp_at(rowptr, rowptr_is64, v + 1);
                for (int64_t e = e0; e < e1; ++e) {
                    const int32_t s = col[e];
                    if (s < 0 || s >= n_nodes) { bad = 1; continue; }
                    const float w = ew[e];
                    const float* xs = x + (size_t)s * F;
#pragma omp simd
                    for (int f = 0; f < F; ++f) acc[f] += w * xs[f];
                    wsum += w;
                }
                const float den = wsum > 1e-12f ? wsum : 1e-12f;
                for (int f = 0; f < F; ++f) acc[f] = acc[f] / den;
            }
            /* z[r][h] = b[h] + sum_k a[r][k] * W[k][h], k ascending.  Register blocking: 4 rows x HC output columns
             * are accumulated in local arrays (vector registers) over the whole k loop; this changes which loads are
             * shared, not the per-element association order. */
            for (int rq = 0; rq < nr; rq += 4) {
                const int rn = (nr - rq) < 4 ? (nr - rq) : 4;
                const float* a0 = a + (size_t)(rq + 0) * K;
                const float* a1 = a + (size_t)(rq + (rn > 1 ? 1 : 0)) * K;
                const float* a2 = a + (size_t)(rq + (rn > 2 ? 2 : 0)) * K;
                const float* a3 = a + (size_t)(rq + (rn > 3 ? 3 : 0)) * K;
                int hc = 0;
                for (; hc + HC <= H; hc += HC) {
                    float c0[HC], c1[HC], c2[HC], c3[HC];
                    for (int j = 0; j < HC; ++j) { c0[j] = b[hc + j]; c1[j] = b[hc + j]; c2[j] = b[hc + j]; c3[j] = b[hc + j]; }
                    for (int k = 0; k < K; ++k) {
                        const float* wk = W + (size_t)k * H + hc;
                        const float x0 = a0[k], x1 = a1[k], x2 = a2[k], x3 = a3[k];
#pragma omp simd
                        for (int j = 0; j < HC; ++j) {
                            const float w = wk[j];
                            c0[j] += x0 * w; c1[j] += x1 * w; c2[j] += x2 * w; c3[j] += x3 * w;
                        }
                    }
                    memcpy(z + (size_t)(rq + 0) * H + hc, c0, sizeof(c0));
                    if (rn > 1) memcpy(z + (size_t)(rq + 1) * H + hc, c1, sizeof(c1));
                    if (rn > 2) memcpy(z + (size_t)(rq + 2) * H + hc, c2, sizeof(c2));
                    if (rn > 3) memcpy(z + (size_t)(rq + 3) * H + hc, c3, sizeof(c3));
                }
                for (; hc < H; ++hc) {                                   /* H not a multiple of HC */
                    for (int r = 0; r < rn; ++r) {
                        const float* ar = a + (size_t)(rq + r) * K;
                        float c = b[hc];
                        for (int k = 0; k < K; ++k) c += ar[k] * W[(size_t)k * H + hc];
                        z[(size_t)(rq + r) * H + hc] = c;
                    }
                }
            }
            for (int r = 0; r < nr; ++r) {
                float* o = out + (size_t)(r0 + r - row_begin) * H;
                const float* zr = z + (size_t)r * H;
                if (relu) { for (int h = 0; h < H; ++h) o[h] = zr[h] > 0.f ? zr[h] : 0.f; }
                else memcpy(o, zr, sizeof(float) * H);
            }
        }
        free(a); free(z);
    }
    return bad ? -2 : 0;
}

/* m is [row_end - row_begin, F]: the weighted mean alone (K1). */
int nerrf_oracle_sage_aggregate(const float* x, int64_t n_nodes, int F, const void* rowptr, int rowptr_is64,
                                const int32_t* col, const float* ew, int64_t row_begin, int64_t row_end, float* m) {
    if (F <= 0 || row_begin < 0 || row_end > n_nodes || row_begin > row_end) return -1;
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 512) reduction(| : bad)
    for (int64_t v = row_begin; v < row_end; ++v) {
        float* acc = m + (size_t)(v - row_begin) * F;
        for (int f = 0; f < F; ++f) acc[f] = 0.f;
        float wsum = 0.f;
        const int64_t e0 = rp_at(rowptr, rowptr_is64, v), e1 = rp_at(rowptr, rowptr_is64, v + 1);
        for (int64_t e = e0; e < e1; ++e) {
            const int32_t s = col[e];
            if (s < 0 || s >= n_nodes) { bad = 1; continue; }
            const float w = ew[e];
            const float* xs = x + (size_t)s * F;
#pragma omp simd
            for (int f = 0; f < F; ++f) acc[f] += w * xs[f];
            wsum += w;
        }
        const float den = wsum > 1e-12f ? wsum : 1e-12f;
        for (int f = 0; f < F; ++f) acc[f] = acc[f] / den;
    }
    return bad ? -2 : 0;
}

int nerrf_oracle_sage_node_head(const float* h, int64_t n_rows, int H, const float* node_w, float node_b, float* score) {
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < n_rows; ++v) {
        const float* hv = h + (size_t)v * H;
        float t = 0.f;
        for (int k = 0; k < H; ++k) t += hv[k] * node_w[k];
        t += node_b;
        score[v] = 1.0f / (1.0f + expf(-t));
    }
    return 0;
}

/* Whole forward: L layers (W[l] is [2F_l, H], F_0 = f_in, F_l = H) + node head.  h_out [n_nodes, H];
 * tmp [n_nodes, H] scratch (needed when L > 1); score_out [n_nodes]. */
int nerrf_oracle_sage_forward(const float* x, int64_t n_nodes, int f_in, const void* rowptr, int rowptr_is64,
                              const int32_t* col, const float* ew, int L, const float* const* W, const float* const* b,
                              int H, const float* node_w, float node_b, float* h_out, float* tmp, float* score_out) {
    if (L < 1) return -1;
    const float* in = x;
    int F = f_in;
    for (int l = 0; l < L; ++l) {
        /* ping-pong so that the last layer lands in h_out */
        float* out = ((L - 1 - l) % 2 == 0) ? h_out : tmp;
        if (!out) return -1;
        int rc = nerrf_oracle_sage_layer(in, n_nodes, F, rowptr, rowptr_is64, col, ew, W[l], b[l], H, 1, 0, n_nodes, out);
        if (rc) return rc;
        in = out;
        F = H;
    }
    if (score_out) nerrf_oracle_sage_node_head(h_out, n_nodes, H, node_w, node_b, score_out);
    return 0;
}
