/* planner oracle in plain C  --  TEST INFRASTRUCTURE (see oracle/__init__.py), second, independent restatement of
 * oracle/rewards_ref.py + oracle/mcts_ref.py (frozen spec: DESIGN.md 1.3; reference prose: README.md:115,
 * docs/content/docs/architecture.mdx:62-72, threat-model.mdx:205-223).
 *
 * Purpose: (1) pin the numpy oracle bit for bit (tests/test_oracle_c.py), (2) serve as the all-host-cores CPU
 * baseline for the MCTS half of the metric (bench.py): rollouts of one iteration run under OpenMP, the result is
 * independent of the thread count (every rollout writes its own val[r]; sums have a fixed order).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC   (no FMA contraction: every fp32 op is
 * individually rounded, like numpy float32 and the CUDA __f*_rn intrinsics).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        if (r != 9) { k0 += W0; k1 += W1; }
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void nerrf_oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

static int nw_for(int A) { return A <= 1024 ? 1 : (A <= 2048 ? 2 : 4); }

typedef struct {
    int A, NW, chunk, A_pad, n_words;
    float *u, *v, *c;   /* [A_pad] */
    float* vw;          /* [A_pad] spec v1: term of an applied action whose guard (kill action, index < 32) is NOT applied */
    int* g;             /* [A_pad] guard index or -1 */
} Terms;

static void terms_init(Terms* t, const float* p, const float* size, const float* cost, const int32_t* guard, int A) {
    t->A = A; t->NW = nw_for(A); t->chunk = 32 * t->NW; t->A_pad = 1024 * t->NW; t->n_words = 32 * t->NW;
    t->u = (float*)calloc(t->A_pad, sizeof(float)); t->v = (float*)calloc(t->A_pad, sizeof(float));
    t->c = (float*)calloc(t->A_pad, sizeof(float));
    t->vw = (float*)calloc(t->A_pad, sizeof(float)); t->g = (int*)malloc(t->A_pad * sizeof(int));
    for (int a = 0; a < t->A_pad; ++a) t->g[a] = -1;
    for (int a = 0; a < A; ++a) {
        t->u[a] = p[a] * size[a];
        const float om = 1.0f - p[a];
        t->v[a] = om * size[a];
        t->c[a] = cost[a];
        t->vw[a] = t->v[a];
        if (guard && guard[a] >= 0 && guard[a] < 32 && guard[a] < A) {   /* w_a = fl(p_g * u_a); vw_a = fl(v_a + w_a) */
            t->g[a] = guard[a];
            const float w = p[guard[a]] * t->u[a];
            t->vw[a] = t->v[a] + w;
        }
    }
}
static void terms_free(Terms* t) { free(t->u); free(t->v); free(t->c); free(t->vw); free(t->g); }

/* fixed-order score: lane l sums its `chunk` consecutive actions sequentially, xor-butterfly 1,2,4,8,16 */
static float score_state(const Terms* t, const uint32_t* s) {
    float dl[32], dt[32], tl[32], tt[32];
    for (int l = 0; l < 32; ++l) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < t->chunk; ++i) {
            const int act = l * t->chunk + i;
            const int ap = (s[act >> 5] >> (act & 31)) & 1u;
            const int gd = t->g[act];
            const int guard_alive = gd >= 0 && !((s[0] >> gd) & 1u);     /* its kill action is not in the state */
            a = a + (ap ? (guard_alive ? t->vw[act] : t->v[act]) : t->u[act]);
            b = b + (ap ? t->c[act] : 0.f);
        }
        dl[l] = a; dt[l] = b;
    }
    for (int st = 1; st < 32; st <<= 1) {
        for (int l = 0; l < 32; ++l) { tl[l] = dl[l] + dl[l ^ st]; tt[l] = dt[l] + dt[l ^ st]; }
        memcpy(dl, tl, sizeof(dl)); memcpy(dt, tt, sizeof(dt));
    }
    const float tenth = 0.1f * dt[0];
    return -(dl[0] + tenth);
}

void nerrf_oracle_score(const uint32_t* states, int64_t B, const float* p, const float* size, const float* cost,
                        const int32_t* guard, int A, float* out) {
    Terms t; terms_init(&t, p, size, cost, guard, A);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) out[b] = score_state(&t, states + b * t.n_words);
    terms_free(&t);
}

static int popc(uint32_t x) { return __builtin_popcount(x); }
static int n_legal(const uint32_t* s, int nw) { int z = 0; for (int k = 0; k < nw; ++k) z += popc(~s[k]); return z; }
/* index of the j-th zero bit (ascending) */
static int kth_zero(const uint32_t* s, int nw, int j) {
    for (int k = 0; k < nw; ++k) {
        uint32_t z = ~s[k];
        const int pz = popc(z);
        if (j < pz) {
            for (int b = 0; b < 32; ++b) if ((z >> b) & 1u) { if (j == 0) return k * 32 + b; --j; }
        }
        j -= pz;
    }
    return -1;
}

static float tree_sum(float* v, int n) {   /* adjacent pairs, in place */
    while (n > 1) { for (int i = 0; i < n / 2; ++i) v[i] = v[2 * i] + v[2 * i + 1]; n >>= 1; }
    return v[0];
}

/* Leaf-parallel UCT exactly as oracle/mcts_ref.py.  root_n int32 [A_pad], root_w fp32 [A_pad] out. */
int nerrf_oracle_mcts(const float* p, const float* size, const float* cost, const int32_t* guard, int A,
                      const uint32_t* root_state, int R, int D,
                      int T, uint64_t seed, float c, float lo, float inv_range, const float* lnN, int32_t* root_n,
                      float* root_w, int32_t* num_nodes_out) {
    Terms t; terms_init(&t, p, size, cost, guard, A);
    const int nw = t.n_words, Ap = t.A_pad;
    const uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
    int32_t* visits = (int32_t*)calloc(T + 1, sizeof(int32_t));
    int32_t* child_n = (int32_t*)calloc((size_t)(T + 1) * Ap, sizeof(int32_t));
    float* child_w = (float*)calloc((size_t)(T + 1) * Ap, sizeof(float));
    int32_t* child_id = (int32_t*)malloc((size_t)(T + 1) * Ap * sizeof(int32_t));
    memset(child_id, 0xFF, (size_t)(T + 1) * Ap * sizeof(int32_t));
    float* val = (float*)malloc((size_t)R * sizeof(float));
    float* tmp = (float*)malloc((size_t)R * sizeof(float));
    uint32_t* root = (uint32_t*)calloc(nw, sizeof(uint32_t));
    uint32_t* state = (uint32_t*)malloc(nw * sizeof(uint32_t));
    int* path_node = (int*)malloc((D + 2) * sizeof(int)); int* path_act = (int*)malloc((D + 2) * sizeof(int));
    for (int k = 0; k < nw; ++k) root[k] = root_state ? root_state[k] : 0u;
    for (int a = A; a < Ap; ++a) root[a >> 5] |= 1u << (a & 31);
    int num_nodes = 1;

    for (int it = 0; it < T; ++it) {
        int node = 0, depth = 0, plen = 0;
        memcpy(state, root, nw * sizeof(uint32_t));
        while (1) {
            if (visits[node] == 0) break;
            if (depth >= D || n_legal(state, nw) == 0) break;
            const float ln = lnN[visits[node]];
            float best = -INFINITY; int ba = -1;
            for (int a = 0; a < Ap; ++a) {
                if ((state[a >> 5] >> (a & 31)) & 1u) continue;
                const int n = child_n[(size_t)node * Ap + a];
                float key;
                if (n == 0) key = INFINITY;
                else {
                    const float nf = (float)n;
                    const float q = child_w[(size_t)node * Ap + a] / nf;
                    const float ratio = ln / nf;
                    const float ex = c * sqrtf(ratio);
                    key = q + ex;
                }
                if (key > best) { best = key; ba = a; }
            }
            path_node[plen] = node; path_act[plen] = ba; ++plen;
            state[ba >> 5] |= 1u << (ba & 31);
            ++depth;
            const int cid = child_id[(size_t)node * Ap + ba];
            if (cid < 0) { child_id[(size_t)node * Ap + ba] = num_nodes; node = num_nodes; ++num_nodes; break; }
            node = cid;
        }
        const int leaf = node, L0 = n_legal(state, nw);
        const int first_move = (D - depth) > 0 && L0 > 0;
#pragma omp parallel
        {
            uint32_t* s = (uint32_t*)malloc(nw * sizeof(uint32_t));
#pragma omp for schedule(static)
            for (int r = 0; r < R; ++r) {
                memcpy(s, state, nw * sizeof(uint32_t));
                int left = D - depth, L = L0;
                if (first_move) { const int a = kth_zero(s, nw, r % L0); s[a >> 5] |= 1u << (a & 31); --left; --L; }
                uint32_t rnd[4] = {0, 0, 0, 0};
                for (int k = 0; k < left && L > 0; ++k) {
                    if ((k & 3) == 0) philox4x32_10((uint32_t)r, (uint32_t)(k >> 2), (uint32_t)it, 0u, k0, k1, rnd);
                    const int j = (int)(((uint64_t)rnd[k & 3] * (uint64_t)(uint32_t)L) >> 32);
                    const int a = kth_zero(s, nw, j);
                    s[a >> 5] |= 1u << (a & 31);
                    --L;
                }
                const float sc = score_state(&t, s);
                const float d = sc - lo;
                val[r] = d * inv_range;
            }
            free(s);
        }
        memcpy(tmp, val, (size_t)R * sizeof(float));
        const float total = tree_sum(tmp, R);
        if (first_move) {
            const int nq = R < L0 ? R : L0;
            for (int q = 0; q < nq; ++q) {
                const int a = kth_zero(state, nw, q);
                float w = child_w[(size_t)leaf * Ap + a];
                int cnt = 0;
                for (int r = q; r < R; r += L0) { w = w + val[r]; ++cnt; }
                child_w[(size_t)leaf * Ap + a] = w;
                child_n[(size_t)leaf * Ap + a] += cnt;
            }
        }
        visits[leaf] += 1;
        for (int i = 0; i < plen; ++i) {
            const size_t idx = (size_t)path_node[i] * Ap + path_act[i];
            child_n[idx] += R;
            child_w[idx] = child_w[idx] + total;
            visits[path_node[i]] += 1;
        }
    }
    memcpy(root_n, child_n, (size_t)Ap * sizeof(int32_t));
    memcpy(root_w, child_w, (size_t)Ap * sizeof(float));
    *num_nodes_out = num_nodes;
    free(visits); free(child_n); free(child_w); free(child_id); free(val); free(tmp); free(root); free(state);
    free(path_node); free(path_act); terms_free(&t);
    return 0;
}
