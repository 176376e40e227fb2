// planner.mcts.search + planner.rewards.score (DESIGN.md "MCTS spec").
//
// Bit-exact restatement of oracle/mcts_ref.py + oracle/rewards_ref.py: integer visit counts and
// action indices, Philox4x32-10 counter RNG, and fp32 arithmetic written with explicitly rounded
// intrinsics (__fadd_rn / __fmul_rn / __fdiv_rn / __fsqrt_rn: never contracted to FMA) in the
// spec'd association order.
//
// One PERSISTENT cooperative kernel runs all T iterations of the leaf-parallel search:
//   select   every CTA redundantly descends the tree (block-wide UCT argmax per level; tree is
//            L2-resident and read with ld.global.cg) -> no grid barrier between select/rollout
//   rollout  one warp per rollout: lane l holds NW 32-bit words of the action bitset; j-th
//            legal action = warp prefix-sum of popcounts + __fns; reward = per-lane sequential
//            sum over its 32*NW actions (terms staged in smem, transposed -> conflict-free)
//            + xor-butterfly; value written to val[r]
//   grid barrier (monotonic counter: one atomicAdd + ld.acquire spin per CTA; co-residency is guaranteed by
//                 the cooperative launch; cheaper than cg::grid.sync with one fat CTA per SM)
//   backup   CTA 0: adjacent-pairs tree sum of val -> path edges; all CTAs: per-first-action
//            child statistics of the leaf (fixed ascending-r order)
//   grid barrier
#include "common.cuh"

namespace nerrf {

constexpr int MAXD = 256;
constexpr int MCTS_THREADS = 896;     // 28 warps: one CTA per SM covers 4096 rollouts in a single pass (148 x 28)
constexpr int MCTS_WARPS = MCTS_THREADS / 32;
constexpr int MAXR = 8192;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Stage u/v/cost into smem, transposed: action a = lane*chunk + i lives at [i*32 + lane].
template <int NW>
__device__ __forceinline__ void stage_terms(const float* __restrict__ p, const float* __restrict__ size,
                                            const float* __restrict__ cost, int A, float* u_s, float* v_s, float* c_s) {
    constexpr int CHUNK = 32 * NW, A_PAD = 1024 * NW;
    for (int a = threadIdx.x; a < A_PAD; a += blockDim.x) {
        float u = 0.f, v = 0.f, c = 0.f;
        if (a < A) {
            const float pa = p[a], sa = size[a];
            u = __fmul_rn(pa, sa);
            v = __fmul_rn(__fsub_rn(1.0f, pa), sa);
            c = cost[a];
        }
        const int lane = a / CHUNK, i = a % CHUNK;
        u_s[i * 32 + lane] = u; v_s[i * 32 + lane] = v; c_s[i * 32 + lane] = c;
    }
}

// score of the state held across a warp (lane holds words w[0..NW)), spec'd order.  All lanes return it.
template <int NW>
__device__ __forceinline__ float warp_score(const uint32_t (&w)[NW], const float* u_s, const float* v_s, const float* c_s,
                                            int lane) {
    float dl = 0.f, dt = 0.f;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
#pragma unroll 8
        for (int b = 0; b < 32; ++b) {
            const int i = k * 32 + b;
            const bool ap = (w[k] >> b) & 1u;
            dl = __fadd_rn(dl, ap ? v_s[i * 32 + lane] : u_s[i * 32 + lane]);
            dt = __fadd_rn(dt, ap ? c_s[i * 32 + lane] : 0.f);
        }
    }
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        dl = __fadd_rn(dl, __shfl_xor_sync(0xffffffffu, dl, s));
        dt = __fadd_rn(dt, __shfl_xor_sync(0xffffffffu, dt, s));
    }
    return -__fadd_rn(dl, __fmul_rn(0.1f, dt));
}

// apply the j-th legal (zero) action, ascending a, to the warp-distributed state.  zl = this lane's
// zero count, incl = inclusive prefix over lanes; both are UPDATED in place (the chosen lane loses one
// zero, so every prefix from that lane on drops by one): no re-scan for the next step.
// The n-th-set-bit search inside the chosen word is done by the whole warp: the word is broadcast,
// lane b tests "bit b is legal and exactly `rank` legal bits lie below it", one ballot finds b.
template <int NW>
__device__ __forceinline__ void apply_jth(uint32_t (&w)[NW], int& zl, int& incl, int j, int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, incl > j);
    const int src = __ffs(m) - 1;
    int rank = j - (incl - zl);                       // meaningful in lane `src` only
    int kk = 0;
    uint32_t zsel = ~w[0];
#pragma unroll
    for (int k = 1; k < NW; ++k) {                    // pick the word that holds the rank-th legal bit of this lane
        const int pz = __popc(zsel);
        const bool next = (kk == k - 1) && (rank >= pz);
        if (next) { rank -= pz; kk = k; zsel = ~w[k]; }
    }
    const uint32_t z = __shfl_sync(0xffffffffu, zsel, src);
    const int rk = __shfl_sync(0xffffffffu, rank, src);
    const bool hit = ((z >> lane) & 1u) && (__popc(z & ((1u << lane) - 1u)) == rk);
    const int bit = __ffs(__ballot_sync(0xffffffffu, hit)) - 1;
    if (lane == src) {
#pragma unroll
        for (int k = 0; k < NW; ++k)
            if (k == kk) w[k] |= 1u << bit;
        zl -= 1;
    }
    if (lane >= src) incl -= 1;
}

template <int NW>
__device__ __forceinline__ void lane_zero_scan(const uint32_t (&w)[NW], int lane, int& zl, int& incl, int& total) {
    zl = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) zl += __popc(~w[k]);
    incl = zl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
}

// q-th zero bit of a state stored as a plain word array (single thread).
__device__ __forceinline__ int kth_zero_serial(const uint32_t* s, int n_words, int q) {
    for (int k = 0; k < n_words; ++k) {
        const uint32_t z = ~s[k];
        const int pz = __popc(z);
        if (q < pz) return k * 32 + (int)__fns(z, 0, q + 1);
        q -= pz;
    }
    return -1;
}

// Grid-wide barrier on a monotonically increasing counter (zeroed by the host before the launch).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();                                   // publish this CTA's writes
        atomicAdd(counter, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v < target) __nanosleep(32);
        } while (v < target);
    }
    __syncthreads();
}

struct MctsArgs {
    const float *p, *size, *cost;
    int A;
    const uint32_t* root_state;
    int R, D, T;
    uint32_t k0, k1;
    float c, lo, inv_range;
    const float* lnN;
    int32_t* root_n;
    float* root_w;
    int32_t* num_nodes_out;
    int32_t* visits;     // [T+1]
    int32_t* child_n;    // [(T+1) * A_pad]
    float* child_w;
    int32_t* child_id;
    float* val;          // [R]
    int32_t* g_num_nodes;
    unsigned* barrier;   // grid barrier counter (zero at launch)
};

template <int NW>
__global__ void __launch_bounds__(MCTS_THREADS) mcts_search_kernel(MctsArgs P) {
    constexpr int A_PAD = 1024 * NW, NWORDS = 32 * NW;
    unsigned bar_target = 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* u_s = reinterpret_cast<float*>(smem_raw);
    float* v_s = u_s + A_PAD;
    float* c_s = v_s + A_PAD;
    __shared__ uint32_t s_state[NWORDS];
    __shared__ int2 s_path[MAXD];
    __shared__ float s_key[MCTS_WARPS];
    __shared__ float s_tree[MCTS_WARPS];
    __shared__ int s_arg[MCTS_WARPS];
    __shared__ int s_node, s_depth, s_plen, s_created, s_stop, s_L0, s_numnodes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_terms<NW>(P.p, P.size, P.cost, P.A, u_s, v_s, c_s);
    __syncthreads();

    for (int t = 0; t < P.T; ++t) {
        // ------------------------------------------------------------------ select (redundant per CTA)
        for (int k = tid; k < NWORDS; k += MCTS_THREADS) {
            uint32_t wv = P.root_state ? P.root_state[k] : 0u;
            // padding bits (a >= A) are always applied
            const int a0 = k * 32;
            if (a0 + 32 > P.A) wv |= (a0 >= P.A) ? 0xffffffffu : (0xffffffffu << (P.A - a0));
            s_state[k] = wv;
        }
        // the legal count shrinks by one per level: L(depth) = L(root) - depth, L(root) computed once per iteration
        if (warp == 0) {
            __syncwarp();
            int z = 0;
            for (int k = lane; k < NWORDS; k += 32) {
                uint32_t wv = P.root_state ? P.root_state[k] : 0u;
                const int a0 = k * 32;
                if (a0 + 32 > P.A) wv |= (a0 >= P.A) ? 0xffffffffu : (0xffffffffu << (P.A - a0));
                z += __popc(~wv);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
            if (lane == 0) {
                s_node = 0; s_depth = 0; s_plen = 0; s_created = 0;
                s_numnodes = __ldcg(P.g_num_nodes);
                s_L0 = z;                                                // legal actions at the root
                s_stop = (__ldcg(P.visits) == 0 || P.D <= 0 || z == 0) ? 1 : 0;
            }
        }
        __syncthreads();
        while (!s_stop) {                                                // s_stop is rewritten only between the two barriers below
            const int node = s_node;
            const float lnN = __ldg(P.lnN + __ldcg(P.visits + node));
            float best_key = -INFINITY;
            int best_a = 0x7fffffff;
            const size_t base = (size_t)node * A_PAD;
            for (int a = tid; a < A_PAD; a += MCTS_THREADS) {
                if ((s_state[a >> 5] >> (a & 31)) & 1u) continue;
                const int n = __ldcg(P.child_n + base + a);
                float key;
                if (n == 0) {
                    key = INFINITY;
                } else {
                    const float nf = (float)n;
                    const float q = __fdiv_rn(__ldcg(P.child_w + base + a), nf);
                    key = __fadd_rn(q, __fmul_rn(P.c, __fsqrt_rn(__fdiv_rn(lnN, nf))));
                }
                if (key > best_key) { best_key = key; best_a = a; }     // ascending a: strict > keeps the lowest
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ok = __shfl_xor_sync(0xffffffffu, best_key, o);
                const int oa = __shfl_xor_sync(0xffffffffu, best_a, o);
                if (ok > best_key || (ok == best_key && oa < best_a)) { best_key = ok; best_a = oa; }
            }
            if (lane == 0) { s_key[warp] = best_key; s_arg[warp] = best_a; }
            __syncthreads();                                             // everyone has read s_stop / s_node / s_state
            if (tid == 0) {
                float bk = s_key[0]; int ba = s_arg[0];
                for (int i = 1; i < MCTS_WARPS; ++i)
                    if (s_key[i] > bk || (s_key[i] == bk && s_arg[i] < ba)) { bk = s_key[i]; ba = s_arg[i]; }
                s_path[s_plen] = make_int2(node, ba);
                s_plen += 1;
                s_state[ba >> 5] |= 1u << (ba & 31);
                const int depth_new = s_depth + 1;
                s_depth = depth_new;
                const int cid = __ldcg(P.child_id + base + ba);
                if (cid < 0) {
                    s_created = 1; s_node = s_numnodes; s_stop = 1;      // new leaf
                } else {
                    s_node = cid;                                        // descend; stop at a terminal node
                    if (__ldcg(P.visits + cid) == 0 || depth_new >= P.D || s_L0 - depth_new == 0) s_stop = 1;
                }
            }
            __syncthreads();
        }
        if (tid == 0) s_L0 = s_L0 - s_depth;                             // legal count of the leaf state
        __syncthreads();
        const int leaf = s_node, depth = s_depth, L0 = s_L0;
        const bool first_move = (P.D - depth) > 0 && L0 > 0;

        // ------------------------------------------------------------------ rollouts (warp per rollout)
        for (int r = (int)blockIdx.x * MCTS_WARPS + warp; r < P.R; r += (int)gridDim.x * MCTS_WARPS) {
            uint32_t w[NW];
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = s_state[lane * NW + k];
            int left = P.D - depth;
            int zl, incl, L;
            lane_zero_scan<NW>(w, lane, zl, incl, L);                     // once; maintained incrementally below
            if (first_move) {
                apply_jth<NW>(w, zl, incl, r % L0, lane);
                L -= 1; left -= 1;
            }
            for (int k = 0; k < left && L > 0; k += 4) {                  // one Philox call feeds four steps
                uint32_t rnd[4];
                philox4x32_10((uint32_t)r, (uint32_t)(k >> 2), (uint32_t)t, 0u, P.k0, P.k1, rnd);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (k + q < left && L > 0) {
                        const int j = (int)__umulhi(rnd[q], (uint32_t)L);
                        apply_jth<NW>(w, zl, incl, j, lane);
                        L -= 1;
                    }
                }
            }
            const float sc = warp_score<NW>(w, u_s, v_s, c_s, lane);
            if (lane == 0) P.val[r] = __fmul_rn(__fsub_rn(sc, P.lo), P.inv_range);
        }
        grid_barrier(P.barrier, bar_target);

        // ------------------------------------------------------------------ backup
        if (first_move) {   // leaf children: rank q -> action, fixed ascending-r accumulation
            const int nq = P.R < L0 ? P.R : L0;
            const size_t lbase = (size_t)leaf * A_PAD;
            for (int q = (int)blockIdx.x * MCTS_THREADS + tid; q < nq; q += (int)gridDim.x * MCTS_THREADS) {
                const int a = kth_zero_serial(s_state, NWORDS, q);
                float wsum = __ldcg(P.child_w + lbase + a);
                int cnt = 0;
                for (int r = q; r < P.R; r += L0) { wsum = __fadd_rn(wsum, __ldcg(P.val + r)); ++cnt; }
                P.child_w[lbase + a] = wsum;
                P.child_n[lbase + a] = __ldcg(P.child_n + lbase + a) + cnt;
            }
        }
        if (blockIdx.x == 0) {
            // adjacent-pairs tree sum of val[0..R): each thread reduces an aligned block of m values in
            // registers / local memory, lanes combine with the xor butterfly (== adjacent pairs, lane i holds
            // block i), warps with a fixed pairing.  Same tree as the oracle for any power-of-two R.
            constexpr int TS = 512;                                     // threads used by the tree sum (power of two)
            const int m = P.R >= TS ? P.R / TS : 1;
            const int nthr = P.R / m;                                   // power of two <= TS
            float v = 0.f;
            if (tid < nthr) {
                float loc[MAXR / TS];
                for (int i = 0; i < m; ++i) loc[i] = __ldcg(P.val + tid * m + i);
                for (int st = 1; st < m; st <<= 1)
                    for (int i = 0; i < m; i += 2 * st) loc[i] = __fadd_rn(loc[i], loc[i + st]);
                v = loc[0];
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float other = __shfl_xor_sync(0xffffffffu, v, o);
                if (o < nthr) v = __fadd_rn(v, other);
            }
            if (lane == 0) s_tree[warp] = v;
            __syncthreads();
            float total = s_tree[0];
            if (nthr > 32) {                                            // second level: warp sums, same adjacent pairing
                const int nw = nthr / 32;                               // 2, 4, 8 or 16
                float u = (warp == 0 && lane < nw) ? s_tree[lane] : 0.f;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const float other = __shfl_xor_sync(0xffffffffu, u, o);
                    if (o < nw) u = __fadd_rn(u, other);
                }
                if (tid == 0) s_tree[0] = u;
                __syncthreads();
                total = s_tree[0];
            }
            const int plen = s_plen;
            for (int i = tid; i < plen; i += MCTS_THREADS) {
                const int2 e = s_path[i];
                const size_t idx = (size_t)e.x * A_PAD + e.y;
                P.child_n[idx] = __ldcg(P.child_n + idx) + P.R;
                P.child_w[idx] = __fadd_rn(__ldcg(P.child_w + idx), total);
                P.visits[e.x] = __ldcg(P.visits + e.x) + 1;
            }
            if (tid == 0) {
                P.visits[leaf] = __ldcg(P.visits + leaf) + 1;
                if (s_created) {
                    const int2 e = s_path[plen - 1];
                    P.child_id[(size_t)e.x * A_PAD + e.y] = leaf;
                    *P.g_num_nodes = s_numnodes + 1;
                }
            }
        }
        grid_barrier(P.barrier, bar_target);
    }
    // ---------------------------------------------------------------------- outputs
    if (blockIdx.x == 0) {
        for (int a = tid; a < A_PAD; a += MCTS_THREADS) {
            P.root_n[a] = __ldcg(P.child_n + a);
            P.root_w[a] = __ldcg(P.child_w + a);
        }
        if (tid == 0) *P.num_nodes_out = __ldcg(P.g_num_nodes);
    }
}

// ------------------------------------------------------------------------------------------ rewards.score
template <int NW>
__global__ void __launch_bounds__(256) reward_score_kernel(const uint32_t* __restrict__ states, int64_t B,
                                                           const float* __restrict__ p, const float* __restrict__ size,
                                                           const float* __restrict__ cost, int A, float* __restrict__ out) {
    constexpr int A_PAD = 1024 * NW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* u_s = reinterpret_cast<float*>(smem_raw);
    float* v_s = u_s + A_PAD;
    float* c_s = v_s + A_PAD;
    stage_terms<NW>(p, size, cost, A, u_s, v_s, c_s);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t s = (int64_t)blockIdx.x * 8 + warp; s < B; s += (int64_t)gridDim.x * 8) {
        uint32_t w[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) w[k] = states[s * (32 * NW) + lane * NW + k];
        const float sc = warp_score<NW>(w, u_s, v_s, c_s, lane);
        if (lane == 0) out[s] = sc;
    }
}

static int nw_for(int A) { return A <= 1024 ? 1 : (A <= 2048 ? 2 : 4); }

struct MctsLayout {
    size_t visits, child_n, child_w, child_id, val, numnodes, barrier, total;
};
static MctsLayout mcts_layout(int A, int T, int R) {
    const size_t A_pad = 1024 * (size_t)nw_for(A);
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    MctsLayout L;
    size_t o = 0;
    L.visits = o; o = al(o + (size_t)(T + 1) * 4);
    L.child_n = o; o = al(o + (size_t)(T + 1) * A_pad * 4);
    L.child_w = o; o = al(o + (size_t)(T + 1) * A_pad * 4);
    L.numnodes = o; o = al(o + 4);
    L.barrier = o; o = al(o + 4);
    L.child_id = o; o = al(o + (size_t)(T + 1) * A_pad * 4);
    L.val = o; o = al(o + (size_t)R * 4);
    L.total = o;
    return L;
}

template <int NW>
static int launch_mcts(MctsArgs& args, cudaStream_t st) {
    const size_t smem = (size_t)3 * 1024 * NW * 4;
    NERRF_CHECK_CUDA(cudaFuncSetAttribute(mcts_search_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NERRF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mcts_search_kernel<NW>, MCTS_THREADS, smem));
    NERRF_REQUIRE(per_sm >= 1, "mcts kernel does not fit on an SM");
    const int max_blocks = per_sm * sm_count();
    int want = (args.R + MCTS_WARPS - 1) / MCTS_WARPS;
    if (want < 1) want = 1;
    const int grid = want < max_blocks ? want : max_blocks;
    void* kargs[] = {(void*)&args};
    NERRF_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)mcts_search_kernel<NW>, dim3(grid), dim3(MCTS_THREADS), kargs, smem, st));
    return NERRF_OK;
}

}  // namespace nerrf

using namespace nerrf;

extern "C" int nerrf_reward_score(const uint32_t* states, int64_t B, const float* p, const float* size,
                                  const float* cost, int A, float* out, nerrf_stream_t stream) {
    NERRF_REQUIRE(states && p && size && cost && out, "null pointer");
    NERRF_REQUIRE(A >= 1 && A <= 4096, "number of actions must be in 1..4096 (got %d)", A);
    NERRF_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return NERRF_OK;
    const int NW = nw_for(A);
    const size_t smem = (size_t)3 * 1024 * NW * 4;
    int64_t g = (B + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (NW == 1) reward_score_kernel<1><<<(unsigned)g, 256, smem, st>>>(states, B, p, size, cost, A, out);
    else if (NW == 2) reward_score_kernel<2><<<(unsigned)g, 256, smem, st>>>(states, B, p, size, cost, A, out);
    else {
        NERRF_CHECK_CUDA(cudaFuncSetAttribute(reward_score_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        reward_score_kernel<4><<<(unsigned)g, 256, smem, st>>>(states, B, p, size, cost, A, out);
    }
    return launch_status("reward_score_kernel");
}

extern "C" int nerrf_mcts_workspace_bytes(int A, int T, int R, size_t* bytes) {
    NERRF_REQUIRE(bytes, "null out");
    NERRF_REQUIRE(A >= 1 && A <= 4096, "number of actions must be in 1..4096 (got %d)", A);
    NERRF_REQUIRE(T >= 1 && T <= (1 << 20), "iterations out of range");
    NERRF_REQUIRE(R >= 1 && R <= MAXR && (R & (R - 1)) == 0, "R must be a power of two in 1..%d (got %d)", MAXR, R);
    *bytes = mcts_layout(A, T, R).total;
    return NERRF_OK;
}

extern "C" int nerrf_mcts_search(const float* p, const float* size, const float* cost, int A, const uint32_t* root_state,
                                 int R, int D, int T, uint64_t seed, float c, float lo, float inv_range,
                                 const float* ln_table, int32_t* root_n, float* root_w, int32_t* num_nodes,
                                 void* workspace, size_t workspace_bytes, nerrf_stream_t stream) {
    size_t need = 0;
    int rc = nerrf_mcts_workspace_bytes(A, T, R, &need);
    if (rc) return rc;
    NERRF_REQUIRE(p && size && cost && ln_table && root_n && root_w && num_nodes && workspace, "null pointer");
    NERRF_REQUIRE(D >= 0 && D <= MAXD, "depth must be in 0..%d (got %d)", MAXD, D);
    if (workspace_bytes < need) {
        set_error("mcts workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return NERRF_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const MctsLayout L = mcts_layout(A, T, R);
    unsigned char* ws = (unsigned char*)workspace;
    // visits, child_n, child_w, num_nodes are contiguous: zero them; child_id = -1
    NERRF_CHECK_CUDA(cudaMemsetAsync(ws + L.visits, 0, L.child_id - L.visits, st));
    NERRF_CHECK_CUDA(cudaMemsetAsync(ws + L.child_id, 0xFF, L.val - L.child_id, st));
    NERRF_CHECK_CUDA(cudaMemsetAsync(ws + L.numnodes, 1, 1, st));      // little-endian int32 1 (the word was zeroed above): root node
    MctsArgs a;
    a.p = p; a.size = size; a.cost = cost; a.A = A; a.root_state = root_state; a.R = R; a.D = D; a.T = T;
    a.k0 = (uint32_t)(seed & 0xffffffffu); a.k1 = (uint32_t)(seed >> 32);
    a.c = c; a.lo = lo; a.inv_range = inv_range; a.lnN = ln_table;
    a.root_n = root_n; a.root_w = root_w; a.num_nodes_out = num_nodes;
    a.visits = (int32_t*)(ws + L.visits); a.child_n = (int32_t*)(ws + L.child_n); a.child_w = (float*)(ws + L.child_w);
    a.child_id = (int32_t*)(ws + L.child_id); a.val = (float*)(ws + L.val); a.g_num_nodes = (int32_t*)(ws + L.numnodes); a.barrier = (unsigned*)(ws + L.barrier);
    const int NW = nw_for(A);
    if (NW == 1) return launch_mcts<1>(a, st);
    if (NW == 2) return launch_mcts<2>(a, st);
    return launch_mcts<4>(a, st);
}

extern "C" int nerrf_mcts_search_host(const float* p, const float* size, const float* cost, int A,
                                      const uint32_t* root_state_host, int R, int D, int T, uint64_t seed, float c,
                                      float lo, float inv_range, const float* ln_table_host, int32_t* root_n_host,
                                      float* root_w_host, int32_t* num_nodes_host) {
    size_t need = 0;
    int rc = nerrf_mcts_workspace_bytes(A, T, R, &need);
    if (rc) return rc;
    NERRF_REQUIRE(p && size && cost && ln_table_host && root_n_host && root_w_host && num_nodes_host, "null pointer");
    const int NW = nw_for(A);
    const size_t A_pad = 1024 * (size_t)NW, nwords = 32 * (size_t)NW;
    unsigned char* d = nullptr;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_p = 0, o_s = al(o_p + A * 4), o_c = al(o_s + A * 4), o_rs = al(o_c + A * 4), o_ln = al(o_rs + nwords * 4);
    const size_t o_rn = al(o_ln + (size_t)(T + 2) * 4), o_rw = al(o_rn + A_pad * 4), o_nn = al(o_rw + A_pad * 4);
    const size_t o_ws = al(o_nn + 4), total = o_ws + need;
    NERRF_CHECK_CUDA(cudaMalloc((void**)&d, total));
    cudaStream_t st = 0;
    rc = NERRF_OK;
    auto fail = [&](cudaError_t e, const char* what) { set_error("%s failed: %s", what, cudaGetErrorString(e)); rc = NERRF_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d + o_p, p, A * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) fail(e, "H2D p");
    if (!rc && (e = cudaMemcpyAsync(d + o_s, size, A * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) fail(e, "H2D size");
    if (!rc && (e = cudaMemcpyAsync(d + o_c, cost, A * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) fail(e, "H2D cost");
    if (!rc && root_state_host && (e = cudaMemcpyAsync(d + o_rs, root_state_host, nwords * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) fail(e, "H2D root");
    if (!rc && (e = cudaMemcpyAsync(d + o_ln, ln_table_host, (size_t)(T + 2) * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) fail(e, "H2D ln");
    if (!rc)
        rc = nerrf_mcts_search((const float*)(d + o_p), (const float*)(d + o_s), (const float*)(d + o_c), A,
                               root_state_host ? (const uint32_t*)(d + o_rs) : nullptr, R, D, T, seed, c, lo, inv_range,
                               (const float*)(d + o_ln), (int32_t*)(d + o_rn), (float*)(d + o_rw), (int32_t*)(d + o_nn),
                               d + o_ws, need, st);
    if (!rc && (e = cudaMemcpyAsync(root_n_host, d + o_rn, A_pad * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) fail(e, "D2H root_n");
    if (!rc && (e = cudaMemcpyAsync(root_w_host, d + o_rw, A_pad * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) fail(e, "D2H root_w");
    if (!rc && (e = cudaMemcpyAsync(num_nodes_host, d + o_nn, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) fail(e, "D2H num_nodes");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess && !rc) fail(e, "sync");
    cudaFree(d);
    return rc;
}
