// Fused GraphSAGE-T layer on 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
//   out[v, :] = act( [x_v || m_v] @ W + b ),   m_v = weighted mean of gathered source rows
//
// computed TRANSPOSED so that the small, reused operand lives in tensor memory:
//
//   D^T[H=128 features, rows] = W^T[128, K=2F] . X^T[K, rows]
//     A operand = W^T, resident in TMEM for the whole (persistent) kernel, bf16 hi and lo halves
//                 (lane = feature, two K elements per 32-bit column)
//     B operand = the gathered tile X[rows, K] exactly as the gather warps produce it: row-major,
//                 K contiguous == the UMMA "K-major" canonical layout (128-byte swizzle), bf16 hi/lo
//     D         = fp32 accumulator in TMEM, [128 lanes x 64 columns] per tile, double buffered
//
// fp32-level accuracy from bf16 tensor cores (the north star's 1e-4 bound rules out plain
// TF32/bf16): every fp32 value v is split v = hi + lo (hi = bf16(v), lo = bf16(v - hi)) and
// three MMAs accumulate x_hi.w_hi + x_lo.w_hi + x_hi.w_lo in fp32 (error ~1e-5 relative).
//
// The aggregate never round-trips HBM: gather warps write hi/lo bf16 straight into the shared
// memory B operand.  Warp roles in one CTA per SM (512 threads):
//   warps 0-3   epilogue: tcgen05.ld accumulator (lane = feature), + bias, ReLU, 128 B/warp stores
//   warp  4     TMEM allocator + single-thread MMA issuer (tcgen05.mma, commit -> mbarriers)
//   warp  5     edge-block loader: stages each tile's rowptr words and its col / ew slices into shared memory --
//               the slices as two BULK async copies (cp.async.bulk, byte count on the stage's mbarrier) --
//               running up to IDX_STAGES tiles ahead; the next tile's rowptr words are prefetched before it
//               waits for a free stage
//   warps 6-..  gather: one CSR row per warp at a time, rows round-robin; every source row (512 B at F=128,
//               128 B at F=32) is ONE bulk async copy into the warp's shared-memory ring, completion counted in
//               bytes on the ring slot's mbarrier (the "TMA-staged edge block" of the north star); indices come
//               from shared memory, so the only global latency a row exposes is its x rows
// Pipelines: idx stages full/empty (loader <-> gather), smem operand stages full/empty (gather <-> MMA),
// TMEM accumulators full/empty (MMA <-> epilogue).
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "sage_gather.cuh"

namespace nerrf {

namespace {

constexpr int TN = 32;                 // destination rows per tile (UMMA N)
constexpr int UM = 128;                // UMMA M = hidden width
constexpr int EPI_WARPS = 4;
constexpr int MMA_WARP = 4;
constexpr int LOADER_WARP = 5;
constexpr int GATHER_WARP0 = 6;
constexpr int EMAX = 512;              // staged edges per tile (4 KB); the rest of a heavier tile is read from global
constexpr int EPAD = EMAX + 4;         // a slice is copied as its 16-byte aligned superset: up to 3 leading + 3 trailing elements
constexpr int IDX_STAGE_BYTES = EPAD * 8 + 512;      // col[EPAD] | ew[EPAD] | rp[TN+1] (relative) | e_lo (int64) | col_off, ew_off
// items (source rows) per sub-batch (= one cp.async group): 4 warp-wide copy instructions, i.e. 4 * (32 / (F/4))
template <int F> constexpr int qs_for() { return 4 * (32 / (F / 4)); }
constexpr int NQ = 4;                  // sub-batch slots in a warp's ring; NQ-1 groups are in flight

// NS = number of bf16 terms each fp32 value is split into (v = p0 + p1 [+ p2], p_i = bf16 of the
// running residual).  NS = 3 with the six products (x0w0, x1w0, x0w1, x1w1, x2w0, x0w2) drops only
// terms below 2^-24: fp32-equivalent.  NS = 2 with three products is ~1e-5 relative.
template <int F, int NS>
struct UmmaCfg {
#ifdef NERRF_EXP_GW128                                            // experiment knob (scripts/exp_gather_warps.sh): fewer gather warps at F=128
    static constexpr int GATHER_WARPS = (F == 32) ? 18 : (F == 128 ? NERRF_EXP_GW128 : 14);
#else
    static constexpr int GATHER_WARPS = (F == 32) ? 18 : 14;  // F=32 leaves room for more rings (and needs more TLP)
#endif
    static constexpr int THREADS = (GATHER_WARP0 + GATHER_WARPS) * 32;
    // F = 32 (quad-mode gather): a warp runs up to 3 units = ~7 tiles ahead of what it has consumed and holds a tile's edge
    // block from issue to completion, so the look-ahead is bounded by the number of edge-block stages, not by the rings
    // (ncu r2g: with 4 stages 77 % of the issue-side polls found the next tile not staged)
    static constexpr int IDX_STAGES = (F == 32) ? 9 : 3;
    static constexpr int K = 2 * F;
    static constexpr int KB = K / 64;                        // 128-byte K blocks
    static constexpr int KSTEPS = K / 16;                    // MMAs (K=16) per product
    static constexpr int PART_BYTES = KB * TN * 128;         // one bf16 term of one stage
    static constexpr int STAGE_BYTES = NS * PART_BYTES;
    static constexpr int QS = qs_for<F>();                    // 4 (F=128), 8 (F=64), 16 (F=32): always 2 KB per sub-batch
    static constexpr int RING_BYTES = NQ * QS * F * 4;       // per gather warp: NQ sub-batches x QS source rows
    static constexpr int STAGES = (F == 128) ? 2 : (F == 32 ? 3 : 4);
    static constexpr int W_PART_COLS = K / 2;                // W^T term p lives in TMEM columns [p*K/2, (p+1)*K/2)
    static constexpr int ACC_COL0 = NS * W_PART_COLS;
    static constexpr int TMEM_COLS = (ACC_COL0 + 2 * TN <= 256) ? 256 : 512;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + IDX_STAGES * IDX_STAGE_BYTES +
                                   (size_t)GATHER_WARPS * RING_BYTES + 1024 /*head reduce*/ + 1024 /*align*/ + 1024 /*barriers*/;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static_assert(STAGES >= 2, "need at least two smem stages");
    static_assert(ACC_COL0 + 2 * TN <= 512, "TMEM overflow");
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// arrive (count 1) and add `bytes` to the phase's pending transaction count: the phase completes when the
// arrival count AND the byte count are both satisfied
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// bulk async copy global -> shared (the non-tensor TMA path, SASS UBLKCP): `bytes` (multiple of 16) from a 16-byte
// aligned global address to a 16-byte aligned shared address; completion is signalled on `bar` as complete_tx(bytes)
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// for the roles that have slack (epilogue, MMA issuer, loader): poll with a sleep so they do not burn
// the issue slots the gather warps need (a tile period is ~10 us)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t done;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(400);
    }
}
// predicated 16-byte async copy (no branch)
__device__ __forceinline__ void cp_async16_pred(uint32_t smem_dst, const void* gsrc, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
        ::"r"(smem_dst), "l"(gsrc), "r"((uint32_t)pred)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]^T   (kind::f16: bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), version 1.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                               // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                     // SBO = 1024 B
    d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}
// Instruction descriptor: D=f32, A=B=bf16, both K-major, N = TN, M = 128.
__host__ __device__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
}

// split fp32 values into NS bf16 terms of the running residual; pairs packed (first element low)
template <int NS>
__device__ __forceinline__ void split_pair(float a, float b, uint32_t (&parts)[NS]) {
#pragma unroll
    for (int p = 0; p < NS; ++p) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        parts[p] = *reinterpret_cast<const uint32_t*>(&h);
        if (p + 1 < NS) {
            const float2 f = __bfloat1622float2(h);
            a -= f.x; b -= f.y;                        // exact: the residual is representable in fp32
        }
    }
}
template <int NS>
__device__ __forceinline__ void split4(const float4 v, uint2 (&parts)[NS]) {
    uint32_t lo[NS], hi[NS];
    split_pair<NS>(v.x, v.y, lo);
    split_pair<NS>(v.z, v.w, hi);
#pragma unroll
    for (int p = 0; p < NS; ++p) parts[p] = make_uint2(lo[p], hi[p]);
}

// v[j] (j = 0..31) per lane  ->  sum over the 32 lanes of v[lane]: a 32x32 transpose-reduce in 31 shuffles
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int n = 16 >> s;                                  // also the xor offset
        const bool up = (lane & n) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = up ? v[i] : v[i + n];
            const float keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return v[0];
}

// byte offset of element (row r, k) inside one hi/lo part of a stage (k multiple of 4)
__device__ __forceinline__ uint32_t b_offset(int r, int k) {
    const int kb = k >> 6, col = k & 63;
    const int chunk = col >> 3;                            // 16-byte chunk in the 128-byte row
    return (uint32_t)(kb * (TN * 128) + r * 128 + (((chunk ^ (r & 7)) << 4) | ((col & 7) << 1)));
}

// ---------------------------------------------------------------- fused exchange (multi-GPU)
// Peer-mapped [n_nodes, 128] embedding buffers of the other ranks (NVLink P2P, CUDA IPC / symmetric memory).
// The epilogue stores every output row slice to the local buffer AND to each peer buffer, so the per-layer
// embedding exchange of the 1-D sharded forward happens inside the layer kernel, overlapped tile by tile.
constexpr int MAX_PEERS = 7;
struct Peers {
    float* out[MAX_PEERS];
    int n;
    const uint8_t* need;  // optional [n_nodes]: bit pr set <=> peer slot pr references this row as a source; rows a
                          // peer never reads are not sent to it (nullptr: send every row to every peer)
    float* mc;            // NVSwitch multicast address of the same buffer on ALL ranks (incl. this one), or nullptr:
                          // one multimem.st per element, replicated by the switch (egress 1x instead of (G-1)x)
};
__device__ __forceinline__ void multimem_st4(float* addr, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// 4x4 transpose across the 4 lanes of a quad: in: lane i holds a[0..3]; out: lane i holds (a[i] of lanes 0,1,2,3).
__device__ __forceinline__ void quad_transpose(float& a0, float& a1, float& a2, float& a3, int lane) {
    const bool b0 = lane & 1, b1 = lane & 2;
    float t;
    t = b0 ? a0 : a1; t = __shfl_xor_sync(0xffffffffu, t, 1); if (b0) a0 = t; else a1 = t;
    t = b0 ? a2 : a3; t = __shfl_xor_sync(0xffffffffu, t, 1); if (b0) a2 = t; else a3 = t;
    t = b1 ? a0 : a2; t = __shfl_xor_sync(0xffffffffu, t, 2); if (b1) a0 = t; else a2 = t;
    t = b1 ? a1 : a3; t = __shfl_xor_sync(0xffffffffu, t, 2); if (b1) a1 = t; else a3 = t;
}

// ---------------------------------------------------------------- long rows (hub destinations)
// A destination row with more than LONG_T in-edges would be gathered by ONE warp.  A deterministic pre-pass
// cuts every such row into fixed chunks of LONG_CH edges; one CTA per chunk computes the unnormalised
// weighted sum + weight sum (fixed order: 8 warps x 64 edges, then warp 0 adds the 8 warp partials in order)
// into a side buffer.  The main kernel then sees the row as the items [self, partial_0 .. partial_{nc-1}]
// (weight 1 each, weight-sum from pw[]).  The workspace is capacity driven: rows that do not fit stay inline.
// Threshold: a 300-edge row gathered by one warp (512 B per edge at ~6 GB/s per warp) takes ~25 us while the other 31 rows
// of its tile are done in 2 -- and a trace graph has exactly such a row every 64 nodes (the process node of each
// component, in-degree = events of its files).  Measured on the trace-structured bench graph: 2.07 ms per F=128 layer with
// the r1 threshold of 512, see profiles/r02_bench_n2.json vs r02 final.
constexpr int LONG_T = 128;
constexpr int LONG_CH = 256;

struct LongWs {
    int* hdr;              // [0] = number of chunk items claimed (may exceed cap), [1] = cap
    int* hash_key;         // row + 1, 0 = empty
    int* hash_val;         // first item of the row
    int2* queue;           // (row - row_begin? no: absolute row as two ints) -> see pack/unpack
    float* partial;        // [cap][128]
    float* pw;             // [cap]
    int cap;
    int hash_mask;         // slots - 1 (slots = power of two >= 2 * cap)
};

__host__ __device__ inline size_t long_ws_bytes_for(int cap, int slots) {
    return 256 + (size_t)slots * 8 + (size_t)cap * (8 + 128 * 4 + 4);
}
static LongWs long_ws_carve(void* ws, size_t bytes) {
    LongWs L{};
    if (!ws || bytes < 8192) return L;
    int cap = (int)((bytes - 256) / (8 + 128 * 4 + 4 + 32));        // 32 B/item covers the hash (>= 2 slots x 8 B, rounded up)
    cap &= ~3;                                                       // keeps every sub-array 16-byte aligned
    if (cap < 4) return L;
    int slots = 1;
    while (slots < 2 * cap) slots <<= 1;
    while (long_ws_bytes_for(cap, slots) > bytes && cap > 4) { cap = (cap * 3 / 4) & ~3; slots = 1; while (slots < 2 * cap) slots <<= 1; }
    if (long_ws_bytes_for(cap, slots) > bytes) return L;
    unsigned char* b = (unsigned char*)ws;
    L.hdr = (int*)b; b += 256;
    L.hash_key = (int*)b; b += (size_t)slots * 4;
    L.hash_val = (int*)b; b += (size_t)slots * 4;
    L.queue = (int2*)b; b += (size_t)cap * 8;
    L.partial = (float*)b; b += (size_t)cap * 128 * 4;
    L.pw = (float*)b;
    L.cap = cap; L.hash_mask = slots - 1;
    return L;
}
__device__ __forceinline__ int long_hash(int64_t row, int mask) { return (int)(((uint64_t)row * 0x9E3779B97F4A7C15ull) >> 40) & mask; }
__device__ __forceinline__ int long_lookup(const LongWs& L, int64_t row) {      // first item of `row`, or -1
    int h = long_hash(row, L.hash_mask);
    const int key = (int)row + 1;
    for (int probe = 0; probe <= L.hash_mask; ++probe) {
        const int k = L.hash_key[h];
        if (k == key) return L.hash_val[h];
        if (k == 0) return -1;
        h = (h + 1) & L.hash_mask;
    }
    return -1;
}

template <typename RP>
__global__ void __launch_bounds__(256) long_scan_kernel(const RP* __restrict__ rowptr, int64_t row_begin, int64_t row_end, LongWs L) {
    for (int64_t row = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < row_end; row += (int64_t)gridDim.x * blockDim.x) {
        const int64_t deg = (int64_t)rowptr[row + 1] - (int64_t)rowptr[row];
        if (deg <= LONG_T) continue;
        const int nch = (int)((deg + LONG_CH - 1) / LONG_CH);
        const int first = atomicAdd(L.hdr, nch);
        if (first + nch > L.cap) continue;                               // does not fit: the row stays inline
        for (int c = 0; c < nch; ++c) L.queue[first + c] = make_int2((int)row, c);
        int h = long_hash(row, L.hash_mask);
        const int key = (int)row + 1;
        while (atomicCAS(L.hash_key + h, 0, key) != 0) h = (h + 1) & L.hash_mask;
        L.hash_val[h] = first;
    }
}

template <int F, typename RP>
__global__ void __launch_bounds__(256) long_chunk_kernel(const float* __restrict__ x, const RP* __restrict__ rowptr,
                                                         const int32_t* __restrict__ col, const float* __restrict__ ew, LongWs L) {
    __shared__ float4 s_acc[8][32];
    __shared__ float s_w[8];
    constexpr int LPR = F / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int n_items = L.hdr[0];
    if (n_items > L.cap) n_items = L.cap;                                // items past cap were never queued ...
    for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
        const int2 it = L.queue[i];
        if (it.x < 0) continue;                                          // claimed by a row that did not fit (never queued)
        const int64_t e0 = (int64_t)rowptr[it.x], e1 = (int64_t)rowptr[it.x + 1];
        const int64_t c0 = e0 + (int64_t)it.y * LONG_CH;
        const int64_t c1 = (c0 + LONG_CH < e1) ? c0 + LONG_CH : e1;
        const int64_t a = c0 + warp * (LONG_CH / 8);
        int64_t b = a + LONG_CH / 8;
        if (b > c1) b = c1;
        float wsum = 0.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a < b) acc = gather_row<F>(x, col, ew, a, b, lane, &wsum);
        if (lane < LPR) s_acc[warp][lane] = acc;
        if (lane == 0) s_w[warp] = wsum;
        __syncthreads();
        if (warp == 0) {
            if (lane < LPR) {
                float4 t = s_acc[0][lane];
#pragma unroll
                for (int k = 1; k < 8; ++k) { t.x += s_acc[k][lane].x; t.y += s_acc[k][lane].y; t.z += s_acc[k][lane].z; t.w += s_acc[k][lane].w; }
                *reinterpret_cast<float4*>(L.partial + (size_t)i * 128 + 4 * lane) = t;
            }
            if (lane == 0) {
                float t = s_w[0];
#pragma unroll
                for (int k = 1; k < 8; ++k) t += s_w[k];
                L.pw[i] = t;
            }
        }
        __syncthreads();
    }
}

// LONG: compiled with the hub-row (pre-aggregated partial items) paths; the plain variant is used when the
// launch has no hub-row scratch, so graphs without hub rows pay nothing for the feature.
// BULK: source rows travel as one bulk async copy each (cp.async.bulk, SASS UBLKCP) instead of 32 lanes x 16-byte
// cp.async (LDGSTS).  Measured on B200 (profiles/r02_gather_transport.md): the TMA unit retires ~one bulk request per
// 20-30 cycles per SM, so at one request per 128..512-byte row it -- not HBM, not issue slots -- becomes the limiter
// (F=128 layer 1.24 ms vs 1.01 ms).  The LDGSTS transport is therefore the default; the edge-block slices (col / ew,
// kilobytes per request) use bulk copies in both variants.
template <int F, int NS, typename RP, bool LONG, bool BULK>
__global__ void __launch_bounds__((UmmaCfg<F, NS>::THREADS), 1)
sage_layer_umma_kernel(const float* __restrict__ x, const RP* __restrict__ rowptr, const int32_t* __restrict__ col,
                       const float* __restrict__ ew, const float* __restrict__ W, const float* __restrict__ bias,
                       float* __restrict__ out, int64_t row_begin, int64_t row_end, int relu,
                       const float* __restrict__ node_w, float node_b, float* __restrict__ score, const LongWs lw,
                       const Peers peers, unsigned* __restrict__ tile_ctr) {
    using C = UmmaCfg<F, NS>;
    constexpr int K = C::K, LPR = F / 4;
    constexpr int GATHER_WARPS = C::GATHER_WARPS, IDX_STAGES = C::IDX_STAGES;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;          // SWIZZLE_128B needs 1024-B alignment
    unsigned char* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    const uint32_t idx_base = smem_base + C::STAGES * C::STAGE_BYTES;
    const uint32_t bar_base = idx_base + IDX_STAGES * IDX_STAGE_BYTES;     // 256 B of barriers, then the gather rings
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto accf_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto acce_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    auto idxf_bar = [&](int q) { return bar_base + 8u * (2 * C::STAGES + 4 + q); };
    auto idxe_bar = [&](int q) { return bar_base + 8u * (2 * C::STAGES + 4 + IDX_STAGES + q); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4 + 2 * IDX_STAGES);
    // per gather warp: NQ ring-slot barriers (bulk-copy byte counts), private to the warp
    auto ring_bar = [&](int g, int q) { return bar_base + 256u + 8u * (uint32_t)(g * NQ + q); };
    static_assert(256 + GATHER_WARPS * NQ * 8 <= 1024 - 64, "barrier area");
    // DYNAMIC TILE SCHEDULE: the loader warp claims tiles from a global counter (tile_ctr, zero at launch) and publishes
    // the id of the CTA's tl-th tile in this ring (-1 = end of stream) before it releases the tile's edge block; every
    // role reads it after the mbarrier wait that orders it behind that release.  Rows are not equally expensive (hub
    // destinations: a 32-row tile can hold 50x the edges of an average one), a static round-robin left whole SMs idle.
    constexpr int TILE_RING = 16;                  // > IDX_STAGES + STAGES + 2 accumulators: an entry outlives its tile
    volatile int* tile_ids = reinterpret_cast<volatile int*>(smem_gen + (bar_base + 1024 - 64 - smem_base));
    // QUAD mode (F = 32): a gather warp works on FOUR rows at a time, one per 8-lane group (a 128-byte source row is 8 lanes
    // x 16 B), so the per-row overhead -- row setup, the cross-group combine, the bf16 split, the operand stores, the barrier
    // arrivals -- is paid once per four rows and by all 32 lanes (r1/r2 ncu: the F=32 layer was issue-bound at ~500
    // instructions per row, most of them this overhead).  Units of 4 rows are dealt round-robin to the warps; a warp need not
    // own a unit in every tile, so the end of the tile stream is also published in `end_tl` (the closing tile's index).
    constexpr bool QUAD = (F == 32) && !BULK;
    volatile int* end_tl = reinterpret_cast<volatile int*>(smem_gen + (bar_base + 1024 - 64 - 8 - smem_base));
    static_assert(256 + GATHER_WARPS * NQ * 8 <= 1024 - 64 - 8, "barrier area");
    unsigned char* idx_gen = smem_gen + (idx_base - smem_base);
    const uint32_t ring_base = bar_base + 1024;
    float* head_red = reinterpret_cast<float*>(smem_gen + (ring_base - smem_base) + (size_t)GATHER_WARPS * C::RING_BYTES);   // [2][4][32]
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t rows = row_end - row_begin;
    const int64_t n_tiles = (rows + TN - 1) / TN;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), QUAD ? TN / 4 : TN); mbar_init(empty_bar(s), 1); }
            *end_tl = 0x7fffffff;
            for (int a = 0; a < 2; ++a) { mbar_init(accf_bar(a), 1); mbar_init(acce_bar(a), EPI_WARPS * 32); }
            // idx full: ONE arrive.expect_tx by the loader (publishes the rowptr words, counts the bytes of the two bulk
            // copies); empty: one per gather warp.  ring slots: one arrive.expect_tx by the issuing lane per use
            for (int q = 0; q < IDX_STAGES; ++q) { mbar_init(idxf_bar(q), 1); mbar_init(idxe_bar(q), QUAD ? TN / 4 : GATHER_WARPS); }
            for (int g = 0; g < GATHER_WARPS; ++g)
                for (int q = 0; q < NQ; ++q) mbar_init(ring_bar(g, q), 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, C::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // ---- W^T -> TMEM (hi | lo), done once per CTA by the four epilogue warps: thread = feature
    float my_bias = 0.f;
    if (warp < EPI_WARPS) {
        const int f = warp * 32 + lane;
        my_bias = __ldg(bias + f);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < K; k0 += 16) {
            uint32_t parts[NS][8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = __ldg(W + (size_t)(k0 + 2 * i) * UM + f);
                const float b = __ldg(W + (size_t)(k0 + 2 * i + 1) * UM + f);
                uint32_t pp[NS];
                split_pair<NS>(a, b, pp);
#pragma unroll
                for (int p = 0; p < NS; ++p) parts[p][i] = pp[p];
            }
#pragma unroll
            for (int p = 0; p < NS; ++p) tmem_st8(lane_addr + (uint32_t)(p * C::W_PART_COLS + (k0 >> 1)), parts[p]);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= GATHER_WARP0) {
      if constexpr (QUAD) {
        // =========================================================== gather warps, QUAD mode (F = 32): see above
        constexpr int UPT = TN / 4;                                       // units (4 rows) per tile
        constexpr int SLOT_BYTES = 4 * 4 * F * 4;                         // a sub-batch = 4 steps x 4 rows x 128 B = 2 KB
        static_assert(SLOT_BYTES * NQ == C::RING_BYTES, "ring layout");
        const int g = warp - GATHER_WARP0;
        const int grp = lane >> 3, sub = lane & 7;
        // ring slot layout [step][group][8 lanes x 16 B] = lane-linear: conflict-free 16-byte shared loads
        const uint32_t ring_u32 = ring_base + (uint32_t)(g * C::RING_BYTES) + (uint32_t)(lane * 16);
        const float* ring_gen = reinterpret_cast<const float*>(smem_gen + (ring_base - smem_base) + (size_t)g * C::RING_BYTES) + lane * 4;
        const float* xs = x + 4 * sub;
        int stream_end = 0x7fffff00;                                      // in units

        // per-unit state; e0 / items / row / long are per GROUP (each 8-lane group has its own row), nb is warp-uniform
        int iu = g, ib = 0, inb = 1, ie0 = 0, iitems = 0, ilong = -1, itl = -1, itile = 0;  uint32_t irow = 0;  bool ifast = true;
        const int32_t* icol = nullptr; int64_t ielo = 0;
        int cu = g, cb = 0, cnb = 1, ce0 = 0, citems = 0, clong = -1, ctl = -1, ctile = 0;  bool cfast = true;
        const float* cew = nullptr; int64_t celo = 0;

        // Set up unit `iu` for issuing.  NEVER blocks: a warp runs up to NQ-1 sub-batches = up to 3 units = ~7 tiles ahead of
        // what it has consumed, more than there are edge-block stages, so blocking here for a stage that is waiting for this
        // very warp's consumption would deadlock.  false = the unit's tile is not staged yet (the caller emits an empty
        // "bubble" group and retries at its next call) or the stream ended before it (stream_end shrinks to iu).
        auto try_setup_issue_unit = [&]() -> bool {
            const int tl = iu / UPT, r = (iu % UPT) * 4 + grp;
            if (tl != itl) {
                uint32_t done;
                // test_wait, not try_wait: the latter may suspend the warp for a hardware time limit, and this poll sits in front of
                // the consumption of data that has already landed
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(idxf_bar(tl % IDX_STAGES)), "r"((uint32_t)(tl / IDX_STAGES) & 1u) : "memory");
                if (!__all_sync(0xffffffffu, done != 0)) {               // every lane must have acquired the loader's writes
                    if (__any_sync(0xffffffffu, *end_tl < tl)) stream_end = iu;    // the tile will never come
                    return false;
                }
                itl = tl;
                itile = tile_ids[tl % TILE_RING];
                if (itile < 0) stream_end = (tl + 1) * UPT;              // the closing tile: its units carry the end of stream on
            }
            const unsigned char* ibp = idx_gen + (size_t)(tl % IDX_STAGES) * IDX_STAGE_BYTES;
            const int32_t* rp_s = reinterpret_cast<const int32_t*>(ibp + EPAD * 8);
            // (the closing tile's stage holds no rowptr words: its slack reads must not follow a stale offset)
            icol = reinterpret_cast<const int32_t*>(ibp) + (itile >= 0 ? rp_s[96 + 2] : 0);
            ielo = *reinterpret_cast<const int64_t*>(ibp + EPAD * 8 + 384);
            const int64_t row = row_begin + (int64_t)itile * TN + r;
            ib = 0; iitems = 0; ilong = -1; ifast = true;
            if (itile < 0) ie0 = 1;
            if (itile >= 0 && row < row_end) {
                irow = (uint32_t)row;
                ie0 = rp_s[r];
                const int deg = rp_s[r + 1] - ie0;
                iitems = deg + 1;                                        // [self, edges...]
                if (LONG && deg > LONG_T && lw.cap > 0) {
                    ilong = long_lookup(lw, row);
                    if (ilong >= 0) iitems = 1 + (deg + LONG_CH - 1) / LONG_CH;
                }
                ifast = ilong < 0 && ie0 + deg <= EMAX;
            }
            inb = __reduce_max_sync(0xffffffffu, (iitems + 3) >> 2);
            inb = inb > 0 ? inb : 1;                                     // an all-invalid unit still travels as one empty sub-batch
            ifast = __all_sync(0xffffffffu, ifast);
            return true;
        };
        auto setup_consume_unit = [&]() {                                // cu < stream_end, tile already staged
            const int tl = cu / UPT, r = (cu % UPT) * 4 + grp;
            if (tl != ctl) ctile = tile_ids[tl % TILE_RING];
            ctl = tl;
            const unsigned char* ibp = idx_gen + (size_t)(tl % IDX_STAGES) * IDX_STAGE_BYTES;
            const int32_t* rp_s = reinterpret_cast<const int32_t*>(ibp + EPAD * 8);
            cew = reinterpret_cast<const float*>(ibp + EPAD * 4) + (ctile >= 0 ? rp_s[96 + 3] : 0);
            celo = *reinterpret_cast<const int64_t*>(ibp + EPAD * 8 + 384);
            const int64_t row = row_begin + (int64_t)ctile * TN + r;
            cb = 0; citems = 0; clong = -1; cfast = true;
            if (ctile < 0) ce0 = 1;
            if (ctile >= 0 && row < row_end) {
                ce0 = rp_s[r];
                const int deg = rp_s[r + 1] - ce0;
                citems = deg + 1;
                if (LONG && deg > LONG_T && lw.cap > 0) {
                    clong = long_lookup(lw, row);
                    if (clong >= 0) citems = 1 + (deg + LONG_CH - 1) / LONG_CH;
                }
                cfast = clong < 0 && ce0 + deg <= EMAX;
            }
            cnb = __reduce_max_sync(0xffffffffu, (citems + 3) >> 2);
            cnb = cnb > 0 ? cnb : 1;
            cfast = __all_sync(0xffffffffu, cfast);
        };
        // one cp.async group = sub-batch `ib` of the issue unit: 4 steps, step s = item ib*4+s of each of the 4 rows
        bool iready = false, cready = false;
        uint32_t bubbles = 0;                                             // bit q: nothing was issued into ring slot q
        auto issue = [&](int slot) {
            if (iu < stream_end && !iready) iready = try_setup_issue_unit();       // may shrink stream_end
            bubbles |= 1u << slot;
            if (iu < stream_end && iready) {
                bubbles &= ~(1u << slot);
                const int first = ib * 4;
                const uint32_t dst = ring_u32 + (uint32_t)(slot * SLOT_BYTES);
                if (ifast) {                                             // every row of the unit staged in full: branch free
                    uint32_t src[4];
#pragma unroll
                    for (int st = 0; st < 4; ++st) src[st] = (uint32_t)icol[ie0 + first + st - 1];   // slack reads stay in smem
                    if (first == 0) src[0] = irow;                       // item 0 of every row is its self row
#pragma unroll
                    for (int st = 0; st < 4; ++st)
                        cp_async16_pred(dst + (uint32_t)(st * 512), xs + (size_t)src[st] * F, first + st < iitems);
                } else {
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        const int it = first + st;
                        if (it < iitems) {
                            const float* srcp;
                            if (it == 0) srcp = xs + (size_t)irow * F;
                            else if (LONG && ilong >= 0) srcp = lw.partial + (size_t)(ilong + it - 1) * 128 + 4 * sub;
                            else {
                                const int k = ie0 + it - 1;
                                srcp = xs + (size_t)(uint32_t)((k < EMAX) ? icol[k] : __ldg(col + ielo + k)) * F;
                            }
                            cp_async16(dst + (uint32_t)(st * 512), srcp);
                        }
                    }
                }
                if (++ib >= inb) { iu += GATHER_WARPS; iready = false; }
            }
            cp_async_commit();                                            // (possibly empty) group keeps the count in step
        };

        int islot = 0, cslot = 0;
#pragma unroll
        for (int d = 0; d < NQ - 1; ++d) { issue(islot); islot = (islot + 1) % NQ; }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), self = acc;
        float wsum = 0.f;
        int last_tl = -1;
        while (cu < stream_end) {
            cp_async_wait<NQ - 2>();
            __syncwarp();
            issue(islot);
            islot = (islot + 1) % NQ;
            if ((bubbles >> cslot) & 1u) {                                // a bubble: the unit's tile was not staged yet
                cslot = (cslot + 1) % NQ;
                __nanosleep(64);
                continue;
            }
            if (!cready) { setup_consume_unit(); cready = true; }         // first sub-batch of the unit (its tile is staged: the
                                                                          // issue side saw it, and this warp has not released it)
            {
                const int first = cb * 4;
                const float* rs = ring_gen + cslot * (SLOT_BYTES / 4);
                if (cfast) {
                    float4 v[4];
                    float w[4];
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        v[st] = *reinterpret_cast<const float4*>(rs + st * 128);
                        w[st] = cew[ce0 + first + st - 1];
                    }
                    if (first == 0) { self = v[0]; w[0] = 0.f; v[0] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        if (first + st < citems) {
                            wsum += w[st];
                            acc.x = fmaf(w[st], v[st].x, acc.x); acc.y = fmaf(w[st], v[st].y, acc.y);
                            acc.z = fmaf(w[st], v[st].z, acc.z); acc.w = fmaf(w[st], v[st].w, acc.w);
                        }
                    }
                } else {
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        const int it = first + st;
                        if (it < citems) {
                            const float4 v = *reinterpret_cast<const float4*>(rs + st * 128);
                            if (it == 0) {
                                self = v;
                            } else if (LONG && clong >= 0) {             // chunk partial: already weighted
                                wsum += lw.pw[clong + it - 1];
                                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                            } else {
                                const int k = ce0 + it - 1;
                                const float w = (k < EMAX) ? cew[k] : __ldg(ew + celo + k);
                                wsum += w;
                                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                                acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                            }
                        }
                    }
                }
            }
            cslot = (cslot + 1) % NQ;
            if (++cb < cnb) continue;

            // ---- unit complete: all 32 lanes normalise, split and store their 4 columns of their group's row
            const int tl = ctl;
            const int r = (cu % UPT) * 4 + grp;
            const int s = tl % C::STAGES;
            if (tl != last_tl) {
                mbar_wait(empty_bar(s), ((uint32_t)(tl / C::STAGES) & 1u) ^ 1u);
                last_tl = tl;
            }
            if (citems > 0) {
                const float inv = 1.0f / fmaxf(wsum, 1e-12f);
                const float4 mean = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
                unsigned char* st = smem_gen + (size_t)s * C::STAGE_BYTES;
                uint2 parts[NS];
                split4<NS>(self, parts);
                uint32_t off = b_offset(r, 4 * sub);
#pragma unroll
                for (int p = 0; p < NS; ++p) *reinterpret_cast<uint2*>(st + p * C::PART_BYTES + off) = parts[p];
                split4<NS>(mean, parts);
                off = b_offset(r, F + 4 * sub);
#pragma unroll
                for (int p = 0; p < NS; ++p) *reinterpret_cast<uint2*>(st + p * C::PART_BYTES + off) = parts[p];
            }
            fence_proxy_async();
            acc = make_float4(0.f, 0.f, 0.f, 0.f); self = acc; wsum = 0.f;
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(full_bar(s));
                mbar_arrive(idxe_bar(tl % IDX_STAGES));                  // one arrival per unit (the stage has UPT of them)
            }
            cu += GATHER_WARPS;
            cready = false;
        }
        cp_async_wait<0>();
      } else {
        // =========================================================== gather warps (producers of the B operand)
        // Row stream of this warp: i = g, g+G, g+2G, ... over the CTA's tiles.  A row is the item list
        // [self, edge_0 .. edge_{deg-1}], cut into sub-batches of QS items.  Every item is ONE bulk async copy of a
        // whole source row (F*4 bytes) into a slot of this warp's shared-memory ring, issued by the lane that owns
        // the item; the slot's mbarrier counts the bytes.  NQ-1 sub-batches stay in flight while the oldest is
        // consumed (LDS + FFMA), so memory-level parallelism is bounded by the ring, not by registers, no load ever
        // stalls a scoreboard, and a source row costs one copy instruction instead of one per 16 bytes.
        constexpr int G = 32 / LPR;                                     // items consumed by one warp-wide LDS.128
        constexpr int QS = C::QS;
        constexpr int SLOT_FLOATS = QS * F;
        constexpr uint32_t ROW_BYTES = F * 4;
        const int g = warp - GATHER_WARP0;
        const int grp = lane / LPR, sub = lane % LPR;
        // LDGSTS transport: lane (grp, sub) moves bytes [16 sub, 16 sub + 16) of item t + grp; bulk: whole rows
        const uint32_t ring_u32 = ring_base + (uint32_t)(g * C::RING_BYTES) + (BULK ? 0u : (uint32_t)((grp * F + 4 * sub) * 4));
        const float* xs = x + 4 * sub;
        const float* ring_gen = reinterpret_cast<const float*>(smem_gen + (ring_base - smem_base) + (size_t)g * C::RING_BYTES) + grp * F + 4 * sub;
        int stream_end = 0x7fffff00;                                     // rows of this CTA's tile stream; known once the
                                                                          // end-of-stream entry (-1) shows up
        int itile = 0, ctile = 0;

        // ---- per-row state, issue side (i*) and consume side (c*)
        int ii = g, ib = 0, inb = 1, ie0 = 0, ideg = -1, itl = -1, ilong = -1, iitems = 0;  uint32_t irow = 0;
        const int32_t* icol = nullptr; int64_t ielo = 0;
        int ci = g, cb = 0, cnb = 1, ce0 = 0, cdeg = -1, ctl = -1, clong = -1, citems = 0;
        const float* cew = nullptr; int64_t celo = 0;

        auto setup_issue_row = [&]() {                                  // ii < stream_end
            const int tl = ii / TN, r = ii % TN;
            if (tl != itl) {
                const int q = tl % IDX_STAGES;
                mbar_wait(idxf_bar(q), (uint32_t)(tl / IDX_STAGES) & 1u);          // edge block staged
                itl = tl;
                itile = tile_ids[tl % TILE_RING];
                if (itile < 0) stream_end = (tl + 1) * TN;                        // a row-less closing tile: it carries the
                                                                                  // end of stream through every pipeline
            }
            const unsigned char* ibp = idx_gen + (size_t)(tl % IDX_STAGES) * IDX_STAGE_BYTES;
            const int32_t* rp_s = reinterpret_cast<const int32_t*>(ibp + EPAD * 8);
            icol = reinterpret_cast<const int32_t*>(ibp) + rp_s[96 + 2];            // + alignment offset of the col slice
            ielo = *reinterpret_cast<const int64_t*>(ibp + EPAD * 8 + 384);
            const int64_t row = row_begin + (int64_t)itile * TN + r;
            ib = 0; inb = 1; ideg = -1;
            if (itile >= 0 && row < row_end) {
                irow = (uint32_t)row;
                ie0 = rp_s[r];
                ideg = rp_s[r + 1] - ie0;
                iitems = ideg + 1;                                       // [self, edges...]
                ilong = -1;
                if (LONG && ideg > LONG_T && lw.cap > 0) {               // hub row: [self, chunk partials...] if pre-aggregated
                    ilong = long_lookup(lw, row);
                    if (ilong >= 0) iitems = 1 + (ideg + LONG_CH - 1) / LONG_CH;
                }
                inb = (iitems + QS - 1) / QS;
            }
        };
        auto setup_consume_row = [&]() {                                // ci < stream_end, tile already staged
            const int tl = ci / TN, r = ci % TN;
            if (tl != ctl) ctile = tile_ids[tl % TILE_RING];
            ctl = tl;
            const unsigned char* ibp = idx_gen + (size_t)(tl % IDX_STAGES) * IDX_STAGE_BYTES;
            const int32_t* rp_s = reinterpret_cast<const int32_t*>(ibp + EPAD * 8);
            cew = reinterpret_cast<const float*>(ibp + EPAD * 4) + rp_s[96 + 3];    // + alignment offset of the ew slice
            celo = *reinterpret_cast<const int64_t*>(ibp + EPAD * 8 + 384);
            const int64_t row = row_begin + (int64_t)ctile * TN + r;
            cb = 0; cnb = 1; cdeg = -1;
            if (ctile >= 0 && row < row_end) {
                ce0 = rp_s[r];
                cdeg = rp_s[r + 1] - ce0;
                citems = cdeg + 1;
                clong = -1;
                if (LONG && cdeg > LONG_T && lw.cap > 0) {
                    clong = long_lookup(lw, row);
                    if (clong >= 0) citems = 1 + (cdeg + LONG_CH - 1) / LONG_CH;
                }
                cnb = (citems + QS - 1) / QS;
            }
        };
        // LDGSTS transport: one cp.async group = sub-batch `ib` of the issue row, into ring slot `slot`
        auto issue_ldgsts = [&](int slot) {
            if (ii < stream_end) {
                if (ideg >= 0) {
                    const int first = ib * QS;
                    const int nitems = iitems - first;                   // items left in the row (>= 1)
                    const uint32_t dst = ring_u32 + (uint32_t)(slot * SLOT_FLOATS * 4);
                    const int kbase = ie0 + first - 1;                   // item `it` of this sub-batch is edge kbase + it
                    if (!(LONG && ilong >= 0) && ie0 + ideg <= EMAX) {   // whole row staged (the common case): branch free
                        uint32_t src[QS / G];
#pragma unroll
                        for (int t = 0; t < QS; t += G) src[t / G] = (uint32_t)icol[kbase + t + grp];   // slack reads stay in smem
                        if (first == 0 && grp == 0) src[0] = irow;       // item 0 of the row is the self row
#pragma unroll
                        for (int t = 0; t < QS; t += G)
                            cp_async16_pred(dst + (uint32_t)(t * F * 4), xs + (size_t)src[t / G] * F, t + grp < nitems);
                    } else if (LONG && ilong >= 0) {                     // hub row: items are pre-aggregated chunk partials
#pragma unroll
                        for (int t = 0; t < QS; t += G) {
                            const int it = t + grp;
                            if (it < nitems) {
                                const float* srcp = (first + it == 0) ? xs + (size_t)irow * F
                                                                      : lw.partial + (size_t)(ilong + first + it - 1) * 128 + 4 * sub;
                                cp_async16(dst + (uint32_t)(t * F * 4), srcp);
                            }
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < QS; t += G) {
                            const int it = t + grp;
                            if (it < nitems) {
                                const int k = kbase + it;
                                uint32_t src = irow;
                                if (first + it > 0) src = (uint32_t)((k < EMAX) ? icol[k] : __ldg(col + ielo + k));
                                cp_async16(dst + (uint32_t)(t * F * 4), xs + (size_t)src * F);
                            }
                        }
                    }
                }
                if (++ib >= inb) {
                    ii += GATHER_WARPS;
                    if (ii < stream_end) setup_issue_row();
                }
            }
            cp_async_commit();                                            // (possibly empty) group keeps the count in step
        };
        // sub-batch `ib` of the issue row -> ring slot `slot`: lane j < nitems copies item j (one bulk copy per source row)
        auto issue_bulk = [&](int slot) {
            const uint32_t bar = ring_bar(g, slot);
            uint32_t tx = 0;
            if (ii < stream_end) {
                if (ideg >= 0) {
                    const int first = ib * QS;
                    int nitems = iitems - first;                         // items left in the row (>= 1)
                    nitems = nitems < QS ? nitems : QS;
                    tx = (uint32_t)nitems * ROW_BYTES;
                    if (lane == 0) mbar_arrive_expect_tx(bar, tx);
                    if (lane < nitems) {
                        const int it = first + lane;                     // item number inside the row; 0 = the self row
                        const int k = ie0 + it - 1;                      // its edge, relative to the staged slice
                        const float* srcp;
                        if (it == 0) srcp = x + (size_t)irow * F;
                        else if (LONG && ilong >= 0) srcp = lw.partial + (size_t)(ilong + it - 1) * 128;   // pre-aggregated chunk
                        else srcp = x + (size_t)(uint32_t)((k < EMAX) ? icol[k] : __ldg(col + ielo + k)) * F;
                        bulk_g2s(ring_u32 + (uint32_t)((slot * QS + lane) * (int)ROW_BYTES), srcp, ROW_BYTES, bar);
                    }
                }
                if (++ib >= inb) {
                    ii += GATHER_WARPS;
                    if (ii < stream_end) setup_issue_row();
                }
            }
            if (tx == 0 && lane == 0) mbar_arrive(bar);                   // empty sub-batch: keeps the slot's phase in step
        };

        auto issue = [&](int slot) { if constexpr (BULK) issue_bulk(slot); else issue_ldgsts(slot); };

        if (ii < stream_end) { setup_issue_row(); }
        if (ci < stream_end) { setup_consume_row(); }
        int islot = 0, cslot = 0;
#pragma unroll
        for (int d = 0; d < NQ - 1; ++d) { issue(islot); islot = (islot + 1) % NQ; }   // prologue: NQ-1 groups in flight
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), self = acc;
        float wsum = 0.f;
        int last_tl = -1;
        uint32_t ring_phase = 0;                                          // bit q = parity to wait for on ring slot q
        while (ci < stream_end) {
            if constexpr (BULK) {
                mbar_wait(ring_bar(g, cslot), (ring_phase >> cslot) & 1u);    // oldest sub-batch landed (all its bytes)
                ring_phase ^= 1u << cslot;
            } else {
                cp_async_wait<NQ - 2>();                                  // oldest group landed (this lane's part) ...
            }
            __syncwarp();                                                 // ... and everyone else's;                                                 // all lanes are done reading the slot consumed
                                                                          // last iteration
            issue(islot);                                                 // refill that slot
            islot = (islot + 1) % NQ;
            if (cdeg >= 0) {
                const int first = cb * QS;
                const int nitems = citems - first;
                const float* rs = ring_gen + cslot * SLOT_FLOATS;
                const int kbase = ce0 + first - 1;
                if (!(LONG && clong >= 0) && ce0 + cdeg <= EMAX) {       // whole row staged: batched LDS, predicated math
                    float4 v[QS / G];
                    float w[QS / G];
#pragma unroll
                    for (int t = 0; t < QS; t += G) {
                        v[t / G] = *reinterpret_cast<const float4*>(rs + t * F);
                        w[t / G] = cew[kbase + t + grp];
                    }
                    if (first == 0 && grp == 0) { self = v[0]; w[0] = 0.f; v[0] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                    for (int t = 0; t < QS; t += G) {
                        if (t + grp < nitems) {
                            const float wt = w[t / G];
                            const float4 vt = v[t / G];
                            wsum += wt;
                            acc.x = fmaf(wt, vt.x, acc.x); acc.y = fmaf(wt, vt.y, acc.y);
                            acc.z = fmaf(wt, vt.z, acc.z); acc.w = fmaf(wt, vt.w, acc.w);
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < QS; t += G) {
                        const int it = t + grp;
                        if (it < nitems) {
                            const float4 v = *reinterpret_cast<const float4*>(rs + t * F);
                            if (t == 0 && first + it == 0) {
                                self = v;
                            } else if (LONG && clong >= 0) {             // chunk partial: already weighted
                                wsum += lw.pw[clong + first + it - 1];
                                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                            } else {
                                const int k = kbase + it;
                                const float w = (k < EMAX) ? cew[k] : __ldg(ew + celo + k);
                                wsum += w;
                                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                                acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                            }
                        }
                    }
                }
            }
            cslot = (cslot + 1) % NQ;
            if (++cb < cnb) continue;

            // ---- row complete: normalise, split to bf16 terms, write the UMMA B operand, signal
            const int tl = ctl;
            const int r = ci % TN;
            const int s = tl % C::STAGES;
            if (G > 1) {                                                  // combine the G lane groups
#pragma unroll
                for (int o = LPR; o < 32; o <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
                    self.x += __shfl_xor_sync(0xffffffffu, self.x, o); self.y += __shfl_xor_sync(0xffffffffu, self.y, o);
                    self.z += __shfl_xor_sync(0xffffffffu, self.z, o); self.w += __shfl_xor_sync(0xffffffffu, self.w, o);
                }
            }
            if (tl != last_tl) {                                         // first write of this warp into the stage
                mbar_wait(empty_bar(s), ((uint32_t)(tl / C::STAGES) & 1u) ^ 1u);
                last_tl = tl;
            }
            if (cdeg >= 0) {
                const float inv = 1.0f / fmaxf(wsum, 1e-12f);
                const float4 mean = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
                if (lane < LPR) {
                    unsigned char* st = smem_gen + (size_t)s * C::STAGE_BYTES;
                    uint2 parts[NS];
                    split4<NS>(self, parts);
                    uint32_t off = b_offset(r, 4 * lane);
#pragma unroll
                    for (int p = 0; p < NS; ++p) *reinterpret_cast<uint2*>(st + p * C::PART_BYTES + off) = parts[p];
                    split4<NS>(mean, parts);
                    off = b_offset(r, F + 4 * lane);
#pragma unroll
                    for (int p = 0; p < NS; ++p) *reinterpret_cast<uint2*>(st + p * C::PART_BYTES + off) = parts[p];
                }
                fence_proxy_async();                                     // generic-proxy writes -> async proxy (UMMA)
                acc = make_float4(0.f, 0.f, 0.f, 0.f); self = acc; wsum = 0.f;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(full_bar(s));
                if ((ci + GATHER_WARPS) / TN != tl) mbar_arrive(idxe_bar(tl % IDX_STAGES));   // this warp's last row in the tile
            }
            ci += GATHER_WARPS;
            if (ci < stream_end) setup_consume_row();
        }
        if constexpr (!BULK) cp_async_wait<0>();
      }
    } else if (warp == LOADER_WARP) {
        // =========================================================== edge-block loader
        constexpr int RPW = (TN + 1 + 31) / 32;
        auto load_rp = [&](int64_t tile, int64_t (&rp)[RPW]) {          // rowptr[row0 .. row0+TN], clamped to row_end
            const int64_t row0 = row_begin + tile * TN;
#pragma unroll
            for (int k = 0; k < RPW; ++k) {
                int64_t rr = row0 + lane + 32 * k;
                if (rr > row_end) rr = row_end;
                rp[k] = (lane + 32 * k <= TN) ? (int64_t)rowptr[rr] : 0;
            }
        };
        // Software pipeline, two tiles deep and free of register copies (the loop body is instantiated twice with the roles of
        // the two rowptr register sets swapped): the atomic that claims tile t+2 and the rowptr loads of tile t+1 are issued
        // in iteration t and first needed one iteration later, so neither latency (~1 us each) sits on the per-tile path --
        // at F=32 a tile period is < 2 us and the r1/r2 loader (claim -> dependent loads -> use, serial) was as slow as that.
        unsigned craw = 0;                                               // lane 0: raw result of the claim in flight
        auto claim_issue = [&]() { if (lane == 0) craw = atomicAdd(tile_ctr, 1u); };
        auto claim_take = [&]() -> int64_t { return (int64_t)__shfl_sync(0xffffffffu, craw, 0); };
        // one iteration: stage tile `tile` (its rowptr words in rp_cur) as the CTA's tl-th tile; returns true at the end of stream
        auto body = [&](int64_t tl, int64_t tile, int64_t (&rp_cur)[RPW], int64_t& tile_nx, int64_t (&rp_nx)[RPW]) -> bool {
            const bool end = tile >= n_tiles;
            tile_nx = n_tiles;
            if (!end) {
                tile_nx = claim_take();                                  // claimed one iteration ago
                claim_issue();                                           // for the iteration after the next
                if (tile_nx < n_tiles) load_rp(tile_nx, rp_nx);          // needed in the next iteration
            }
            const int q = (int)(tl % IDX_STAGES);
            mbar_wait_relaxed(idxe_bar(q), ((uint32_t)(tl / IDX_STAGES) & 1u) ^ 1u);
            if (lane == 0) {
                tile_ids[tl % TILE_RING] = end ? -1 : (int)tile;
                if (end) *end_tl = (int)tl;                              // ordered before the closing release below
            }
            if (end) {                                                   // closing entry: no rows, no edges
                __syncwarp();
                if (lane == 0) mbar_arrive_expect_tx(idxf_bar(q), 0u);
                return true;
            }
            const int64_t e_lo = __shfl_sync(0xffffffffu, rp_cur[0], 0);
            const int64_t e_hi = __shfl_sync(0xffffffffu, rp_cur[TN / 32], TN % 32);
            unsigned char* ib = idx_gen + (size_t)q * IDX_STAGE_BYTES;
            int32_t* rp_s = reinterpret_cast<int32_t*>(ib + EPAD * 8);
#pragma unroll
            for (int k = 0; k < RPW; ++k)
                if (lane + 32 * k <= TN) rp_s[lane + 32 * k] = (int32_t)(rp_cur[k] - e_lo);
            const int n = (int)((e_hi - e_lo) < EMAX ? (e_hi - e_lo) : EMAX);
            // the col / ew slices [e_lo, e_lo + n) as the 16-byte aligned supersets bulk copies need: the slice starts
            // `off` elements into its staged array (the gather warps add it); an over-read stays inside the 16-byte
            // line that holds the first / last element
            const uintptr_t ca = reinterpret_cast<uintptr_t>(col + e_lo), wa = reinterpret_cast<uintptr_t>(ew + e_lo);
            const uint32_t c_off = (uint32_t)(ca & 15u) >> 2, w_off = (uint32_t)(wa & 15u) >> 2;
            const uint32_t c_bytes = n ? (((uint32_t)n + c_off) * 4u + 15u) & ~15u : 0u;
            const uint32_t w_bytes = n ? (((uint32_t)n + w_off) * 4u + 15u) & ~15u : 0u;
            if (lane == 0) {
                *reinterpret_cast<int64_t*>(ib + EPAD * 8 + 384) = e_lo;
                rp_s[96 + 2] = (int32_t)c_off; rp_s[96 + 3] = (int32_t)w_off;
            }
            __syncwarp();                                                // every lane's rowptr words are written ...
            if (lane == 0) {
                const uint32_t cs = idx_base + (uint32_t)(q * IDX_STAGE_BYTES), ws = cs + EPAD * 4;
                mbar_arrive_expect_tx(idxf_bar(q), c_bytes + w_bytes);   // ... and published by this release-arrive
                if (n) {
                    bulk_g2s(cs, reinterpret_cast<const void*>(ca & ~(uintptr_t)15), c_bytes, idxf_bar(q));
                    bulk_g2s(ws, reinterpret_cast<const void*>(wa & ~(uintptr_t)15), w_bytes, idxf_bar(q));
                }
            }
            return false;
        };
        int64_t rp_a[RPW], rp_b[RPW];
        claim_issue();
        int64_t tile_a = claim_take(), tile_b = n_tiles;
        claim_issue();                                                   // in flight for the first body()
        if (tile_a < n_tiles) load_rp(tile_a, rp_a);
        for (int64_t tl = 0;; tl += 2) {
            if (body(tl, tile_a, rp_a, tile_b, rp_b)) break;
            if (body(tl + 1, tile_b, rp_b, tile_a, rp_a)) break;
        }
    } else if (warp == MMA_WARP) {
        // =========================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc();
            for (int64_t tl = 0;; ++tl) {
                const int s = (int)(tl % C::STAGES);
                const uint32_t n = (uint32_t)(tl / C::STAGES);
                const int a = (int)(tl & 1);
                const uint32_t m = (uint32_t)(tl >> 1);
                mbar_wait_relaxed(acce_bar(a), (m & 1u) ^ 1u);           // epilogue drained this accumulator
                mbar_wait_relaxed(full_bar(s), n & 1u);                  // gather filled this stage
                tc_fence_after();
                if (tile_ids[tl % TILE_RING] < 0) {                      // closing entry: hand the end of stream to the epilogue
                    umma_commit(accf_bar(a));
                    break;
                }
                const uint32_t d_tmem = tmem_base + (uint32_t)(C::ACC_COL0 + a * TN);
                const uint32_t st0 = smem_base + (uint32_t)(s * C::STAGE_BYTES);
#pragma unroll 2
                for (int j = 0; j < C::KSTEPS; ++j) {
                    const uint32_t boff = (uint32_t)((j >> 2) * (TN * 128) + (j & 3) * 32);
                    uint64_t xb[NS];
                    uint32_t wa[NS];
#pragma unroll
                    for (int p = 0; p < NS; ++p) {
                        xb[p] = make_b_desc(st0 + (uint32_t)(p * C::PART_BYTES) + boff);     // x term p (smem)
                        wa[p] = tmem_base + (uint32_t)(p * C::W_PART_COLS + j * 8);          // W^T term p (TMEM)
                    }
                    umma_ts(d_tmem, wa[0], xb[0], idesc, j > 0 ? 1u : 0u);
                    umma_ts(d_tmem, wa[0], xb[1], idesc, 1u);
                    umma_ts(d_tmem, wa[1], xb[0], idesc, 1u);
                    if constexpr (NS == 3) {
                        umma_ts(d_tmem, wa[1], xb[1], idesc, 1u);
                        umma_ts(d_tmem, wa[0], xb[2], idesc, 1u);
                        umma_ts(d_tmem, wa[2], xb[0], idesc, 1u);
                    }
                }
                umma_commit(empty_bar(s));                               // smem stage reusable once the MMAs retire
                umma_commit(accf_bar(a));                                // accumulator ready for the epilogue
            }
        }
    } else {
        // =========================================================== epilogue warps
        const int f = warp * 32 + lane;
        const float my_nw = node_w ? __ldg(node_w + f) : 0.f;
        for (int64_t tl = 0;; ++tl) {
            const int a = (int)(tl & 1);
            const uint32_t m = (uint32_t)(tl >> 1);
            mbar_wait_relaxed(accf_bar(a), m & 1u);
            tc_fence_after();
            const int64_t tile = tile_ids[tl % TILE_RING];
            if (tile < 0) break;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(C::ACC_COL0 + a * TN);
            const int64_t row0 = row_begin + tile * TN;
            uint32_t v[32];
            tmem_ld32(taddr, v);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(acce_bar(a));                                    // accumulator drained: the MMA may reuse it
            float hv[32], ov[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float o = __uint_as_float(v[j]) + my_bias;
                if (relu) o = fmaxf(o, 0.f);
                ov[j] = o;
                hv[j] = o * my_nw;
            }
            if (peers.n == 0 && !peers.mc) {
                // single GPU: thread = feature, 32 row stores of 128 B per warp (no extra shuffles: the epilogue
                // shares issue slots with the gather warps)
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (row0 + j < row_end) out[(row0 + j) * UM + f] = ov[j];
            } else {
                // fused exchange: transpose 4x4 inside each lane quad so that a lane owns 4 consecutive features of
                // one row (16-byte stores; a warp instruction writes four full 128-byte row segments) -- locally and
                // to every peer's buffer over NVLink, or once through the NVSwitch multicast address.
                const int qi = lane & 3;                                  // row within the block of 4
                const int fq = warp * 32 + (lane & ~3);                   // first of this lane's 4 features
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float a0 = ov[4 * r], a1 = ov[4 * r + 1], a2 = ov[4 * r + 2], a3 = ov[4 * r + 3];
                    quad_transpose(a0, a1, a2, a3, lane);
                    const int64_t row = row0 + 4 * r + qi;
                    if (row < row_end) {
                        const int64_t off = row * UM + fq;
                        const float4 val = make_float4(a0, a1, a2, a3);
                        if (peers.mc) {
                            multimem_st4(peers.mc + off, val);
                        } else {
                            *reinterpret_cast<float4*>(out + off) = val;
                            const unsigned need = peers.need ? (unsigned)__ldg(peers.need + row) : 0xFFu;
#pragma unroll
                            for (int pr = 0; pr < MAX_PEERS; ++pr)
                                if (pr < peers.n && ((need >> pr) & 1u)) *reinterpret_cast<float4*>(peers.out[pr] + off) = val;
                        }
                    }
                }
            }
            if (node_w) {
                // fused node head: score[row] = sigmoid(h[row,:] . node_w + node_b)
                const float part = warp_transpose_reduce32(hv, lane);    // lane l: this warp's 32 features of row l
                head_red[(a * EPI_WARPS + warp) * 32 + lane] = part;
                asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps
                if (warp == 0) {
                    const float* hr = head_red + a * EPI_WARPS * 32 + lane;
                    const float dot = (hr[0] + hr[32]) + (hr[64] + hr[96]);
                    if (row0 + lane < row_end) score[row0 + lane] = 1.f / (1.f + expf(-(dot + node_b)));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// the per-row bulk-copy transport is an experiment knob (NERRF_SAGE_ROW_COPY=bulk), compiled for the headline shapes only
template <int F, int NS, typename RP>
constexpr bool kBulkVariant = (NS == 3) && (F == 32 || F == 128) && (sizeof(RP) == 4);
inline bool bulk_rows_requested() {
    static const bool v = [] { const char* e = getenv("NERRF_SAGE_ROW_COPY"); return e && e[0] == 'b'; }();
    return v;
}

// per-device ring of launch-private tile counters (1 KB, allocated on first use and kept): a launch zeroes its slot on its
// own stream, so launches on different streams do not share a counter unless 256 of them are in flight at once
inline unsigned* tile_counter_slot(cudaStream_t st) {
    static unsigned* ring[64] = {};
    static unsigned next[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    dev_ &= 63;
    if (!ring[dev_] && cudaMalloc((void**)&ring[dev_], 256 * sizeof(unsigned)) != cudaSuccess) return nullptr;
    unsigned* slot = ring[dev_] + (__atomic_fetch_add(&next[dev_], 1u, __ATOMIC_RELAXED) & 255u);
    if (cudaMemsetAsync(slot, 0, sizeof(unsigned), st) != cudaSuccess) return nullptr;
    return slot;
}

template <int F, int NS, typename RP>
int launch_umma(const float* x, const RP* rowptr, const int32_t* col, const float* ew, const float* W, const float* b,
                float* out, int64_t row_begin, int64_t row_end, int relu, const float* node_w, float node_b, float* score,
                void* long_ws, size_t long_ws_bytes, bool reuse_scan, const Peers& peers, cudaStream_t st) {
    using C = UmmaCfg<F, NS>;
    static bool attr_set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& attr_set = attr_set_dev[dev_ & 63];          // function attributes are per device
    if (!attr_set) {
        NERRF_CHECK_CUDA(cudaFuncSetAttribute(sage_layer_umma_kernel<F, NS, RP, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        NERRF_CHECK_CUDA(cudaFuncSetAttribute(sage_layer_umma_kernel<F, NS, RP, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        if constexpr (kBulkVariant<F, NS, RP>)
            NERRF_CHECK_CUDA(cudaFuncSetAttribute(sage_layer_umma_kernel<F, NS, RP, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        attr_set = true;
    }
    const int64_t rows = row_end - row_begin;
    const int64_t tiles = (rows + TN - 1) / TN;
    if (tiles == 0) return NERRF_OK;
    LongWs lw = long_ws_carve(long_ws, long_ws_bytes);
    if (lw.cap > 0) {
        // hub-row pre-pass: clear header + hash, mark the queue empty, find long rows, aggregate their chunks
        const int sms = sm_count();
        if (!reuse_scan) {       // the scan (hub-row list, hash, chunk queue) depends on the graph only, not on x
            const size_t hash_bytes = (size_t)(lw.hash_mask + 1) * 8;
            NERRF_CHECK_CUDA(cudaMemsetAsync(lw.hdr, 0, 256 + hash_bytes, st));
            NERRF_CHECK_CUDA(cudaMemsetAsync(lw.queue, 0xFF, (size_t)lw.cap * 8, st));
            long_scan_kernel<RP><<<sms * 4, 256, 0, st>>>(rowptr, row_begin, row_end, lw);
        }
        long_chunk_kernel<F, RP><<<sms * 4, 256, 0, st>>>(x, rowptr, col, ew, lw);
        int rc = launch_status("long-row pre-pass");
        if (rc) return rc;
    }
    const int64_t grid = tiles < sm_count() ? tiles : sm_count();
    unsigned* tile_ctr = tile_counter_slot(st);
    if (!tile_ctr) { set_error("could not allocate the tile counter"); return NERRF_ERR_CUDA; }
    // the per-row bulk-copy transport needs 16-byte aligned rows; every transport stages col / ew with bulk copies,
    // which only need 4-byte aligned slices (the aligned superset is copied)
    if (lw.cap > 0)
        sage_layer_umma_kernel<F, NS, RP, true, false><<<(unsigned)grid, C::THREADS, C::SMEM, st>>>(x, rowptr, col, ew, W, b, out, row_begin, row_end, relu, node_w, node_b, score, lw, peers, tile_ctr);
    else if (kBulkVariant<F, NS, RP> && bulk_rows_requested() && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        if constexpr (kBulkVariant<F, NS, RP>)
            sage_layer_umma_kernel<F, NS, RP, false, true><<<(unsigned)grid, C::THREADS, C::SMEM, st>>>(x, rowptr, col, ew, W, b, out, row_begin, row_end, relu, node_w, node_b, score, lw, peers, tile_ctr);
    } else
        sage_layer_umma_kernel<F, NS, RP, false, false><<<(unsigned)grid, C::THREADS, C::SMEM, st>>>(x, rowptr, col, ew, W, b, out, row_begin, row_end, relu, node_w, node_b, score, lw, peers, tile_ctr);
    return launch_status("sage_layer_umma_kernel");
}

}  // namespace

bool sage_umma_available() { return true; }

int sage_layer_umma(const float* x, const void* rowptr, int is64, const int32_t* col, const float* ew, const float* W,
                    const float* b, float* out, int64_t n_nodes, int64_t row_begin, int64_t row_end, int F, int relu,
                    int nsplit, const float* node_w, float node_b, float* score, void* long_ws, size_t long_ws_bytes,
                    bool reuse_scan, float* const* peer_out, int n_peers, const uint8_t* peer_need, cudaStream_t st) {
    (void)n_nodes;
    Peers peers{};
    peers.need = peer_need;
    if (n_peers == -1) {                      // peer_out[0] is a multicast address covering every rank's buffer
        peers.mc = peer_out[0];
    } else {
        if (n_peers < 0 || n_peers > MAX_PEERS) { set_error("n_peers must be -1 (multicast) or 0..%d", MAX_PEERS); return NERRF_ERR_INVALID; }
        for (int i = 0; i < n_peers; ++i) peers.out[i] = peer_out[i];
        peers.n = n_peers;
    }
#define GO(FV, NSV)                                                                                                                                                                               \
    return is64 ? launch_umma<FV, NSV, int64_t>(x, (const int64_t*)rowptr, col, ew, W, b, out, row_begin, row_end, relu, node_w, node_b, score, long_ws, long_ws_bytes, reuse_scan, peers, st) \
                : launch_umma<FV, NSV, int32_t>(x, (const int32_t*)rowptr, col, ew, W, b, out, row_begin, row_end, relu, node_w, node_b, score, long_ws, long_ws_bytes, reuse_scan, peers, st)
    if (nsplit == 3) {
        switch (F) {
            case 32: GO(32, 3);
            case 64: GO(64, 3);
            case 128: GO(128, 3);
        }
    } else if (nsplit == 2) {
        switch (F) {
            case 32: GO(32, 2);
            case 64: GO(64, 2);
            case 128: GO(128, 2);
        }
    }
#undef GO
    set_error("UMMA layer: unsupported feature width F=%d / split %d (supported: F in 32, 64, 128; split 2 or 3)", F, nsplit);
    return NERRF_ERR_INVALID;
}

}  // namespace nerrf
