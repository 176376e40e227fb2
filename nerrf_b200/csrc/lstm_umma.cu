// BiLSTM on 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only -- the recurrent half of lstm.forward.
//
// One LSTM layer-direction step is   gates[4H, seqs] = G_x[:, t] + W_hh[4H, H] . h_{t-1}[H, seqs]   followed by the cell
// update.  The input projection G_x = X W_ih^T + b has no recurrence: it is computed for all (b, t) up front by the
// fused GraphSAGE-T layer kernel used as a dense GEMM (sage_umma.cu with an identity-shift graph: x_v = features
// [0,128) of row v, its single in-edge brings features [128,256) -- see lstm_forward_umma below).  This file is the
// recurrence:
//
//   * W_hh (1024 x 256) is cut into 8 slices of 128 gate rows = the 4 gates of 32 hidden units; slice s lives in the
//     tensor memory of one CTA for the whole launch (A operand, bf16 x 3 split, 384 columns), exactly like W^T in
//     sage_umma.cu.  8 CTAs (one per slice) form a GROUP that advances one tile of 128 sequences of one direction.
//   * per step every CTA of the group computes gates^T[128, 128 seqs] (+)= W_slice . h_{t-1}^T: the accumulator
//     (128 TMEM columns) is pre-loaded with G_x[:, t] by the epilogue warps (tcgen05.st) while the group is still
//     exchanging h, so the 96 MMAs (16 k-steps x 6 split products) only ever accumulate.
//   * epilogue (8 warps): tcgen05.ld the gates, activations, a shared-memory exchange brings the four gates of a unit
//     to one thread, cell update with c and h in registers, h slice (32 units x 128 seqs) to the group's ping-pong
//     buffer in global memory (L2) and to the layer output, then ONE release-increment of the group's step counter.
//   * producer warps (8) acquire-poll the counter (bounded spin), read the full h_t (128 x 256 fp32, ld.global.cg),
//     split it into three bf16 planes in the K-major SWIZZLE_128B layout the UMMA B operand wants, and signal the MMA
//     issuer.
//
// Groups are independent (no grid barrier): group g owns direction g & 1 and the sequence tiles (g >> 1) + k * G/2.
// The 8 CTAs of a group must be co-resident: cooperative launch, one CTA per SM, grid = 8 * G <= #SMs.
//
// STATUS: parity-green on B200 against the oracle (max err 1.2e-7 on B=6/130/300; scripts/lstm_umma_check.py); first
// timing 56 ms for B=4096, T=100 (v0 FFMA kernel: 52 ms) -- opt-in (NERRF_LSTM_ALGO=umma) until it is faster.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace nerrf {

int sage_layer_umma(const float* x, const void* rowptr, int is64, const int32_t* col, const float* ew, const float* W,
                    const float* b, float* out, int64_t n_nodes, int64_t row_begin, int64_t row_end, int F, int relu,
                    int nsplit, const float* node_w, float node_b, float* score, void* long_ws, size_t long_ws_bytes,
                    bool reuse_scan, float* const* peer_out, int n_peers, const uint8_t* peer_need, cudaStream_t st);

namespace {

constexpr int LH = 256;                 // hidden size
constexpr int NSEQ = 128;               // sequences per tile = UMMA N
constexpr int SLICES = 8;               // CTAs per group
constexpr int UM = 128;                 // gate rows per slice = UMMA M
constexpr int NS = 3;                   // bf16 terms per fp32 value
constexpr int KB = LH / 64;             // 128-byte K blocks of the B operand
constexpr int KSTEPS = LH / 16;
constexpr int PART_BYTES = KB * NSEQ * 128;          // 64 KB: one bf16 plane of h [128 seqs, 256]
constexpr int OPER_BYTES = NS * PART_BYTES;          // 192 KB
constexpr int GATE_BYTES = 4 * 32 * 32 * 4;          // 16 KB: [gate][seq in chunk][unit] per epilogue half
constexpr int W_PART_COLS = LH / 2;                  // TMEM columns of one W plane
constexpr int ACC_COL0 = NS * W_PART_COLS;           // 384
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 8;                         // warps 0-3: sequences 0-63, warps 4-7: sequences 64-127
constexpr int MMA_WARP = 8;
constexpr int PROD_WARP0 = 9;
constexpr int PROD_WARPS = 8;
constexpr int THREADS = (PROD_WARP0 + PROD_WARPS) * 32;          // 544
constexpr size_t SMEM = (size_t)OPER_BYTES + 2 * GATE_BYTES + NSEQ * 4 + 1024 /*align*/ + 64 /*barriers*/;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
static_assert(ACC_COL0 + NSEQ <= TMEM_COLS, "TMEM budget");

// ---------------------------------------------------------------- PTX wrappers (same forms as sage_umma.cu)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// every wait in this kernel is bounded: a protocol bug must end in a trap (launch error), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin) {
        if (spin > (1u << 16)) __nanosleep(64);
        if (spin > (1u << 26)) __trap();
    }
}
// roles with slack poll with a sleep so they do not take issue slots from the producer warps
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin) {
        __nanosleep(100);
        if (spin > (1u << 25)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
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
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc() {      // D=f32, A=B=bf16 K-major, N = NSEQ, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NSEQ >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t (&parts)[NS]) {
#pragma unroll
    for (int p = 0; p < NS; ++p) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        parts[p] = *reinterpret_cast<const uint32_t*>(&h);
        if (p + 1 < NS) {
            const float2 f = __bfloat1622float2(h);
            a -= f.x; b -= f.y;
        }
    }
}
// byte offset of element (sequence r, hidden index k) inside one plane of the B operand (k multiple of 4)
__device__ __forceinline__ uint32_t b_offset(int r, int k) {
    const int kb = k >> 6, col = k & 63;
    const int chunk = col >> 3;
    return (uint32_t)(kb * (NSEQ * 128) + r * 128 + (((chunk ^ (r & 7)) << 4) | ((col & 7) << 1)));
}
__device__ __forceinline__ float4 ld_cg4(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// activations: |abs error| ~ 2e-7 (ex2.approx + rcp), far inside the 1e-4 spec bound; tanh(x) = 2 sigmoid(2x) - 1
__device__ __forceinline__ float sigmoid_fast(float x) { return __frcp_rn(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.0f, sigmoid_fast(2.0f * x), -1.0f); }

struct RecArgs {
    const float* gx_a;        // [2 dirs * 8 slices][R = B*T rows][128]: W_ih part (first K half) + bias, gate-slice layout
    const float* gx_b;        // second K half of the input projection, or nullptr
    const float* whh;         // [2 dirs * 8 slices][256 (k)][128 (gate row in slice)]
    const int32_t* len;       // [B]
    float* out_slabs;         // [4][R][128]: features [128q, 128q+128) of the layer output, or nullptr (top layer)
    float* hfin;              // [B][512]
    float* hx;                // [groups][2][128][256] ping-pong h exchange
    unsigned* counter;        // [groups], zero at launch
    int64_t B;
    int T;
    int groups;
};

__global__ void __launch_bounds__(THREADS, 1) lstm_rec_umma_kernel(RecArgs P) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    float* gate_s = reinterpret_cast<float*>(smem_gen + OPER_BYTES);                // [2 halves][4][32][32]
    int* len_s = reinterpret_cast<int*>(smem_gen + OPER_BYTES + 2 * GATE_BYTES);    // [128]
    const uint32_t bar_base = smem_base + OPER_BYTES + 2 * GATE_BYTES + NSEQ * 4;
    const uint32_t opfull_bar = bar_base, accinit_bar = bar_base + 8, accf_bar = bar_base + 16, tmem_slot = bar_base + 24;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = blockIdx.x / SLICES, slice = blockIdx.x % SLICES;
    const int dir = group & 1;
    const int64_t n_tiles = (P.B + NSEQ - 1) / NSEQ;
    const int tiles_stride = P.groups >> 1;                       // groups per direction
    const int64_t R = P.B * (int64_t)P.T;
    // number of items (tiles) this group processes, and global steps
    const int64_t first_tile = group >> 1;
    const int64_t n_items = first_tile < n_tiles ? (n_tiles - first_tile + tiles_stride - 1) / tiles_stride : 0;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            mbar_init(opfull_bar, PROD_WARPS * 32);
            mbar_init(accinit_bar, EPI_WARPS * 32);
            mbar_init(accf_bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // ---- W_hh slice -> TMEM (three bf16 planes), once: thread = gate row (warps 0-3)
    if (warp < 4) {
        const int f = warp * 32 + lane;
        const float* W = P.whh + (size_t)(dir * SLICES + slice) * LH * UM;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < LH; k0 += 16) {
            uint32_t parts[NS][8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = __ldg(W + (size_t)(k0 + 2 * i) * UM + f);
                const float b = __ldg(W + (size_t)(k0 + 2 * i + 1) * UM + f);
                uint32_t pp[NS];
                split_pair(a, b, pp);
#pragma unroll
                for (int p = 0; p < NS; ++p) parts[p][i] = pp[p];
            }
#pragma unroll
            for (int p = 0; p < NS; ++p) tmem_st8(lane_addr + (uint32_t)(p * W_PART_COLS + (k0 >> 1)), parts[p]);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const float* gxa = P.gx_a + (size_t)(dir * SLICES + slice) * R * UM;
    const float* gxb = P.gx_b ? P.gx_b + (size_t)(dir * SLICES + slice) * R * UM : nullptr;
    float* hx = P.hx + (size_t)group * 2 * NSEQ * LH;
    unsigned* counter = P.counter + group;

    if (warp < EPI_WARPS) {
        // =========================================================== epilogue warps
        const int half = warp >> 2;                 // sequences [64 half, 64 half + 64)
        const int wq = warp & 3;                    // TMEM lane quarter = gate type (i, f, g, o)
        const int et = (warp & 3) * 32 + lane;      // thread index inside the half (0..127)
        float* gs = gate_s + half * (GATE_BYTES / 4);
        const uint32_t acc_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ACC_COL0 + half * 64);
        const int gcol = wq * 32 + lane;            // gate row inside the slice = column of the G_x slab
        const int ucol = slice * 32 + lane;         // hidden unit of this thread in the cell-update phase
        uint32_t q = 0;                             // global step index of this group
        for (int64_t it = 0; it < n_items; ++it) {
            const int64_t b0 = (first_tile + it * tiles_stride) * NSEQ;
            // sequence lengths of the tile (both halves load their own 64)
            if (et < 64) { const int64_t b = b0 + half * 64 + et; len_s[half * 64 + et] = b < P.B ? P.len[b] : 0; }
            named_bar(1 + half, 128);
            float c_reg[16], h_reg[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { c_reg[j] = 0.f; h_reg[j] = 0.f; }

            auto init_acc = [&](int step) {        // accumulator <- G_x[:, t] for this thread's gate row, 64 sequences
                const int t = dir ? (P.T - 1 - step) : step;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    // the 16 (or 32) loads of a chunk are issued back to back: rows past the batch are clamped (their
                    // sequences have length 0, the value is never used), so there is no branch between the loads
                    float ga[16];
                    const int64_t bb = b0 + half * 64 + c * 16;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int64_t b = (bb + j < P.B) ? bb + j : P.B - 1;
                        ga[j] = __ldg(gxa + (size_t)(b * P.T + t) * UM + gcol);
                    }
                    if (gxb) {
                        float gb[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int64_t b = (bb + j < P.B) ? bb + j : P.B - 1;
                            gb[j] = __ldg(gxb + (size_t)(b * P.T + t) * UM + gcol);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) ga[j] += gb[j];
                    }
                    uint32_t v0[8], v1[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { v0[j] = __float_as_uint(ga[j]); v1[j] = __float_as_uint(ga[8 + j]); }
                    tmem_st8(acc_addr + (uint32_t)(c * 16), v0);
                    tmem_st8(acc_addr + (uint32_t)(c * 16 + 8), v1);
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(accinit_bar);
            };
            init_acc(0);

            for (int step = 0; step < P.T; ++step, ++q) {
                const int t = dir ? (P.T - 1 - step) : step;
                mbar_wait_relaxed(accf_bar, q & 1u);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    // phase A: activated gates of this thread's gate row for 32 sequences -> shared [gate][seq][unit]
                    uint32_t v[32];
                    tmem_ld32(acc_addr + (uint32_t)(c * 32), v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = __uint_as_float(v[j]);
                        gs[(wq * 32 + j) * 32 + lane] = (wq == 2) ? tanh_fast(x) : sigmoid_fast(x);
                    }
                    named_bar(1 + half, 128);
                    // phase B: cell update, thread = (unit lane, sequences wq*8 .. wq*8+7 of the chunk)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = wq * 8 + jj;                 // sequence inside the chunk
                        const int n = half * 64 + c * 32 + j;      // sequence inside the tile
                        const float ig = gs[(0 * 32 + j) * 32 + lane], fg = gs[(1 * 32 + j) * 32 + lane];
                        const float gg = gs[(2 * 32 + j) * 32 + lane], og = gs[(3 * 32 + j) * 32 + lane];
                        const bool valid = t < len_s[n];
                        float& cr = c_reg[c * 8 + jj];
                        float& hr = h_reg[c * 8 + jj];
                        float oval = 0.f;
                        if (valid) {
                            cr = fmaf(fg, cr, ig * gg);
                            hr = og * tanh_fast(cr);
                            oval = hr;
                        }
                        hx[((size_t)(q & 1u) * NSEQ + n) * LH + ucol] = hr;
                        const int64_t b = b0 + n;
                        if (P.out_slabs && b < P.B) {
                            const int feat = dir * LH + ucol;
                            P.out_slabs[((size_t)(feat >> 7) * R + (size_t)(b * P.T + t)) * UM + (feat & 127)] = oval;
                        }
                    }
                    named_bar(1 + half, 128);                       // the gate buffer is reused by the next chunk
                }
                tc_fence_before();
                named_bar(3, EPI_WARPS * 32);                       // both halves have written their h
                if (tid == 0) {
                    __threadfence();
                    atomicAdd(counter, 1u);                          // release: this CTA's h_t slice is published
                }
                if (step + 1 < P.T) init_acc(step + 1);
            }
            // final state of the tile
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int64_t b = b0 + half * 64 + c * 32 + wq * 8 + jj;
                    if (b < P.B) P.hfin[b * (2 * LH) + dir * LH + ucol] = h_reg[c * 8 + jj];
                }
        }
    } else if (warp == MMA_WARP) {
        // =========================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc();
            const uint32_t d_tmem = tmem_base + (uint32_t)ACC_COL0;
            const uint32_t total = (uint32_t)(n_items * P.T);
            for (uint32_t q = 0; q < total; ++q) {
                mbar_wait_relaxed(accinit_bar, q & 1u);              // accumulator holds G_x[:, t]
                mbar_wait(opfull_bar, q & 1u);                       // B operand holds h_{t-1}
                tc_fence_after();
#pragma unroll 2
                for (int j = 0; j < KSTEPS; ++j) {
                    const uint32_t boff = (uint32_t)((j >> 2) * (NSEQ * 128) + (j & 3) * 32);
                    uint64_t xb[NS];
                    uint32_t wa[NS];
#pragma unroll
                    for (int p = 0; p < NS; ++p) {
                        xb[p] = make_b_desc(smem_base + (uint32_t)(p * PART_BYTES) + boff);
                        wa[p] = tmem_base + (uint32_t)(p * W_PART_COLS + j * 8);
                    }
                    umma_ts(d_tmem, wa[0], xb[0], idesc, 1u);
                    umma_ts(d_tmem, wa[0], xb[1], idesc, 1u);
                    umma_ts(d_tmem, wa[1], xb[0], idesc, 1u);
                    umma_ts(d_tmem, wa[1], xb[1], idesc, 1u);
                    umma_ts(d_tmem, wa[0], xb[2], idesc, 1u);
                    umma_ts(d_tmem, wa[2], xb[0], idesc, 1u);
                }
                umma_commit(accf_bar);
            }
        }
    } else {
        // =========================================================== producers of the B operand (h_{t-1} planes)
        const int pt = tid - PROD_WARP0 * 32;          // 0..255
        uint32_t q = 0;
        for (int64_t it = 0; it < n_items; ++it) {
            for (int step = 0; step < P.T; ++step, ++q) {
                if (q > 0) {
                    // every CTA of the group has published step q-1 (and is therefore done reading the buffer it is
                    // about to overwrite two steps later)
                    if (pt == 0) {
                        const unsigned want = (unsigned)SLICES * q;
                        for (uint32_t spin = 0; ld_acquire(counter) < want; ++spin) {
                            if (spin > 64) __nanosleep(32);
                            if (spin > (1u << 25)) __trap();
                        }
                    }
                    named_bar(4, PROD_WARPS * 32);
                }
                if (step == 0) {
                    const uint2 z = make_uint2(0u, 0u);
                    for (int i = pt; i < NSEQ * (LH / 4); i += PROD_WARPS * 32) {
                        const uint32_t off = b_offset(i >> 6, (i & 63) * 4);
#pragma unroll
                        for (int p = 0; p < NS; ++p) *reinterpret_cast<uint2*>(smem_gen + p * PART_BYTES + off) = z;
                    }
                } else {
                    const float* src = hx + (size_t)((q - 1) & 1u) * NSEQ * LH;
                    constexpr int PB = 8;                                   // float4 loads in flight per thread
                    for (int i0 = pt; i0 < NSEQ * (LH / 4); i0 += PB * PROD_WARPS * 32) {
                        float4 v[PB];
#pragma unroll
                        for (int u = 0; u < PB; ++u) v[u] = ld_cg4(src + (size_t)(i0 + u * PROD_WARPS * 32) * 4);
#pragma unroll
                        for (int u = 0; u < PB; ++u) {
                            const int i = i0 + u * PROD_WARPS * 32;
                            uint32_t lo[NS], hi[NS];
                            split_pair(v[u].x, v[u].y, lo);
                            split_pair(v[u].z, v[u].w, hi);
                            const uint32_t off = b_offset(i >> 6, (i & 63) * 4);
#pragma unroll
                            for (int p = 0; p < NS; ++p)
                                *reinterpret_cast<uint2*>(smem_gen + p * PART_BYTES + off) = make_uint2(lo[p], hi[p]);
                        }
                    }
                }
                fence_proxy_async();
                mbar_arrive(opfull_bar);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------- packing / helper kernels
// Wt [K_src, 1024] (k-major, gate g = w*256 + unit) -> slab weights [2? no: one direction][8 slices][K_dst][128]:
// dst[s][k][w*32+u] = (k_lo <= k < k_lo + K_dst_real) ? Wt[(k_lo + k) * 1024 + w*256 + s*32 + u] : 0
__global__ void pack_slices_kernel(const float* __restrict__ Wt, int k_lo, int k_real, int K_dst, float* __restrict__ dst) {
    const int total = SLICES * K_dst * UM;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int j = i % UM, k = (i / UM) % K_dst, s = i / (UM * K_dst);
        const int g = (j >> 5) * LH + s * 32 + (j & 31);
        dst[i] = k < k_real ? Wt[(size_t)(k_lo + k) * (4 * LH) + g] : 0.f;
    }
}
__global__ void pack_bias_kernel(const float* __restrict__ bias, float* __restrict__ dst) {      // [8][128]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < SLICES * UM) {
        const int j = i % UM, s = i / UM;
        dst[i] = bias[(j >> 5) * LH + s * 32 + (j & 31)];
    }
}
// seq [R, D] -> xpad [R, 32] (zero padded)
__global__ void pad_rows_kernel(const float* __restrict__ seq, int D, int64_t R, float* __restrict__ xpad) {
    const int64_t total = R * 32;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i & 31);
        xpad[i] = k < D ? seq[(i >> 5) * D + k] : 0.f;
    }
}
// identity-shift graph: row v has one in-edge from row R + v (weight 1); rp0 = all-zero rowptr (no edges)
__global__ void shift_graph_kernel(int64_t R, int32_t* __restrict__ rp1, int32_t* __restrict__ col, float* __restrict__ ew,
                                   int32_t* __restrict__ rp0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= R; i += (int64_t)gridDim.x * blockDim.x) {
        rp1[i] = (int32_t)i; rp0[i] = 0;
        if (i < R) { col[i] = (int32_t)(R + i); ew[i] = 1.0f; }
    }
}

struct Ws {
    size_t gxa, gxb, slabs, xpad, wih, whh, bias, zero_bias, rp1, rp0, col, ew, hx, counter, total;
};
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
Ws carve(int64_t B, int T, int groups) {
    const size_t R = (size_t)B * T;
    Ws w{};
    size_t o = 0;
    w.gxa = o; o += up256(2 * SLICES * R * UM * 4);
    w.gxb = o; o += up256(2 * SLICES * R * UM * 4);
    w.slabs = o; o += up256(4 * R * UM * 4);
    w.xpad = o; o += up256(R * 32 * 4);
    w.wih = o; o += up256((size_t)SLICES * 256 * UM * 4);           // one (dir, K half) at a time
    w.whh = o; o += up256((size_t)2 * SLICES * LH * UM * 4);
    w.bias = o; o += up256((size_t)SLICES * UM * 4);
    w.zero_bias = o; o += up256((size_t)UM * 4);
    w.rp1 = o; o += up256((R + 1) * 4);
    w.rp0 = o; o += up256((R + 1) * 4);
    w.col = o; o += up256(R * 4);
    w.ew = o; o += up256(R * 4);
    w.hx = o; o += up256((size_t)groups * 2 * NSEQ * LH * 4);
    w.counter = o; o += 256;
    w.total = o;
    return w;
}
int groups_for_device() {
    int g = sm_count() / SLICES;
    g &= ~1;                                         // groups alternate directions
    return g < 2 ? 0 : g;
}

}  // namespace

bool lstm_umma_enabled() {
    const char* e = getenv("NERRF_LSTM_ALGO");
    return e && (e[0] == 'u' || e[0] == 'U');         // "umma"; default stays the FFMA kernel until validated on hardware
}

size_t lstm_umma_workspace_bytes(int64_t B, int T) { return carve(B, T, groups_for_device() > 0 ? groups_for_device() : 18).total + 256; }

// Runs every layer; fills hfin [B, 2H] with the top layer's final states.  Weights in the layout of nerrf_lstm_forward.
int lstm_layers_umma(const float* seq, const int32_t* len, int64_t B, int T, int D_in, int num_layers, const float* const* Wih_t,
                     const float* const* Whh_t, const float* const* bias, float* hfin, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
    const int groups = groups_for_device();
    NERRF_REQUIRE(groups >= 2, "tensor-core LSTM needs at least 16 SMs");
    NERRF_REQUIRE(D_in <= 32, "tensor-core LSTM: D_in <= 32 (got %d)", D_in);
    NERRF_REQUIRE((int64_t)B * T * 2 < ((int64_t)1 << 31), "tensor-core LSTM: B*T too large for 32-bit row ids");
    const Ws w = carve(B, T, groups);
    NERRF_REQUIRE(workspace_bytes >= w.total, "lstm workspace too small for the tensor-core path");
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const int64_t R = B * (int64_t)T;
    float* gxa = (float*)(ws + w.gxa); float* gxb = (float*)(ws + w.gxb); float* slabs = (float*)(ws + w.slabs);
    float* xpad = (float*)(ws + w.xpad); float* wih = (float*)(ws + w.wih); float* whh = (float*)(ws + w.whh);
    float* bp = (float*)(ws + w.bias); float* zb = (float*)(ws + w.zero_bias);
    int32_t* rp1 = (int32_t*)(ws + w.rp1); int32_t* rp0 = (int32_t*)(ws + w.rp0); int32_t* col = (int32_t*)(ws + w.col);
    float* ew = (float*)(ws + w.ew); float* hx = (float*)(ws + w.hx); unsigned* counter = (unsigned*)(ws + w.counter);
    const int blk = sm_count() * 4;
    NERRF_CHECK_CUDA(cudaMemsetAsync(zb, 0, UM * 4, st));
    shift_graph_kernel<<<blk, 256, 0, st>>>(R, rp1, col, ew, rp0);
    pad_rows_kernel<<<blk, 256, 0, st>>>(seq, D_in, R, xpad);
    static bool attr_set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set_dev[dev_ & 63]) {
        NERRF_CHECK_CUDA(cudaFuncSetAttribute(lstm_rec_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        attr_set_dev[dev_ & 63] = true;
    }
    for (int l = 0; l < num_layers; ++l) {
        // ---- input projection for both directions: 16 (layer 1) / 32 (deeper layers) dense-GEMM launches
        for (int d = 0; d < 2; ++d) {
            const float* Wt = Wih_t[2 * l + d];
            pack_bias_kernel<<<(SLICES * UM + 255) / 256, 256, 0, st>>>(bias[2 * l + d], bp);
            const int halves = l == 0 ? 1 : 2;
            for (int kh = 0; kh < halves; ++kh) {
                if (l == 0) pack_slices_kernel<<<blk, 256, 0, st>>>(Wt, 0, D_in, 64, wih);            // K = [x(32) || m(32)=0]
                else pack_slices_kernel<<<blk, 256, 0, st>>>(Wt, kh * 256, 256, 256, wih);
                for (int s = 0; s < SLICES; ++s) {
                    float* dst = (kh == 0 ? gxa : gxb) + (size_t)(d * SLICES + s) * R * UM;
                    int rc;
                    if (l == 0)
                        rc = sage_layer_umma(xpad, rp0, 0, col, ew, wih + (size_t)s * 64 * UM, bp + s * UM, dst, R, 0, R, 32, 0, 3,
                                             nullptr, 0.f, nullptr, nullptr, 0, false, nullptr, 0, nullptr, st);
                    else
                        rc = sage_layer_umma(slabs + (size_t)kh * 2 * R * UM, rp1, 0, col, ew, wih + (size_t)s * 256 * UM,
                                             kh == 0 ? bp + s * UM : zb, dst, 2 * R, 0, R, 128, 0, 3, nullptr, 0.f, nullptr, nullptr,
                                             0, false, nullptr, 0, nullptr, st);
                    if (rc) return rc;
                }
            }
            pack_slices_kernel<<<blk, 256, 0, st>>>(Whh_t[2 * l + d], 0, LH, LH, whh + (size_t)d * SLICES * LH * UM);
        }
        // ---- recurrence
        NERRF_CHECK_CUDA(cudaMemsetAsync(counter, 0, 256, st));
        RecArgs a{};
        a.gx_a = gxa; a.gx_b = l == 0 ? nullptr : gxb; a.whh = whh; a.len = len;
        a.out_slabs = (l == num_layers - 1) ? nullptr : slabs; a.hfin = hfin; a.hx = hx; a.counter = counter;
        a.B = B; a.T = T; a.groups = groups;
        void* args[] = {&a};
        NERRF_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_rec_umma_kernel, dim3(groups * SLICES), dim3(THREADS), args,
                                                     SMEM, st));
    }
    return launch_status("lstm_rec_umma_kernel");
}

}  // namespace nerrf
