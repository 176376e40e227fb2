// BiLSTM on 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only -- lstm.forward.
//
// One LSTM layer-direction step is   gates[4H, seqs] = G_x[:, t] + W_hh[4H, H] . h_{t-1}[H, seqs]   followed by the cell
// update.  Everything that multiplies a weight matrix runs on tcgen05 with the weight slice RESIDENT IN TENSOR MEMORY
// (A operand, fp32 split into three bf16 terms, six products per K step: fp32-equivalent accuracy, same scheme as
// sage_umma.cu) and the activations as the B operand, bulk-copied (cp.async.bulk, byte-counting mbarrier) from a
// global-memory OPERAND IMAGE -- the exact byte layout the UMMA reads (K-major, SWIZZLE_128B, three bf16 planes):
//
//   image of a time-major activation X[t][b][K]:   tile (t, bt = b / 64, khalf):  [plane 3][kblock KB][row 64][128 B]
//
//   * lstm_image_kernel     layer-1 input seq [B,T,D] -> image (K padded to 64, one K block)
//   * lstm_proj_kernel      input projection G_x = X W_ih^T (+ b) for all (t, b): persistent CTAs, each keeps ONE
//                           (direction, gate slice) of W_ih^T (one K half of 256) in TMEM and streams 64-row tiles of the
//                           image through a two-stage bulk-copy pipeline; no conversion work, the tensor pipe is the limit
//   * lstm_rec_kernel       the recurrence.  W_hh (1024 x 256) is cut into 8 slices of 128 gate rows (4 gates x 32 units);
//                           8 CTAs (one per slice) form a GROUP that advances tiles of 64 sequences of one direction.
//                           Per step every CTA computes gates^T[128, 64] (+)= W_slice . h_{t-1}^T with the accumulator
//                           pre-loaded with G_x[:, t]; the epilogue applies the activations, updates c / h in registers,
//                           splits the new h into its three bf16 terms and stores them INTO THE LAYER'S OUTPUT IMAGE
//                           (tile (t, bt, khalf = direction)) -- which is at once the next layer's projection operand and
//                           the h_{t} exchange between the 8 CTAs: after one release-increment of the group's counter every
//                           CTA bulk-copies the whole 96 KB tile back as its next B operand.
//                           The step is a latency chain (MMA -> epilogue -> publish -> poll -> 96 KB copy -> MMA), so every
//                           group works on TWO tiles at once (two operand buffers, two accumulators, two epilogue
//                           warpgroups): while one tile waits for its exchange the other one's MMAs run.
//
// Groups are independent (no grid barrier): group g owns direction g & 1.  The 8 CTAs of a group must be co-resident:
// cooperative launch, one CTA per SM, grid = 8 * G <= #SMs.  Every wait is bounded and ends in a trap, never a hang.
//
// History (profiles/r02_lstm.md): v0 FFMA kernel 51.6 ms for B=4096 / T=100; v1 (fp32 exchange, every CTA re-splits h,
// projections through the GraphSAGE-T kernel as a dense GEMM) 31.7 ms; v2 (bf16x3 image exchange) 30.3 ms -- the
// recurrence was a 19 us latency chain per step with the tensor pipe idle 85 % of the time, hence the two-tile interleave
// and the dedicated projection kernel of this version.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace nerrf {

namespace {

constexpr int LH = 256;                 // hidden size
constexpr int NSEQ = 64;                // sequences per tile = UMMA N
constexpr int SLICES = 8;               // CTAs per group
constexpr int UM = 128;                 // gate rows per slice = UMMA M
constexpr int NS = 3;                   // bf16 terms per fp32 value
constexpr int KB_H = LH / 64;           // 128-byte K blocks of a 256-wide operand
constexpr int KSTEPS_H = LH / 16;
constexpr int PLANE_BYTES_H = KB_H * NSEQ * 128;     // 32 KB: one bf16 plane of a [64 rows, 256] tile
constexpr int TILE_BYTES_H = NS * PLANE_BYTES_H;     // 96 KB
constexpr int W_PART_COLS = LH / 2;                  // TMEM columns of one W plane (K = 256)
constexpr int ACC_COL0 = NS * W_PART_COLS;           // 384
constexpr int TMEM_COLS = 512;

// ---------------------------------------------------------------- PTX wrappers (same forms as sage_umma.cu)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
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
// every wait in this file is bounded: a protocol bug must end in a trap (launch error), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin) {
        if (spin > (1u << 16)) __nanosleep(64);
        if (spin > (1u << 26)) __trap();
    }
}
// global -> shared bulk async copy (async proxy, SASS UBLKCP); completes `bytes` transactions on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
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
// byte offset of ONE bf16 element (row r of the tile, K index k inside the K block range) inside a plane: K-major
// SWIZZLE_128B, 128-byte rows of one 64-wide K block, 16-byte chunks XOR-swizzled by the row (what the UMMA reads)
__device__ __forceinline__ uint32_t b_elem_offset(int r, int k) {
    const int kb = k >> 6, col = k & 63;
    return (uint32_t)(kb * (NSEQ * 128) + r * 128 + ((((col >> 3) ^ (r & 7)) << 4) | ((col & 7) << 1)));
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// activations: |abs error| ~ 2e-7 (ex2.approx + rcp), far inside the 1e-4 spec bound; tanh(x) = 2 sigmoid(2x) - 1
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.0f, sigmoid_fast(2.0f * x), -1.0f); }

// W^T slice [K, 128] (fp32, k-major) -> TMEM planes: thread = gate row f (lane of TMEM), 2 bf16 per 32-bit column.
// K / 2 columns per plane; plane p starts at column p * plane_stride.
__device__ __forceinline__ void load_w_to_tmem(const float* __restrict__ W, int K, uint32_t tmem_base, int plane_stride, int warp, int lane) {
    const int f = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 16) {
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
        for (int p = 0; p < NS; ++p) tmem_st8(lane_addr + (uint32_t)(p * plane_stride + (k0 >> 1)), parts[p]);
    }
    tmem_wait_st();
}

// the six split products of one 64-row tile: D[128, 64] (+)= W[128, K] . X[64, K]^T, W planes in TMEM, X planes in smem
template <int KSTEPS>
__device__ __forceinline__ void issue_tile_mmas(uint32_t d_tmem, uint32_t w_tmem, int plane_stride, uint32_t x_smem, uint32_t plane_bytes,
                                                bool accumulate_first) {
    constexpr uint32_t idesc = make_idesc();
#pragma unroll 2
    for (int j = 0; j < KSTEPS; ++j) {
        const uint32_t boff = (uint32_t)((j >> 2) * (NSEQ * 128) + (j & 3) * 32);
        uint64_t xb[NS];
        uint32_t wa[NS];
#pragma unroll
        for (int p = 0; p < NS; ++p) {
            xb[p] = make_b_desc(x_smem + (uint32_t)p * plane_bytes + boff);
            wa[p] = w_tmem + (uint32_t)(p * plane_stride + j * 8);
        }
        umma_ts(d_tmem, wa[0], xb[0], idesc, (j > 0 || accumulate_first) ? 1u : 0u);
        umma_ts(d_tmem, wa[0], xb[1], idesc, 1u);
        umma_ts(d_tmem, wa[1], xb[0], idesc, 1u);
        umma_ts(d_tmem, wa[1], xb[1], idesc, 1u);
        umma_ts(d_tmem, wa[0], xb[2], idesc, 1u);
        umma_ts(d_tmem, wa[2], xb[0], idesc, 1u);
    }
}

// ================================================================================================ layer-1 input image
// seq [B, T, D] fp32 -> image tiles (t, bt): [plane 3][1 K block][64 rows][128 B], K padded to 64, rows past B zero
__global__ void lstm_image_kernel(const float* __restrict__ seq, int D, int64_t B, int T, int64_t n_bt, unsigned char* __restrict__ img) {
    constexpr int TILE = NS * NSEQ * 128;                            // 24 KB
    const int64_t total = (int64_t)T * n_bt * NSEQ * 32;             // one thread per (t, bt, row, k pair)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int kp = (int)(i & 31);
        const int r = (int)((i >> 5) % NSEQ);
        const int64_t tile = i / (32 * NSEQ);
        const int64_t t = tile / n_bt, bt = tile % n_bt;
        const int64_t b = bt * NSEQ + r;
        float a = 0.f, c = 0.f;
        if (b < B) {
            const float* s = seq + ((size_t)b * T + t) * D;
            if (2 * kp < D) a = s[2 * kp];
            if (2 * kp + 1 < D) c = s[2 * kp + 1];
        }
        uint32_t pp[NS];
        split_pair(a, c, pp);
        unsigned char* dst = img + (size_t)tile * TILE + (r * 128 + ((((kp >> 2) ^ (r & 7)) << 4) | ((kp & 3) << 2)));
#pragma unroll
        for (int p = 0; p < NS; ++p) *reinterpret_cast<uint32_t*>(dst + p * (NSEQ * 128)) = pp[p];
    }
}

// ================================================================================================ input projection
// G[dir][slice][row = t*Bp + b][128] = X_tile . W_ih^T slice (+ bias), for one K range of KB 64-wide blocks.
//   img       activation image; tile (t, bt) = the K range starts at img + tile * tile_stride + k_off
//   wih       [16 combos = dir*8 + slice][K = 64*KB][128] fp32 (packed by pack_slices_kernel), bias [16][128] or nullptr
struct ProjArgs {
    const unsigned char* img;
    size_t tile_stride;        // bytes between consecutive (t, bt) tiles
    size_t k_off;              // byte offset of this K range inside a tile
    const float* wih;
    const float* bias;
    float* gx;                 // [16][T*Bp][128]
    int64_t n_tiles;           // T * n_bt
    int64_t rows_total;        // T * Bp
    int ctas_per_combo;
};

template <int KB>
__global__ void __launch_bounds__(192, 1) lstm_proj_kernel(ProjArgs P) {
    constexpr int K = 64 * KB, KSTEPS = K / 16;
    constexpr int PLANE = KB * NSEQ * 128, TILE = NS * PLANE;       // bytes of one (tile, K range)
    constexpr int WCOLS = K / 2;                                    // TMEM columns per W plane
    constexpr int ACC0 = NS * WCOLS;
    constexpr int STAGES = 2;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + STAGES * TILE;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto accf_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto acce_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;          // warps 0-3 epilogue, 4 MMA, 5 loader
    const int combo = blockIdx.x / P.ctas_per_combo, j0 = blockIdx.x % P.ctas_per_combo;
    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(accf_bar(a), 1); mbar_init(acce_bar(a), 128); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    if (warp < 4) load_w_to_tmem(P.wih + (size_t)combo * K * UM, K, tmem_base, WCOLS, warp, lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int64_t n_mine = j0 < P.n_tiles ? (P.n_tiles - j0 + P.ctas_per_combo - 1) / P.ctas_per_combo : 0;
    if (warp == 5) {
        if (lane == 0) {
            for (int64_t i = 0; i < n_mine; ++i) {
                const int s = (int)(i % STAGES);
                mbar_wait(empty_bar(s), ((uint32_t)(i / STAGES) & 1u) ^ 1u);
                const unsigned char* src = P.img + (size_t)(j0 + i * P.ctas_per_combo) * P.tile_stride + P.k_off;
                mbar_arrive_expect_tx(full_bar(s), (uint32_t)TILE);
                constexpr int CH = 8192;                                 // divides both tile sizes (24 KB, 96 KB)
                static_assert(TILE % CH == 0, "tile must be a whole number of bulk copies");
#pragma unroll 1
                for (int c = 0; c < TILE / CH; ++c)
                    bulk_g2s(smem_base + (uint32_t)(s * TILE + c * CH), src + (size_t)c * CH, (uint32_t)CH, full_bar(s));
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            for (int64_t i = 0; i < n_mine; ++i) {
                const int s = (int)(i % STAGES), a = (int)(i & 1);
                mbar_wait(acce_bar(a), ((uint32_t)(i >> 1) & 1u) ^ 1u);
                mbar_wait(full_bar(s), (uint32_t)(i / STAGES) & 1u);
                tc_fence_after();
                issue_tile_mmas<KSTEPS>(tmem_base + (uint32_t)(ACC0 + a * NSEQ), tmem_base, WCOLS, smem_base + (uint32_t)(s * TILE),
                                        (uint32_t)PLANE, false);
                umma_commit(empty_bar(s));
                umma_commit(accf_bar(a));
            }
        }
    } else {
        const int gcol = warp * 32 + lane;
        const float my_bias = P.bias ? __ldg(P.bias + (size_t)combo * UM + gcol) : 0.f;
        float* g = P.gx + (size_t)combo * P.rows_total * UM + gcol;
        for (int64_t i = 0; i < n_mine; ++i) {
            const int a = (int)(i & 1);
            mbar_wait(accf_bar(a), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ACC0 + a * NSEQ);
            const int64_t row0 = (j0 + i * P.ctas_per_combo) * NSEQ;      // rows of the padded time-major matrix
            uint32_t v[32], w[32];
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, w);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(acce_bar(a));
#pragma unroll
            for (int r = 0; r < 32; ++r) g[(size_t)(row0 + r) * UM] = __uint_as_float(v[r]) + my_bias;
#pragma unroll
            for (int r = 0; r < 32; ++r) g[(size_t)(row0 + 32 + r) * UM] = __uint_as_float(w[r]) + my_bias;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ================================================================================================ recurrence
// The epilogue is instruction-issue bound (2 warps per scheduler ran at IPC ~0.25 each: 20 k cycles per step,
// profiles/r02_lstm.md), so every tile slot gets EIGHT epilogue warps: two per TMEM lane quarter, each owning 32 of the
// tile's 64 sequences.
constexpr int REC_EPI_WARPS = 16;                    // warps 0-7: tile slot 0, warps 8-15: tile slot 1
constexpr int REC_MMA_WARP = 16;
constexpr int REC_INIT_WARP0 = 17;                   // warps 17-20: slot 0, warps 21-24: slot 1 (accumulator init + exchange)
constexpr int REC_THREADS = (REC_INIT_WARP0 + 8) * 32;               // 800
constexpr int GATE_BYTES = 2 * 4 * 16 * 32 * 4;      // 16 KB per slot: [half][gate][seq in chunk of 16][unit]
constexpr size_t REC_SMEM = (size_t)2 * TILE_BYTES_H + 2 * GATE_BYTES + 2 * NSEQ * 4 + 1024 /*align*/ + 128 /*barriers*/;
static_assert(REC_SMEM <= 227 * 1024, "shared memory budget");
static_assert(ACC_COL0 + 2 * NSEQ <= TMEM_COLS, "TMEM budget");

struct RecArgs {
    const float* gx_a;        // [2 dirs * 8 slices][Rp = T*Bp rows][128]: first K range of the input projection + bias
    const float* gx_b;        // second K half, or nullptr
    const float* whh;         // [2 dirs * 8 slices][256 (k)][128 (gate row in slice)]
    const int32_t* len;       // [B]
    unsigned char* out_img;   // output image: tile (t, bt) at (t*n_bt + bt) * 2*TILE_BYTES_H, K half = direction
    float* hfin;              // [B][512]
    unsigned* counter;        // [groups][2 slots], zero at launch
    int64_t B, Bp, n_bt;
    int T;
    int groups;
};

__global__ void __launch_bounds__(REC_THREADS, 1) lstm_rec_kernel(RecArgs P) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    float* gate_s = reinterpret_cast<float*>(smem_gen + 2 * TILE_BYTES_H);                       // [2 slots][4][32][32]
    int* len_s = reinterpret_cast<int*>(smem_gen + 2 * TILE_BYTES_H + 2 * GATE_BYTES);           // [2 slots][64]
    const uint32_t bar_base = smem_base + 2 * TILE_BYTES_H + 2 * GATE_BYTES + 2 * NSEQ * 4;
    auto opfull_bar = [&](int s) { return bar_base + 8u * s; };
    auto accinit_bar = [&](int s) { return bar_base + 16u + 8u * s; };
    auto accf_bar = [&](int s) { return bar_base + 32u + 8u * s; };
    auto drained_bar = [&](int s) { return bar_base + 48u + 8u * s; };
    const uint32_t tmem_slot = bar_base + 64u;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = blockIdx.x / SLICES, slice = blockIdx.x % SLICES;
    const int dir = group & 1;
    const int groups_dir = P.groups >> 1;                          // groups per direction
    const int gd = group >> 1;                                     // this group's index inside its direction
    // tile sequence of (group, slot): bt = gd + (2 * i + slot) * groups_dir, i = 0, 1, ...
    auto n_items = [&](int slot) -> int64_t {
        const int64_t first = gd + (int64_t)slot * groups_dir;
        return first < P.n_bt ? (P.n_bt - first + 2 * groups_dir - 1) / (2 * groups_dir) : 0;
    };
    const int64_t Rp = (int64_t)P.T * P.Bp;

    if (warp == REC_MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(opfull_bar(s), 1);                 // one arrive (+ the bulk copies' transaction bytes)
                mbar_init(accinit_bar(s), 128);              // the slot's accumulator-init warps
                mbar_init(accf_bar(s), 1);
                mbar_init(drained_bar(s), 256);              // the slot's epilogue warps
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    if (warp < 4) load_w_to_tmem(P.whh + (size_t)(dir * SLICES + slice) * LH * UM, LH, tmem_base, W_PART_COLS, warp, lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const float* gxa = P.gx_a + (size_t)(dir * SLICES + slice) * Rp * UM;
    const float* gxb = P.gx_b ? P.gx_b + (size_t)(dir * SLICES + slice) * Rp * UM : nullptr;

    if (warp < REC_EPI_WARPS) {
        // =========================================================== epilogue warps of tile slot `slot`
        const int slot = warp >> 3;
        const int half = (warp >> 2) & 1;           // sequences [32 half, 32 half + 32) of the tile
        const int wq = warp & 3;                    // TMEM lane quarter = gate type (i, f, g, o)
        const int et = wq * 32 + lane;              // thread index inside the (slot, half) warpgroup (0..127)
        const int bar_id = 1 + slot * 2 + half;     // named barrier of this warpgroup
        float* gs = gate_s + slot * (GATE_BYTES / 4) + half * (GATE_BYTES / 8);
        int* ls = len_s + slot * NSEQ;
        const uint32_t acc_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ACC_COL0 + slot * NSEQ + half * 32);
        const int ucol = slice * 32 + lane;         // hidden unit of this thread in the cell-update phase
        unsigned* counter = P.counter + (size_t)group * 2 + slot;
        uint32_t q = 0;                             // global step index of this (group, slot)
        const int64_t items = n_items(slot);
        const bool prof_on = blockIdx.x == 0 && tid == 0;           // diagnostics: cycles of CTA 0's slot-0 epilogue thread 0
        long long pf[4] = {0, 0, 0, 0}, tc = prof_on ? clock64() : 0;
        for (int64_t it = 0; it < items; ++it) {
            const int64_t bt = gd + (2 * it + slot) * groups_dir;
            const int64_t b0 = bt * NSEQ;
            if (et < 32) { const int64_t b = b0 + half * 32 + et; ls[half * 32 + et] = b < P.B ? P.len[b] : 0; }
            named_bar(bar_id, 128);
            float c_reg[8], h_reg[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { c_reg[j] = 0.f; h_reg[j] = 0.f; }
            for (int step = 0; step < P.T; ++step, ++q) {
                const int t = dir ? (P.T - 1 - step) : step;
                unsigned char* tile = P.out_img + ((size_t)t * P.n_bt + bt) * (2 * TILE_BYTES_H) + (size_t)dir * TILE_BYTES_H;
                mbar_wait(accf_bar(slot), q & 1u);
                tc_fence_after();
                if (prof_on) { const long long c_ = clock64(); pf[0] += c_ - tc; tc = c_; }      // waiting for the MMAs
                uint32_t v0[16], v1[16];
                tmem_ld16(acc_addr, v0);
                tmem_ld16(acc_addr + 16, v1);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(drained_bar(slot));                     // this thread has read all of its accumulator columns
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    // phase A: activated gates of this thread's gate row for 16 sequences -> shared [gate][seq][unit]
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float x = __uint_as_float(c == 0 ? v0[j] : v1[j]);
                        gs[(wq * 16 + j) * 32 + lane] = (wq == 2) ? tanh_fast(x) : sigmoid_fast(x);
                    }
                    named_bar(bar_id, 128);
                    // phase B: cell update, thread = (unit lane, sequences wq*4 .. wq*4+3 of the chunk)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = wq * 4 + jj;                 // sequence inside the chunk
                        const int n = half * 32 + c * 16 + j;      // sequence inside the tile
                        const float ig = gs[(0 * 16 + j) * 32 + lane], fg = gs[(1 * 16 + j) * 32 + lane];
                        const float gg = gs[(2 * 16 + j) * 32 + lane], og = gs[(3 * 16 + j) * 32 + lane];
                        float& cr = c_reg[c * 4 + jj];
                        float& hr = h_reg[c * 4 + jj];
                        if (t < ls[n]) {
                            cr = fmaf(fg, cr, ig * gg);
                            hr = og * tanh_fast(cr);
                        }
                        // h (state, kept when masked) -> three bf16 terms -> output image = exchange; even lanes store the
                        // pair.  Masked positions of the layer output are never read by the next layer (its own steps
                        // are masked there), so the kept state is as good as the zero packed sequences would hold.
                        float r = hr;
                        uint32_t term[NS];
#pragma unroll
                        for (int p = 0; p < NS; ++p) {
                            const __nv_bfloat16 hb = __float2bfloat16_rn(r);
                            term[p] = (uint32_t)__bfloat16_as_ushort(hb);
                            r -= __bfloat162float(hb);
                        }
                        unsigned char* dstp = tile + b_elem_offset(n, ucol);
#pragma unroll
                        for (int p = 0; p < NS; ++p) {
                            const uint32_t nb = __shfl_down_sync(0xffffffffu, term[p], 1);
                            if (!(lane & 1)) *reinterpret_cast<uint32_t*>(dstp + p * PLANE_BYTES_H) = term[p] | (nb << 16);
                        }
                    }
                    named_bar(bar_id, 128);                         // the gate buffer is reused by the next chunk
                }
                if (prof_on) { const long long c_ = clock64(); pf[1] += c_ - tc; tc = c_; }      // activations + cell update + stores
                fence_proxy_async_all();                            // generic stores above are read by bulk (async-proxy) copies
                named_bar(5 + slot, 256);                           // both halves of the slot have written their h
                if (half == 0 && et == 0) {
                    __threadfence();
                    atomicAdd(counter, 1u);                          // release: this CTA's h_t slice is published
                }
                if (prof_on) { const long long c_ = clock64(); pf[2] += c_ - tc; tc = c_; }      // fences + publish
            }
            // final state of the tile
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int64_t b = b0 + half * 32 + c * 16 + wq * 4 + jj;
                    if (b < P.B) P.hfin[b * (2 * LH) + dir * LH + ucol] = h_reg[c * 4 + jj];
                }
        }
        if (prof_on) for (int i = 0; i < 3; ++i) P.counter[128 + i] = (unsigned)(pf[i] >> 6);
    } else if (warp == REC_MMA_WARP) {
        // =========================================================== MMA issuer (one thread), slots interleaved
        if (lane == 0) {
            const uint32_t total[2] = {(uint32_t)(n_items(0) * P.T), (uint32_t)(n_items(1) * P.T)};
            uint32_t q[2] = {0, 0};
            uint32_t idle = 0;
            while (q[0] < total[0] || q[1] < total[1]) {
                bool issued = false;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    // whichever tile is ready goes next: accumulator holds G_x[:, t] AND the B operand holds h_{t-1}
                    if (q[s] >= total[s] || !mbar_try(accinit_bar(s), q[s] & 1u) || !mbar_try(opfull_bar(s), q[s] & 1u)) continue;
                    tc_fence_after();
                    issue_tile_mmas<KSTEPS_H>(tmem_base + (uint32_t)(ACC_COL0 + s * NSEQ), tmem_base, W_PART_COLS,
                                              smem_base + (uint32_t)(s * TILE_BYTES_H), (uint32_t)PLANE_BYTES_H, true);
                    umma_commit(accf_bar(s));
                    ++q[s];
                    issued = true;
                }
                if (issued) idle = 0;
                else if (++idle > (1u << 24)) __trap();                  // bounded: a protocol bug ends in a launch error
            }
        }
    } else {
        // =========================================================== accumulator-init warps + the slot's exchange thread
        const int slot = (warp - REC_INIT_WARP0) >> 2;
        const int wq = warp & 3;                       // TMEM lane quarter this warp may touch (warp % 4)
        const int pt = ((warp - REC_INIT_WARP0) & 3) * 32 + lane;     // 0..127 inside the slot's init warpgroup
        const int gcol = wq * 32 + lane;               // gate row inside the slice = column of the G_x slab
        const uint32_t acc_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ACC_COL0 + slot * NSEQ);
        unsigned* counter = P.counter + (size_t)group * 2 + slot;
        unsigned char* op_gen = smem_gen + (size_t)slot * TILE_BYTES_H;
        const uint32_t op_u32 = smem_base + (uint32_t)(slot * TILE_BYTES_H);
        uint32_t q = 0;
        const int64_t items = n_items(slot);
        const bool prof_on = blockIdx.x == 0 && slot == 0 && pt == 0;
        long long pf[4] = {0, 0, 0, 0}, tc = prof_on ? clock64() : 0;
        for (int64_t it = 0; it < items; ++it) {
            const int64_t bt = gd + (2 * it + slot) * groups_dir;
            const int64_t b0 = bt * NSEQ;
            for (int step = 0; step < P.T; ++step, ++q) {
                const int t = dir ? (P.T - 1 - step) : step;
                if (prof_on) tc = clock64();
                // ---- accumulator <- G_x[:, t] (rows of the padded time-major matrix: always in range)
                const float* ra = gxa + (size_t)((int64_t)t * P.Bp + b0) * UM + gcol;
                const float* rb = gxb ? gxb + (size_t)((int64_t)t * P.Bp + b0) * UM + gcol : nullptr;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float ga[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) ga[j] = __ldg(ra + (size_t)(c * 16 + j) * UM);
                    if (rb) {
                        float gb[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) gb[j] = __ldg(rb + (size_t)(c * 16 + j) * UM);
#pragma unroll
                        for (int j = 0; j < 16; ++j) ga[j] += gb[j];
                    }
                    if (c == 0 && q > 0) {                           // the epilogue of step q-1 has drained the accumulator
                        if (prof_on) { const long long c_ = clock64(); pf[0] += c_ - tc; tc = c_; }   // G_x loads issued + landed
                        mbar_wait(drained_bar(slot), (q - 1) & 1u);
                        tc_fence_after();
                        if (prof_on) { const long long c_ = clock64(); pf[1] += c_ - tc; tc = c_; }   // waiting for the drain
                    }
                    uint32_t v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(ga[j]);
                    tmem_st16(acc_addr + (uint32_t)(c * 16), v);
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(accinit_bar(slot));

                // ---- B operand <- h_{t-1}
                if (step == 0) {
                    // (the MMAs of the previous tile's last step have retired: its epilogue ran, and this warpgroup waited
                    // for that epilogue's drained barrier above)
                    named_bar(7 + slot, 128);
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                    for (int i = pt; i < TILE_BYTES_H / 16; i += 128) reinterpret_cast<uint4*>(op_gen)[i] = z;
                    fence_proxy_async();
                    named_bar(7 + slot, 128);
                    if (pt == 0) mbar_arrive(opfull_bar(slot));
                } else if (pt == 0) {
                    // every CTA of the group has published step q-1 of this slot (so this CTA's MMA of step q-1, which
                    // read the operand buffer, has retired as well)
                    const unsigned want = (unsigned)SLICES * q;
                    if (prof_on) { const long long c_ = clock64(); pf[2] += c_ - tc; tc = c_; }       // second chunk + tmem st
                    for (uint32_t spin = 0; ld_acquire(counter) < want; ++spin) {
                        if (spin > 64) __nanosleep(32);
                        if (spin > (1u << 25)) __trap();
                    }
                    if (prof_on) { const long long c_ = clock64(); pf[3] += c_ - tc; tc = c_; }       // polling the group's counter
                    fence_proxy_async_all();                         // acquired generic writes -> async-proxy reads
                    const int tp = dir ? t + 1 : t - 1;              // the step that produced h_{t-1}
                    const unsigned char* src = P.out_img + ((size_t)tp * P.n_bt + bt) * (2 * TILE_BYTES_H) + (size_t)dir * TILE_BYTES_H;
                    mbar_arrive_expect_tx(opfull_bar(slot), (uint32_t)TILE_BYTES_H);
#pragma unroll 1
                    for (int i = 0; i < TILE_BYTES_H / 16384; ++i)
                        bulk_g2s(op_u32 + (uint32_t)(i * 16384), src + (size_t)i * 16384, 16384u, opfull_bar(slot));
                }
            }
        }
        if (prof_on) for (int i = 0; i < 4; ++i) P.counter[136 + i] = (unsigned)(pf[i] >> 6);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == REC_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------- packing kernels
// Wt [K_src, 1024] (k-major, gate g = w*256 + unit) -> [8 slices][K_dst][128]:
// dst[s][k][w*32+u] = (k < k_real) ? Wt[(k_lo + k) * 1024 + w*256 + s*32 + u] : 0
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

struct Ws {
    size_t gxa, gxb, img0, img_a, img_b, wih, whh, bias, counter, total;
};
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
Ws carve(int64_t B, int T, int num_layers) {
    const int64_t n_bt = (B + NSEQ - 1) / NSEQ;
    const size_t Rp = (size_t)T * n_bt * NSEQ;
    const size_t tiles = (size_t)T * n_bt;
    Ws w{};
    size_t o = 0;
    w.gxa = o; o += up256(2 * SLICES * Rp * UM * 4);
    w.gxb = o; o += up256(num_layers > 1 ? 2 * SLICES * Rp * UM * 4 : 0);
    w.img0 = o; o += up256(tiles * NS * NSEQ * 128);                 // layer-1 input image (K = 64)
    w.img_a = o; o += up256(tiles * 2 * TILE_BYTES_H);               // layer outputs, ping-pong
    w.img_b = o; o += up256(num_layers > 1 ? tiles * 2 * TILE_BYTES_H : 0);
    w.wih = o; o += up256((size_t)2 * SLICES * LH * UM * 4);         // one K range of both directions
    w.whh = o; o += up256((size_t)2 * SLICES * LH * UM * 4);
    w.bias = o; o += up256((size_t)2 * SLICES * UM * 4);
    w.counter = o; o += 1024;
    w.total = o;
    return w;
}
int groups_for_device() {
    int g = sm_count() / SLICES;
    g &= ~1;                                         // groups alternate directions
    return g < 2 ? 0 : g;
}

template <int KB>
int launch_proj(const ProjArgs& a, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * NS * KB * NSEQ * 128 + 1024 + 128;
    static bool attr_set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set_dev[dev_ & 63]) {
        NERRF_CHECK_CUDA(cudaFuncSetAttribute(lstm_proj_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set_dev[dev_ & 63] = true;
    }
    lstm_proj_kernel<KB><<<16 * a.ctas_per_combo, 192, smem, st>>>(a);
    return launch_status("lstm_proj_kernel");
}

}  // namespace

// NERRF_LSTM_ALGO=ffma selects the fp32 CUDA-core kernel of lstm.cu (kept as an independent cross-check of this path)
bool lstm_umma_enabled() {
    const char* e = getenv("NERRF_LSTM_ALGO");
    return !(e && (e[0] == 'f' || e[0] == 'F'));
}

size_t lstm_umma_workspace_bytes(int64_t B, int T, int num_layers) { return carve(B, T, num_layers).total + 256; }

// Runs every layer; fills hfin [B, 2H] with the top layer's final states.  Weights in the layout of nerrf_lstm_forward.
int lstm_layers_umma(const float* seq, const int32_t* len, int64_t B, int T, int D_in, int num_layers, const float* const* Wih_t,
                     const float* const* Whh_t, const float* const* bias, float* hfin, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
    const int groups = groups_for_device();
    NERRF_REQUIRE(groups >= 2, "tensor-core LSTM needs at least 16 SMs");
    NERRF_REQUIRE(D_in <= 64, "tensor-core LSTM: D_in <= 64 (got %d)", D_in);
    const Ws w = carve(B, T, num_layers);
    NERRF_REQUIRE(workspace_bytes >= w.total, "lstm workspace too small for the tensor-core path");
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const int64_t n_bt = (B + NSEQ - 1) / NSEQ, Bp = n_bt * NSEQ;
    const int64_t n_tiles = (int64_t)T * n_bt;
    float* gxa = (float*)(ws + w.gxa); float* gxb = (float*)(ws + w.gxb);
    unsigned char* img0 = (unsigned char*)(ws + w.img0);
    unsigned char* img[2] = {(unsigned char*)(ws + w.img_a), (unsigned char*)(ws + w.img_b)};
    float* wih = (float*)(ws + w.wih); float* whh = (float*)(ws + w.whh); float* bp = (float*)(ws + w.bias);
    unsigned* counter = (unsigned*)(ws + w.counter);
    const int blk = sm_count() * 4;
    const int ctas_per_combo = sm_count() / 16 > 0 ? sm_count() / 16 : 1;
    lstm_image_kernel<<<blk, 256, 0, st>>>(seq, D_in, B, T, n_bt, img0);
    static bool attr_set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set_dev[dev_ & 63]) {
        NERRF_CHECK_CUDA(cudaFuncSetAttribute(lstm_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REC_SMEM));
        attr_set_dev[dev_ & 63] = true;
    }
    for (int l = 0; l < num_layers; ++l) {
        // ---- input projection: one launch per K range, 16 (direction, slice) combos each
        const unsigned char* in_img = l == 0 ? img0 : img[(l - 1) & 1];
        for (int d = 0; d < 2; ++d) pack_bias_kernel<<<(SLICES * UM + 255) / 256, 256, 0, st>>>(bias[2 * l + d], bp + (size_t)d * SLICES * UM);
        const int ranges = l == 0 ? 1 : 2;
        for (int kr = 0; kr < ranges; ++kr) {
            const int Kdst = l == 0 ? 64 : LH;
            for (int d = 0; d < 2; ++d) {
                if (l == 0) pack_slices_kernel<<<blk, 256, 0, st>>>(Wih_t[2 * l + d], 0, D_in, 64, wih + (size_t)d * SLICES * Kdst * UM);
                else pack_slices_kernel<<<blk, 256, 0, st>>>(Wih_t[2 * l + d], kr * LH, LH, LH, wih + (size_t)d * SLICES * Kdst * UM);
            }
            ProjArgs a{};
            a.img = in_img;
            a.tile_stride = l == 0 ? (size_t)NS * NSEQ * 128 : (size_t)2 * TILE_BYTES_H;
            a.k_off = l == 0 ? 0 : (size_t)kr * TILE_BYTES_H;
            a.wih = wih; a.bias = kr == 0 ? bp : nullptr; a.gx = kr == 0 ? gxa : gxb;
            a.n_tiles = n_tiles; a.rows_total = (int64_t)T * Bp; a.ctas_per_combo = ctas_per_combo;
            const int rc = l == 0 ? launch_proj<1>(a, st) : launch_proj<KB_H>(a, st);
            if (rc) return rc;
        }
        for (int d = 0; d < 2; ++d)
            pack_slices_kernel<<<blk, 256, 0, st>>>(Whh_t[2 * l + d], 0, LH, LH, whh + (size_t)d * SLICES * LH * UM);
        // ---- recurrence
        NERRF_CHECK_CUDA(cudaMemsetAsync(counter, 0, 1024, st));
        RecArgs r{};
        r.gx_a = gxa; r.gx_b = l == 0 ? nullptr : gxb; r.whh = whh; r.len = len;
        r.out_img = img[l & 1]; r.hfin = hfin; r.counter = counter;
        r.B = B; r.Bp = Bp; r.n_bt = n_bt; r.T = T; r.groups = groups;
        void* args[] = {&r};
        NERRF_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_rec_kernel, dim3(groups * SLICES), dim3(REC_THREADS), args,
                                                     REC_SMEM, st));
        if (getenv("NERRF_LSTM_PROF")) {           // diagnostics: phase cycle counts of CTA 0 (see lstm_rec_kernel)
            unsigned h[16];
            NERRF_CHECK_CUDA(cudaStreamSynchronize(st));
            NERRF_CHECK_CUDA(cudaMemcpy(h, counter + 128, sizeof(h), cudaMemcpyDeviceToHost));
            const double steps = (double)T * ((n_bt + (groups / 2) * 2 - 1) / ((groups / 2) * 2));
            fprintf(stderr, "[lstm prof] layer %d, cycles per step of CTA 0 / slot 0: epilogue wait-MMA %.0f, compute %.0f, fence+publish %.0f | "
                            "init: G_x loads %.0f, wait-drain %.0f, st+arrive %.0f, poll %.0f\n", l, h[0] * 64.0 / steps, h[1] * 64.0 / steps,
                    h[2] * 64.0 / steps, h[8] * 64.0 / steps, h[9] * 64.0 / steps, h[10] * 64.0 / steps, h[11] * 64.0 / steps);
        }
    }
    return launch_status("lstm_rec_kernel");
}

}  // namespace nerrf
