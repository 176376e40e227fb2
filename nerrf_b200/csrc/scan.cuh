// Exclusive prefix sum of int32 values in place (device-wide), shared by the node interning (creating-mention flags -> node
// ids) and the radix sort (digit histograms -> scatter offsets).  Three launches: per-tile scan + tile sums, scan of the
// sums by one CTA, add.  Deterministic (integer).
#pragma once
#include "common.cuh"

namespace nerrf {

// ---- exclusive scan of int32 [m] in place: per-block (1024 elements) scan + block sums, scan of the sums, add
constexpr int SCAN_TILE = 1024;
static __global__ void __launch_bounds__(256) scan_tiles_kernel(int32_t* __restrict__ v, int64_t m, int32_t* __restrict__ sums) {
    __shared__ int32_t warp_tot[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    int32_t x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = (base + k < m) ? v[base + k] : 0;
    const int32_t mine = x[0] + x[1] + x[2] + x[3];
    int32_t inc = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int32_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
    int32_t run = wbase + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (base + k < m) v[base + k] = run; run += x[k]; }
    if (threadIdx.x == 255) sums[blockIdx.x] = wbase + inc;
}
static __global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ sums, int64_t nb, int32_t* __restrict__ total) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < nb; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const int32_t x = i < nb ? sums[i] : 0;
        int32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int32_t wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
        const int32_t carry = carry_s;
        if (i < nb) sums[i] = carry + wbase + inc - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}
static __global__ void __launch_bounds__(256) scan_add_kernel(int32_t* __restrict__ v, int64_t m, const int32_t* __restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    const int32_t add = sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k < m) v[base + k] += add;
}


// v[0..m) -> exclusive prefix sums in place; *total (device) = the sum.  sums: scratch of ceil(m / SCAN_TILE) + 1 ints.
inline void exclusive_scan_i32(int32_t* v, int64_t m, int32_t* sums, int32_t* total, cudaStream_t st) {
    const int64_t nb = (m + SCAN_TILE - 1) / SCAN_TILE;
    scan_tiles_kernel<<<(unsigned)nb, 256, 0, st>>>(v, m, sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nb, total);
    scan_add_kernel<<<(unsigned)nb, 256, 0, st>>>(v, m, sums);
}

}  // namespace nerrf
