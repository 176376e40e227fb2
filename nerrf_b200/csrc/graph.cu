// Temporal-graph constructor, device half: edge list -> CSR-by-destination with temporal
// weights, the exact input layout of the GraphSAGE-T kernels.  Spec = nerrf_b200/graph.py csr_from_edges
// (reference prose: docs/content/docs/architecture.mdx:32-42 "sliding window ... edge weight"):
//
//   order  = stable sort of the edges by (dst, t)            (ties keep input order, like np.lexsort)
//   col[i] = src[order[i]]
//   ew[i]  = conf[e] * exp(-(t_ref - t[e]) / tau),  e = order[i]   (every fp32 op individually rounded)
//   rowptr[v] = number of edges with dst < v
//
// One 64-bit key per edge = dst << 32 | orderable(t); ONE stable LSD radix sort over the low 32 + ceil(log2 N)
// bits (library code: cub::DeviceRadixSort, part of the CUDA toolkit -- HBM-bound byte shuffling, nothing to fuse),
// then one fused pass of ours that gathers src/t/conf through the permutation, evaluates the weight and writes
// rowptr from the dst boundaries of the sorted keys (no atomics anywhere: the result is deterministic).
// String work (path/inode dedup, feature counts) stays on the host: nerrf_b200/graph.py graph_from_events.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace nerrf {

__global__ void __launch_bounds__(256) csr_keys_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                        const float* __restrict__ t, int64_t E, int64_t N,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                        int* __restrict__ bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += stride) {
        int32_t d = dst[i];
        const int32_t s = src[i];
        if ((uint32_t)d >= (uint64_t)N || (uint32_t)s >= (uint64_t)N) { *bad = 1; d = 0; }   // keep the later passes in range
        const uint32_t b = __float_as_uint(t[i] + 0.0f);                 // -0.0 -> +0.0: they compare equal on the host
        const uint32_t ord = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);   // monotone float -> uint
        keys[i] = ((uint64_t)(uint32_t)d << 32) | ord;
        vals[i] = (uint32_t)i;
    }
}

// rowptr[lo..hi] = val; short runs inline by the owning lane, long runs (empty-row gaps, the tail after the last
// edge) by the whole warp so a sparse destination range cannot serialise one thread.
template <typename RP>
__device__ __forceinline__ void fill_rows(RP* __restrict__ rowptr, int64_t lo, int64_t hi, int64_t val, bool active, int lane) {
    const bool is_long = active && (hi - lo) >= 16;
    if (active && !is_long) for (int64_t v = lo; v <= hi; ++v) rowptr[v] = (RP)val;
    unsigned m = __ballot_sync(0xffffffffu, is_long);
    while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const long long lo_ = __shfl_sync(0xffffffffu, (long long)lo, l), hi_ = __shfl_sync(0xffffffffu, (long long)hi, l);
        const long long val_ = __shfl_sync(0xffffffffu, (long long)val, l);
        for (long long v = lo_ + lane; v <= hi_; v += 32) rowptr[v] = (RP)val_;
    }
}

template <typename RP>
__global__ void __launch_bounds__(256) csr_finish_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                                                          const int32_t* __restrict__ src, const float* __restrict__ t,
                                                          const float* __restrict__ conf, int64_t E, int64_t N, float t_ref,
                                                          float tau, RP* __restrict__ rowptr, int32_t* __restrict__ col,
                                                          float* __restrict__ ew) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t wb = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); wb < E; wb += stride) {   // warp-uniform trip count
        const int64_t i = wb + lane;
        const bool ok = i < E;
        int64_t d = 0, dprev = 0;
        if (ok) {
            d = (int64_t)(keys[i] >> 32);
            dprev = i > 0 ? (int64_t)(keys[i - 1] >> 32) : -1;
        }
        fill_rows(rowptr, dprev + 1, d, i, ok && d > dprev, lane);          // first edge of row d and the empty rows before it
        fill_rows(rowptr, d + 1, N, E, ok && i == E - 1, lane);             // rows after the last edge, and rowptr[N]
        if (ok) {
            const uint32_t e = perm[i];
            col[i] = __ldg(src + e);
            const float age = __fsub_rn(t_ref, __ldg(t + e));
            const float z = __fdiv_rn(-age, tau);
            ew[i] = __fmul_rn(__ldg(conf + e), expf(z));
        }
    }
}

template <typename RP>
__global__ void csr_empty_kernel(RP* rowptr, int64_t N) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v <= N; v += stride) rowptr[v] = 0;
}

struct CsrWs {
    size_t keys_a, keys_b, vals_a, vals_b, bad, temp, total, temp_bytes;
};

static inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

static int key_bits(int64_t N) {
    int b = 0;
    while (((int64_t)1 << b) < N) ++b;
    return 32 + (b < 1 ? 1 : b);
}

static int csr_layout(int64_t E, int64_t N, CsrWs* L) {
    size_t temp = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp, k, v, (int64_t)(E > 0 ? E : 1), 0, key_bits(N));
    if (e != cudaSuccess) {
        set_error("cub::DeviceRadixSort size query failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return NERRF_ERR_CUDA;
    }
    size_t o = 0;
    L->keys_a = o; o += up256((size_t)E * 8);
    L->keys_b = o; o += up256((size_t)E * 8);
    L->vals_a = o; o += up256((size_t)E * 4);
    L->vals_b = o; o += up256((size_t)E * 4);
    L->bad = o;    o += 256;
    L->temp = o;   o += up256(temp);
    L->temp_bytes = temp;
    L->total = o;
    return NERRF_OK;
}

}  // namespace nerrf

using namespace nerrf;

extern "C" int nerrf_graph_csr_workspace_bytes(int64_t n_edges, int64_t n_nodes, int64_t* bytes) {
    NERRF_REQUIRE(bytes != nullptr, "null bytes pointer");
    NERRF_REQUIRE(n_edges >= 0 && n_edges < ((int64_t)1 << 32), "n_edges=%lld out of range [0, 2^32)", (long long)n_edges);
    NERRF_REQUIRE(n_nodes >= 1 && n_nodes < ((int64_t)1 << 31), "n_nodes=%lld out of range [1, 2^31)", (long long)n_nodes);
    CsrWs L;
    const int rc = csr_layout(n_edges, n_nodes, &L);
    if (rc != NERRF_OK) return rc;
    *bytes = (int64_t)L.total;
    return NERRF_OK;
}

extern "C" int nerrf_graph_build_csr(const int32_t* src, const int32_t* dst, const float* t, const float* conf,
                                     int64_t n_edges, int64_t n_nodes, float t_ref, float tau, void* rowptr_out,
                                     int rowptr_is64, int32_t* col_out, float* ew_out, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
    NERRF_REQUIRE(n_edges >= 0 && n_edges < ((int64_t)1 << 32), "n_edges=%lld out of range [0, 2^32)", (long long)n_edges);
    NERRF_REQUIRE(n_nodes >= 1 && n_nodes < ((int64_t)1 << 31), "n_nodes=%lld out of range [1, 2^31)", (long long)n_nodes);
    NERRF_REQUIRE(rowptr_out != nullptr, "null rowptr_out");
    NERRF_REQUIRE(rowptr_is64 || n_edges < ((int64_t)1 << 31), "n_edges >= 2^31 needs a 64-bit rowptr");
    NERRF_REQUIRE(tau > 0.f, "tau must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = sm_count() * 8;
    if (n_edges == 0) {
        if (rowptr_is64) csr_empty_kernel<int64_t><<<grid, 256, 0, st>>>((int64_t*)rowptr_out, n_nodes);
        else csr_empty_kernel<int32_t><<<grid, 256, 0, st>>>((int32_t*)rowptr_out, n_nodes);
        return launch_status("csr_empty_kernel");
    }
    NERRF_REQUIRE(src && dst && t && conf && col_out && ew_out, "null edge array");
    CsrWs L;
    int rc = csr_layout(n_edges, n_nodes, &L);
    if (rc != NERRF_OK) return rc;
    NERRF_REQUIRE(workspace != nullptr && workspace_bytes >= (int64_t)L.total,
                  "workspace too small: %lld < %lld bytes (nerrf_graph_csr_workspace_bytes)", (long long)workspace_bytes,
                  (long long)L.total);
    NERRF_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    char* ws = (char*)workspace;
    uint64_t* ka = (uint64_t*)(ws + L.keys_a); uint64_t* kb = (uint64_t*)(ws + L.keys_b);
    uint32_t* va = (uint32_t*)(ws + L.vals_a); uint32_t* vb = (uint32_t*)(ws + L.vals_b);
    int* bad = (int*)(ws + L.bad);
    NERRF_CHECK_CUDA(cudaMemsetAsync(bad, 0, 4, st));
    csr_keys_kernel<<<grid, 256, 0, st>>>(src, dst, t, n_edges, n_nodes, ka, va, bad);
    rc = launch_status("csr_keys_kernel");
    if (rc != NERRF_OK) return rc;
    cub::DoubleBuffer<uint64_t> k(ka, kb);
    cub::DoubleBuffer<uint32_t> v(va, vb);
    size_t temp = L.temp_bytes;
    NERRF_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(ws + L.temp, temp, k, v, n_edges, 0, key_bits(n_nodes), st));
    if (rowptr_is64)
        csr_finish_kernel<int64_t><<<grid, 256, 0, st>>>(k.Current(), v.Current(), src, t, conf, n_edges, n_nodes, t_ref, tau,
                                                         (int64_t*)rowptr_out, col_out, ew_out);
    else
        csr_finish_kernel<int32_t><<<grid, 256, 0, st>>>(k.Current(), v.Current(), src, t, conf, n_edges, n_nodes, t_ref, tau,
                                                         (int32_t*)rowptr_out, col_out, ew_out);
    rc = launch_status("csr_finish_kernel");
    if (rc != NERRF_OK) return rc;
    // the vertex-id check is the one reason this call waits for the stream: out-of-range ids were clamped to row 0
    // on the device (so nothing was written out of bounds) and are reported here as NERRF_ERR_INVALID
    int h_bad = 0;
    NERRF_CHECK_CUDA(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st));
    NERRF_CHECK_CUDA(cudaStreamSynchronize(st));
    NERRF_REQUIRE(h_bad == 0, "edge list holds a vertex id outside [0, n_nodes)");
    return NERRF_OK;
}
