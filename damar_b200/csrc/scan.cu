// scan.cu -- device-wide exclusive prefix sums used by the index / merge stages.
// Three launches: per-tile reduce, one-block scan of the tile sums, per-tile downsweep.
#include "common.cuh"

namespace damar {

namespace {
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t *total) {
    __shared__ uint64_t wsum[SC_THREADS / 32];
    __shared__ uint64_t tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t inc = warp_incl_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint64_t x = (lane < SC_THREADS / 32) ? wsum[lane] : 0;
        uint64_t xi = warp_incl_scan(x);
        if (lane < SC_THREADS / 32) wsum[lane] = xi - x;
        if (lane == SC_THREADS / 32 - 1) tot = xi;
    }
    __syncthreads();
    uint64_t r = wsum[w] + inc - v;
    *total = tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SC_THREADS) k_tile_reduce(const uint32_t *__restrict__ in, int64_t n,
                                                           uint64_t *__restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        int64_t j = base + i * SC_THREADS + threadIdx.x;
        if (j < n) s += in[j];
    }
    uint64_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// single block: exclusive scan of m sums in place, total to *total
__global__ void __launch_bounds__(SC_THREADS) k_scan_sums(uint64_t *sums, int64_t m, uint64_t *total) {
    uint64_t carry = 0;
    for (int64_t base = 0; base < m; base += SC_THREADS) {
        int64_t j = base + threadIdx.x;
        uint64_t v = (j < m) ? sums[j] : 0, tot;
        uint64_t ex = block_excl_scan(v, &tot);
        if (j < m) sums[j] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

template <typename OutT>
__global__ void __launch_bounds__(SC_THREADS) k_tile_downsweep(const uint32_t *in, OutT *out,
                                                              int64_t n, const uint64_t *__restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        int64_t j = base + i;
        v[i] = (j < n) ? in[j] : 0;
        s += v[i];
    }
    uint64_t tot;
    uint64_t ex = block_excl_scan(s, &tot) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        int64_t j = base + i;
        if (j < n) out[j] = (OutT)ex;
        ex += v[i];
    }
}


template <typename OutT>
void scan_impl(const uint32_t *in, OutT *out, int64_t n, uint64_t *total_dev, cudaStream_t st) {
    if (n <= 0) {
        if (total_dev) CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(uint64_t), st));
        return;
    }
    int64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    DevBuf<uint64_t> &g_sums = scratch().scan_sums;
    g_sums.reserve((size_t)tiles + 1);
    k_tile_reduce<<<(unsigned)tiles, SC_THREADS, 0, st>>>(in, n, g_sums.p);
    KERNEL_CHECK();
    k_scan_sums<<<1, SC_THREADS, 0, st>>>(g_sums.p, tiles, total_dev);
    KERNEL_CHECK();
    k_tile_downsweep<OutT><<<(unsigned)tiles, SC_THREADS, 0, st>>>(in, out, n, g_sums.p);
    KERNEL_CHECK();
    count_launch(3);
}
}  // namespace

void exclusive_scan_u32_to_u64(const uint32_t *in, uint64_t *out, int64_t n, uint64_t *total_dev, cudaStream_t st) {
    scan_impl<uint64_t>(in, out, n, total_dev, st);
}

// total_dev (optional) receives the low 32 bits of the sum
void exclusive_scan_u32_inplace(uint32_t *data, int64_t n, uint32_t *total_dev, cudaStream_t st) {
    DevBuf<uint64_t> &tot = scratch().scan_tot;
    tot.reserve(1);
    scan_impl<uint32_t>(data, data, n, tot.p, st);
    if (total_dev) CUDA_CHECK(cudaMemcpyAsync(total_dev, tot.p, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
}

}  // namespace damar
