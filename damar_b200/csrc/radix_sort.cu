// radix_sort.cu -- stable LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits.
//
// Replaces lex_sort/lex_thread (reference dalign/filter.c:230-435): same contract (stable,
// ascending on the selected key bits), different machine.  Per pass:
//   rs_hist    : per-tile digit histogram (coalesced 8-byte loads, shared-memory bins)
//   scan       : exclusive scan of the digit-major [256][tiles] table
//   rs_scatter : warp match_any ranking (stable), tile-local reorder in shared memory, then
//                coalesced runs to global memory.
#include <mutex>

#include "common.cuh"

namespace damar {
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 pairs per CTA
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WTILE = RS_TILE / RS_WARPS;     // 512 pairs per warp

__global__ void __launch_bounds__(RS_THREADS) rs_hist(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                     uint32_t *__restrict__ tile_hist, int ntiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        int64_t j = base + i * RS_THREADS + threadIdx.x;
        if (j < n) atomicAdd(&h[(keys[j] >> shift) & 255], 1u);
    }
    __syncthreads();
    tile_hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS, 3) rs_scatter(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                        uint64_t *__restrict__ kout, uint32_t *__restrict__ vout,
                                                        int64_t n, int shift, const uint32_t *__restrict__ tile_off,
                                                        int ntiles) {
    __shared__ uint32_t whist[RS_WARPS][256];
    __shared__ uint32_t dstart[256];
    __shared__ uint32_t gbase[256];
    __shared__ uint32_t wtot[RS_WARPS];
    extern __shared__ __align__(16) unsigned char dyn[];
    uint64_t *skeys = (uint64_t *)dyn;
    uint32_t *svals = (uint32_t *)(dyn + sizeof(uint64_t) * RS_TILE);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    const int nvalid = (int)((n - base < RS_TILE) ? (n - base) : RS_TILE);

    // keys stay in registers for the whole pass; the values are only fetched when their slot is known, which keeps
    // the kernel at 80 registers = 3 CTAs (24 warps) per SM
    uint64_t key[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        int t = w * RS_WTILE + i * 32 + lane;
        key[i] = (t < nvalid) ? kin[base + t] : ~0ull;
    }
    for (int d = tid; d < RS_WARPS * 256; d += RS_THREADS) (&whist[0][0])[d] = 0;
    __syncthreads();

#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t lower = peers & ((1u << lane) - 1u);
        const uint32_t prev = whist[w][d];
        __syncwarp();
        if (lower == 0) whist[w][d] = prev + __popc(peers);
        __syncwarp();
        rank[i] = (uint16_t)(prev + __popc(lower));
    }
    __syncthreads();

    // per digit: exclusive scan over warps, tile count
    uint32_t cnt;
    {
        uint32_t run = 0;
#pragma unroll
        for (int x = 0; x < RS_WARPS; x++) { uint32_t t = whist[x][tid]; whist[x][tid] = run; run += t; }
        if (tid == 255) run -= (uint32_t)(RS_TILE - nvalid);   // padding keys all carry digit 255 and sort last
        cnt = run;
    }
    // exclusive scan of the 256 tile counts
    {
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wtot[w] = inc;
        __syncthreads();
        uint32_t off = 0;
#pragma unroll
        for (int x = 0; x < RS_WARPS; x++) if (x < w) off += wtot[x];
        dstart[tid] = off + inc - cnt;
        gbase[tid] = tile_off[(size_t)tid * ntiles + blockIdx.x];
    }
    __syncthreads();

#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
        const uint32_t p = dstart[d] + whist[w][d] + rank[i];
        const int t = w * RS_WTILE + i * 32 + lane;
        if (t < nvalid) { skeys[p] = key[i]; svals[p] = vin[base + t]; }
    }
    __syncthreads();

    for (int j = tid; j < nvalid; j += RS_THREADS) {
        const uint64_t k = skeys[j];
        const uint32_t d = (uint32_t)(k >> shift) & 255u;
        const size_t g = (size_t)gbase[d] + (uint32_t)(j - dstart[d]);
        kout[g] = k;
        vout[g] = svals[j];
    }
}

}  // namespace

// Sorts on key bits [0, nbits).  On return keys/vals point at the sorted data and
// keys_alt/vals_alt at the scratch copy (the pointers are swapped as needed).
void radix_sort_pairs(uint64_t *&keys, uint32_t *&vals, uint64_t *&keys_alt, uint32_t *&vals_alt,
                      int64_t n, int nbits, cudaStream_t st) {
    if (n <= 1 || nbits <= 0) return;
    if (n >= (int64_t)0xffffffffll) fatal("radix_sort_pairs: %lld elements exceed the 32-bit offset range", (long long)n);
    const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
    WorkerScratch &ws = scratch();
    DevBuf<uint32_t> &g_tile_hist = ws.tile_hist;
    g_tile_hist.reserve((size_t)ntiles * 256);
    static std::once_flag attr_set;
    const size_t dyn = (sizeof(uint64_t) + sizeof(uint32_t)) * RS_TILE;
    std::call_once(attr_set, [dyn] { CUDA_CHECK(cudaFuncSetAttribute(rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); });
    cudaEvent_t (&ev)[16][2] = ws.rs_ev;
    if (!ws.rs_ev_init) { for (auto &e : ev) { CUDA_CHECK(cudaEventCreateWithFlags(&e[0], timing_event_flags())); CUDA_CHECK(cudaEventCreateWithFlags(&e[1], timing_event_flags())); } ws.rs_ev_init = true; }
    int npass = 0;
    for (int shift = 0; shift < nbits; shift += 8, npass++) {
        rs_hist<<<ntiles, RS_THREADS, 0, st>>>(keys, n, shift, g_tile_hist.p, ntiles);
        KERNEL_CHECK();
        exclusive_scan_u32_inplace(g_tile_hist.p, (int64_t)ntiles * 256, nullptr, st);
        CUDA_CHECK(cudaEventRecord(ev[npass][0], st));
        rs_scatter<<<ntiles, RS_THREADS, dyn, st>>>(keys, vals, keys_alt, vals_alt, n, shift, g_tile_hist.p, ntiles);
        KERNEL_CHECK();
        CUDA_CHECK(cudaEventRecord(ev[npass][1], st));
        count_launch(2);
        uint64_t *tk = keys; keys = keys_alt; keys_alt = tk;
        uint32_t *tv = vals; vals = vals_alt; vals_alt = tv;
    }
    // per-launch device time of the scatter kernel (feeds bench.py's roofline line)
    CUDA_CHECK(cudaEventSynchronize(ev[npass - 1][1]));
    for (int i = 0; i < npass; i++) {
        float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, ev[i][0], ev[i][1]));
        ctx().tm.scatter_ms += ms; ctx().tm.scatter_elems += n; ctx().tm.scatter_launches += 1;
    }
}

}  // namespace damar
