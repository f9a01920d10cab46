// radix_sort.cu -- stable LSD radix sort of u64 keys (optionally with a u32 payload) on a bit range, one sweep per digit.
//
// Replaces lex_sort/lex_thread (reference dalign/filter.c:230-435): same contract (stable, ascending on the selected
// key bits), different machine:
//   os_hist  : ONE read of the keys builds the digit histograms of every pass (shared-memory bins, flushed with one
//              global atomic per non-empty bin and CTA)
//   os_scan  : exclusive scan of each pass's histogram -> first output slot of every digit
//   os_pass  : per digit, a persistent kernel (two 512-thread CTAs per SM).  A CTA draws tile numbers from a global
//              counter; a tile's keys arrive in shared memory through ONE bulk copy (cp.async.bulk + mbarrier, no
//              per-thread loads), are ranked stably (ballot matching per digit bit, warp-private 16-bit counters), the
//              tile's per-digit counts are published and the digit offsets of all earlier tiles are collected by
//              decoupled look-back (Merrill & Garland's chained scan as used by Onesweep, Adinets & Merrill 2022) while
//              the tile is reordered in shared memory; runs of equal digits then leave as coalesced stores.
// Against the round-1 sort (histogram kernel + 3-kernel scan + scatter per 8-bit digit): the keys are read once per
// pass instead of twice, there are no per-pass scan launches, and digits are up to 10 bits wide (28-bit k-mer codes:
// 3 passes instead of 4; 44-bit seed-hit keys: 5 instead of 6).  Keys-only sorts (records packed into one 64-bit
// word by the callers) move 16 B per element and pass instead of 24.
#include <mutex>

#include "common.cuh"

namespace damar {
namespace {

constexpr int OS_THREADS = 512;
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_MAX_PASSES = 8;
constexpr int OS_MAX_RADIX = 1024;
constexpr uint32_t ST_AGG = 1u << 30, ST_INCL = 2u << 30, ST_VAL = (1u << 30) - 1u;

// ---- mbarrier / bulk-copy helpers (PTX ISA 8.x, sm_90+; SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
// generic-proxy accesses of shared memory before this point are ordered before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct HistArgs {
    const uint64_t *keys; int64_t n; int npass;
    int shift[OS_MAX_PASSES]; uint32_t mask[OS_MAX_PASSES];
    uint32_t *ghist;                                   // [npass][OS_MAX_RADIX]
};

__global__ void __launch_bounds__(OS_THREADS) os_hist(HistArgs a) {
    extern __shared__ uint32_t sh[];                   // [npass][OS_MAX_RADIX]
    for (int i = threadIdx.x; i < a.npass * OS_MAX_RADIX; i += OS_THREADS) sh[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * OS_THREADS * 2;
    const bool aligned = (reinterpret_cast<uintptr_t>(a.keys) & 15) == 0;
    for (int64_t i = ((int64_t)blockIdx.x * OS_THREADS + threadIdx.x) * 2; i < a.n; i += stride) {
        uint64_t k0, k1 = 0;
        const bool two = i + 1 < a.n;
        if (two && aligned) { const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(a.keys + i); k0 = v.x; k1 = v.y; }
        else { k0 = a.keys[i]; if (two) k1 = a.keys[i + 1]; }
        for (int p = 0; p < a.npass; p++) {
            atomicAdd(&sh[p * OS_MAX_RADIX + ((uint32_t)(k0 >> a.shift[p]) & a.mask[p])], 1u);
            if (two) atomicAdd(&sh[p * OS_MAX_RADIX + ((uint32_t)(k1 >> a.shift[p]) & a.mask[p])], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.npass * OS_MAX_RADIX; i += OS_THREADS)
        if (sh[i]) atomicAdd(&a.ghist[i], sh[i]);
}

// one CTA per pass: ghist[p][d] <- number of keys whose digit in pass p is smaller than d
__global__ void __launch_bounds__(OS_MAX_RADIX) os_scan(uint32_t *ghist) {
    __shared__ uint32_t wsum[32];
    uint32_t *h = ghist + (size_t)blockIdx.x * OS_MAX_RADIX;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t v = h[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = wsum[lane], xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
        wsum[lane] = xi - x;
    }
    __syncthreads();
    h[threadIdx.x] = wsum[w] + inc - v;
}

struct PassArgs {
    const uint64_t *kin; const uint32_t *vin; uint64_t *kout; uint32_t *vout;
    int64_t n; int shift; uint32_t mask;
    const uint32_t *digit_base;                        // [RADIX] first output slot of every digit (os_scan)
    uint32_t *state;                                   // [ntiles][RADIX] look-back words: flag (2 bits) | count (30 bits)
    unsigned int *counter; unsigned int ntiles;
    int bulk;                                          // sources are 16-byte aligned: full tiles arrive by bulk copy
};

template <int BITS, int ITEMS, bool HAS_VALS>
struct OsCfg {
    static constexpr int RADIX = 1 << BITS;
    static constexpr int TILE = OS_THREADS * ITEMS;
    static constexpr int WTILE = 32 * ITEMS;
    static constexpr int KEY_BYTES = TILE * 8;
    static constexpr int VAL_BYTES = HAS_VALS ? TILE * 4 : 0;
    static constexpr int HIST_BYTES = OS_WARPS * RADIX * 2;      // warp-private u16 counters
    static constexpr int SMEM = KEY_BYTES + VAL_BYTES + HIST_BYTES + 2 * RADIX * 4;
    static_assert(2 * (SMEM + 1024) <= 227 * 1024, "two CTAs per SM");
    static_assert(RADIX / 2 <= OS_THREADS, "one thread per pair of digits");
    static_assert(TILE < 65536, "16-bit ranks");
};

template <int BITS, int ITEMS, bool HAS_VALS>
__global__ void __launch_bounds__(OS_THREADS, 2) os_pass(PassArgs a) {
    using C = OsCfg<BITS, ITEMS, HAS_VALS>;
    constexpr int RADIX = C::RADIX, HW = RADIX / 2, TILE = C::TILE, WTILE = C::WTILE;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *skeys = reinterpret_cast<uint64_t *>(smem);
    uint32_t *svals = reinterpret_cast<uint32_t *>(smem + C::KEY_BYTES);
    uint32_t *whist = reinterpret_cast<uint32_t *>(smem + C::KEY_BYTES + C::VAL_BYTES);   // [OS_WARPS][HW], digits 2j / 2j+1 in the halves of word j
    uint32_t *dstart = reinterpret_cast<uint32_t *>(smem + C::KEY_BYTES + C::VAL_BYTES + C::HIST_BYTES);
    uint32_t *gbase = dstart + RADIX;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ unsigned s_tile;
    __shared__ uint32_t wtot[OS_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    if (tid == 0) mbar_init(&mbar, 1);
    uint32_t phase = 0;
    for (;;) {
        fence_async_smem();
        if (tid == 0) s_tile = atomicAdd(a.counter, 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= a.ntiles) break;
        const int64_t base = (int64_t)tile * TILE;
        const int nvalid = (int)((a.n - base < TILE) ? (a.n - base) : TILE);
        const bool full = nvalid == TILE;
        const bool by_bulk = full && a.bulk;

        // ---- the tile's keys (and payloads) into shared memory: one bulk copy each ----
        for (int j = lane; j < HW; j += 32) whist[w * HW + j] = 0;
        if (by_bulk) {
            if (tid == 0) {
                mbar_expect_tx(&mbar, (uint32_t)(C::KEY_BYTES + C::VAL_BYTES));
                bulk_g2s(skeys, a.kin + base, C::KEY_BYTES, &mbar);
                if (HAS_VALS) bulk_g2s(svals, a.vin + base, C::VAL_BYTES, &mbar);
            }
            mbar_wait(&mbar, phase);
            phase ^= 1u;
        } else {                                       // last, partial tile (or unaligned source): plain loads, a bulk copy needs 16-byte multiples
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int t = w * WTILE + i * 32 + lane;
                skeys[t] = (t < nvalid) ? a.kin[base + t] : ~0ull;
                if (HAS_VALS && t < nvalid) svals[t] = a.vin[base + t];
            }
        }
        __syncwarp();                                  // a warp ranks the keys it loaded (or that the bulk copy delivered)

        // ---- stable rank inside the warp's 32 x ITEMS keys; the keys stay in shared memory ----
        uint32_t rank2[(ITEMS + 1) / 2];               // 16-bit ranks, two per register
#pragma unroll
        for (int i = 0; i < (ITEMS + 1) / 2; i++) rank2[i] = 0;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t d = (uint32_t)(skeys[w * WTILE + i * 32 + lane] >> a.shift) & a.mask;
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < BITS; b++)             // lanes whose digit agrees with mine in bit b: 4 instructions per bit
                asm("{ .reg .pred p; .reg .b32 t, bal;\n"
                    "  and.b32 t, %1, %2; setp.ne.u32 p, t, 0; vote.sync.ballot.b32 bal, p, 0xffffffff;\n"
                    "  @p and.b32 %0, %0, bal; @!p lop3.b32 %0, %0, bal, 0, 0x30; }"
                    : "+r"(peers) : "r"(d), "r"(1u << b));
            const uint32_t lower = peers & lt;
            const int sh16 = (int)(d & 1u) * 16;
            uint32_t prev = 0;
            if (lower == 0) prev = atomicAdd(&whist[w * HW + (d >> 1)], (uint32_t)__popc(peers) << sh16);
            prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
            rank2[i >> 1] |= (((prev >> sh16) & 0xffffu) + (uint32_t)__popc(lower)) << ((i & 1) * 16);
        }
        __syncthreads();

        // ---- per digit: exclusive scan over the warps, tile count; then the exclusive scan over the digits ----
        uint32_t cnt_lo = 0, cnt_hi = 0;
        if (tid < HW) {
            uint32_t run = 0;
#pragma unroll
            for (int x = 0; x < OS_WARPS; x++) { const uint32_t t = whist[x * HW + tid]; whist[x * HW + tid] = run; run += t; }
            cnt_lo = run & 0xffffu; cnt_hi = run >> 16;
            if (!full) {                               // padding keys carry the largest digit and rank last
                const uint32_t pad = (uint32_t)(TILE - nvalid);
                if ((uint32_t)(2 * tid) == a.mask) cnt_lo -= pad;
                if ((uint32_t)(2 * tid + 1) == a.mask) cnt_hi -= pad;
            }
            // publish the tile's counts (tile 0: they are final)
            const uint32_t f = (tile == 0) ? ST_INCL : ST_AGG;
            const uint64_t word = (uint64_t)(f | cnt_lo) | ((uint64_t)(f | cnt_hi) << 32);
            *reinterpret_cast<volatile uint64_t *>(a.state + (size_t)tile * RADIX + 2 * tid) = word;
        }
        {
            const uint32_t s = cnt_lo + cnt_hi;
            uint32_t inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) wtot[w] = inc;
            __syncthreads();
            uint32_t off = 0;
#pragma unroll
            for (int x = 0; x < OS_WARPS; x++) if (x < w) off += wtot[x];
            if (tid < HW) { dstart[2 * tid] = off + inc - s; dstart[2 * tid + 1] = off + inc - s + cnt_lo; }
        }
        __syncthreads();

        // ---- slot of every key inside the reordered tile; the keys pass through registers on their way ----
        uint64_t key[ITEMS];
        uint32_t val[HAS_VALS ? ITEMS : 1];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int t = w * WTILE + i * 32 + lane;
            key[i] = skeys[t];
            const uint32_t d = (uint32_t)(key[i] >> a.shift) & a.mask;
            const uint32_t wp = (whist[w * HW + (d >> 1)] >> ((d & 1u) * 16)) & 0xffffu;
            rank2[i >> 1] += (dstart[d] + wp) << ((i & 1) * 16);   // slot < TILE < 2^16: no carry into the neighbour
            if (HAS_VALS) val[i] = (t < nvalid) ? svals[t] : 0u;
        }
        __syncthreads();                               // every key is in registers: the buffer becomes the reordered tile
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t p = (rank2[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
            if (p < (uint32_t)nvalid) { skeys[p] = key[i]; if (HAS_VALS) svals[p] = val[i]; }
        }

        // ---- decoupled look-back: keys of this digit in all earlier tiles ----
        if (tid < HW) {
            uint32_t ex_lo = 0, ex_hi = 0;
            if (tile > 0) {
                const volatile uint64_t *st = reinterpret_cast<const volatile uint64_t *>(a.state) + tid;
                for (int64_t j = (int64_t)tile - 1;;) {
                    const uint64_t s = st[(size_t)j * HW];
                    const uint32_t f = (uint32_t)s & ~ST_VAL;
                    if (f == 0) continue;              // not published yet
                    ex_lo += (uint32_t)s & ST_VAL; ex_hi += (uint32_t)(s >> 32) & ST_VAL;
                    if (f == ST_INCL) break;
                    j--;
                }
                const uint64_t word = (uint64_t)(ST_INCL | (ex_lo + cnt_lo)) | ((uint64_t)(ST_INCL | (ex_hi + cnt_hi)) << 32);
                *reinterpret_cast<volatile uint64_t *>(a.state + (size_t)tile * RADIX + 2 * tid) = word;
            }
            gbase[2 * tid] = a.digit_base[2 * tid] + ex_lo - dstart[2 * tid];
            gbase[2 * tid + 1] = a.digit_base[2 * tid + 1] + ex_hi - dstart[2 * tid + 1];
        }
        __syncthreads();

        // ---- runs of equal digits leave as coalesced stores ----
        for (int j = tid; j < nvalid; j += OS_THREADS) {
            const uint64_t k = skeys[j];
            const uint32_t d = (uint32_t)(k >> a.shift) & a.mask;
            const uint32_t g = gbase[d] + (uint32_t)j;
            a.kout[g] = k;
            if (HAS_VALS) a.vout[g] = svals[j];
        }
    }
}

template <int BITS, int ITEMS, bool HAS_VALS>
void launch_pass(const PassArgs &a, cudaStream_t st) {
    using C = OsCfg<BITS, ITEMS, HAS_VALS>;
    static std::once_flag attr;
    std::call_once(attr, [] { CUDA_CHECK(cudaFuncSetAttribute(os_pass<BITS, ITEMS, HAS_VALS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM)); });
    const unsigned grid = a.ntiles < 2u * NUM_SMS_B200 ? a.ntiles : 2u * NUM_SMS_B200;
    os_pass<BITS, ITEMS, HAS_VALS><<<grid, OS_THREADS, C::SMEM, st>>>(a);
    KERNEL_CHECK();
}

constexpr int ITEMS_KEYS = 16, ITEMS_PAIRS = 12;
int keys_items() {                                     // DAMAR_SORT_ITEMS=12|14|16: keys per thread of the keys-only kernel (tuning knob)
    static const int v = [] { const char *e = getenv("DAMAR_SORT_ITEMS"); const int x = (e && *e) ? atoi(e) : ITEMS_KEYS; return (x == 12 || x == 14) ? x : ITEMS_KEYS; }();
    return v;
}
template <int BITS, bool HAS_VALS>
void launch_items(const PassArgs &a, int items, cudaStream_t st) {
    if constexpr (HAS_VALS) launch_pass<BITS, ITEMS_PAIRS, true>(a, st);
    else {
        if (items == 12) launch_pass<BITS, 12, false>(a, st);
        else if (items == 14) launch_pass<BITS, 14, false>(a, st);
        else launch_pass<BITS, ITEMS_KEYS, false>(a, st);
    }
}

int max_digit_bits() {
    static const int b = [] { const char *e = getenv("DAMAR_SORT_BITS"); const int v = (e && *e) ? atoi(e) : 10; return v < 8 ? 8 : (v > 10 ? 10 : v); }();
    return b;
}

}  // namespace

// Sorts on key bits [lo_bit, hi_bit).  vals may be null (keys-only).  On return keys/vals point at the sorted data and
// keys_alt/vals_alt at the scratch copy (the pointers are swapped as needed).  Returns 0, or 1 with the message set
// when the input is too large for the 30-bit look-back counters.
int radix_sort(uint64_t *&keys, uint32_t *&vals, uint64_t *&keys_alt, uint32_t *&vals_alt, int64_t n, int lo_bit, int hi_bit,
               cudaStream_t st) {
    const int nbits = hi_bit - lo_bit;
    if (n <= 1 || nbits <= 0) return 0;
    if (n >= (int64_t)ST_VAL) { set_error("radix_sort: %lld elements exceed the 2^30 range of one sort", (long long)n); return 1; }
    const bool has_vals = vals != nullptr;
    const int maxb = max_digit_bits();
    int npass = (nbits + maxb - 1) / maxb;
    if (npass > OS_MAX_PASSES) { set_error("radix_sort: %d key bits need more than %d passes", nbits, OS_MAX_PASSES); return 1; }
    int bits[OS_MAX_PASSES], cap = 8;
    for (int p = 0; p < npass; p++) { bits[p] = nbits / npass + (p < nbits % npass ? 1 : 0); if (bits[p] > cap) cap = bits[p]; }
    const int items = has_vals ? ITEMS_PAIRS : keys_items();
    const int tile = OS_THREADS * items;
    const unsigned ntiles = (unsigned)((n + tile - 1) / tile);
    const int radix = 1 << cap;

    WorkerScratch &ws = scratch();
    DevBuf<uint32_t> &buf = ws.tile_hist;             // [64: counter] [npass][1024: digit bases] [ntiles][radix: look-back state]
    const size_t state_off = 64 + (size_t)OS_MAX_PASSES * OS_MAX_RADIX;
    buf.reserve(state_off + (size_t)ntiles * radix);
    uint32_t *ghist = buf.p + 64, *state = buf.p + state_off;
    CUDA_CHECK(cudaMemsetAsync(buf.p, 0, sizeof(uint32_t) * (64 + (size_t)npass * OS_MAX_RADIX), st));

    cudaEvent_t (&ev)[16][2] = ws.rs_ev;
    if (!ws.rs_ev_init) { for (auto &e : ev) { CUDA_CHECK(cudaEventCreateWithFlags(&e[0], timing_event_flags())); CUDA_CHECK(cudaEventCreateWithFlags(&e[1], timing_event_flags())); } ws.rs_ev_init = true; }
    CUDA_CHECK(cudaEventRecord(ev[15][0], st));        // the whole sort: histogram, scan and every pass
    HistArgs h{};
    h.keys = keys; h.n = n; h.npass = npass; h.ghist = ghist;
    for (int p = 0, s = lo_bit; p < npass; s += bits[p], p++) { h.shift[p] = s; h.mask[p] = (1u << bits[p]) - 1u; }
    {
        int64_t want = (n + 2 * OS_THREADS * 8 - 1) / (2 * OS_THREADS * 8);
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 4 * NUM_SMS_B200 ? 4 * NUM_SMS_B200 : want));
        os_hist<<<grid, OS_THREADS, sizeof(uint32_t) * npass * OS_MAX_RADIX, st>>>(h);
        KERNEL_CHECK();
        os_scan<<<npass, OS_MAX_RADIX, 0, st>>>(ghist);
        KERNEL_CHECK();
        count_launch(2);
    }
    for (int p = 0; p < npass; p++) {
        CUDA_CHECK(cudaMemsetAsync(buf.p, 0, sizeof(uint32_t) * 64, st));
        CUDA_CHECK(cudaMemsetAsync(state, 0, sizeof(uint32_t) * (size_t)ntiles * radix, st));
        PassArgs a{};
        a.kin = keys; a.vin = vals; a.kout = keys_alt; a.vout = vals_alt; a.n = n; a.shift = h.shift[p]; a.mask = h.mask[p];
        a.bulk = ((reinterpret_cast<uintptr_t>(keys) & 15) == 0 && (!has_vals || (reinterpret_cast<uintptr_t>(vals) & 15) == 0)) ? 1 : 0;
        a.digit_base = ghist + (size_t)p * OS_MAX_RADIX; a.state = state; a.counter = buf.p; a.ntiles = ntiles;
        CUDA_CHECK(cudaEventRecord(ev[p][0], st));
        if (has_vals) {
            if (cap <= 8) launch_items<8, true>(a, items, st);
            else if (cap == 9) launch_items<9, true>(a, items, st);
            else launch_items<10, true>(a, items, st);
        } else {
            if (cap <= 8) launch_items<8, false>(a, items, st);
            else if (cap == 9) launch_items<9, false>(a, items, st);
            else launch_items<10, false>(a, items, st);
        }
        CUDA_CHECK(cudaEventRecord(ev[p][1], st));
        count_launch(1);
        uint64_t *tk = keys; keys = keys_alt; keys_alt = tk;
        if (has_vals) { uint32_t *tv = vals; vals = vals_alt; vals_alt = tv; }
    }
    // per-launch device time of the pass kernel (feeds bench.py's roofline line)
    CUDA_CHECK(cudaEventSynchronize(ev[npass - 1][1]));
    {
        float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, ev[15][0], ev[npass - 1][1]));
        ctx().tm.sort_ms += ms; ctx().tm.sort_elems += n;
        ctx().tm.sort_contract_bytes += 16 * n * (1 + 2 * (int64_t)((nbits + 7) / 8));
    }
    for (int p = 0; p < npass; p++) {
        float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, ev[p][0], ev[p][1]));
        ctx().tm.scatter_ms += ms; ctx().tm.scatter_elems += n; ctx().tm.scatter_launches += 1;
        ctx().tm.scatter_bytes += (int64_t)n * (has_vals ? 24 : 16);
    }
    return 0;
}

// ---- self test / micro benchmark (damar_sort_selftest): random keys, sorted, checked on the device ----
namespace {
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}
// key = random sort bits above lo_bit, the element's index below (when it fits): a stable sort leaves equal keys in index order
__global__ void st_fill(uint64_t *keys, uint32_t *vals, int64_t n, int lo_bit, int hi_bit, uint64_t seed, int skew) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r = mix64(seed ^ (uint64_t)i);
    if (skew) r &= mix64(r);                           // fewer set bits: long runs of equal digits
    const uint64_t hi = (hi_bit - lo_bit >= 64) ? r : (r & ((1ull << (hi_bit - lo_bit)) - 1ull));
    const uint64_t low = lo_bit > 0 ? ((uint64_t)i & ((1ull << lo_bit) - 1ull)) : 0ull;
    keys[i] = (hi << lo_bit) | low;
    if (vals) vals[i] = (uint32_t)i;
}
__global__ void st_check(const uint64_t *keys, const uint32_t *vals, int64_t n, int lo_bit, int hi_bit, int low_is_index,
                         unsigned long long *bad, unsigned long long *sum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h = 0; bool v = false;
    if (i < n) {
        const uint64_t m = (hi_bit - lo_bit >= 64) ? ~0ull : ((1ull << (hi_bit - lo_bit)) - 1ull);
        const uint64_t k = keys[i];
        h = mix64(k ^ (vals ? ((uint64_t)vals[i] << 1) : 0ull));
        if (i > 0 && hi_bit > lo_bit) {
            const uint64_t a = (keys[i - 1] >> lo_bit) & m, b = (k >> lo_bit) & m;
            v = a > b;
            if (a == b) {
                if (vals && vals[i - 1] >= vals[i]) v = true;
                if (!vals && low_is_index && (keys[i - 1] & ((1ull << lo_bit) - 1ull)) >= (k & ((1ull << lo_bit) - 1ull))) v = true;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    const unsigned nv = __popc(__ballot_sync(0xffffffffu, v));
    if ((threadIdx.x & 31) == 0) { if (h) atomicAdd(sum, h); if (nv) atomicAdd(bad, (unsigned long long)nv); }
}
}  // namespace

int sort_selftest(int64_t n, int lo_bit, int hi_bit, int has_vals, int reps, uint64_t seed, int skew, double *ms_out) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    DevBuf<uint64_t> k0, k1; DevBuf<uint32_t> v0, v1; DevBuf<unsigned long long> res;
    k0.reserve((size_t)n + 2); k1.reserve((size_t)n + 2); res.reserve(4);
    if (has_vals) { v0.reserve((size_t)n + 2); v1.reserve((size_t)n + 2); }
    const unsigned nb = (unsigned)((n + 255) / 256);
    const int low_is_index = (lo_bit > 0 && lo_bit < 63 && (1ll << lo_bit) >= n) ? 1 : 0;
    double total = 0;
    int rc = 0;
    for (int r = 0; r < reps && rc == 0; r++) {
        uint64_t *ka = k0.p, *kb = k1.p; uint32_t *va = has_vals ? v0.p : nullptr, *vb = has_vals ? v1.p : nullptr;
        CUDA_CHECK(cudaMemsetAsync(res.p, 0, 4 * sizeof(unsigned long long), st));
        st_fill<<<nb, 256, 0, st>>>(ka, va, n, lo_bit, hi_bit, seed + r, skew);
        st_check<<<nb, 256, 0, st>>>(ka, va, n, lo_bit, 0, 0, res.p + 2, res.p + 1);      // checksum of the input (hi_bit 0: no order test)
        KERNEL_CHECK();
        Timer t(st);
        rc = radix_sort(ka, va, kb, vb, n, lo_bit, hi_bit, st);
        total += t.stop();
        if (rc) break;
        st_check<<<nb, 256, 0, st>>>(ka, va, n, lo_bit, hi_bit, low_is_index, res.p, res.p + 3);
        KERNEL_CHECK();
        unsigned long long h[4];
        CUDA_CHECK(cudaMemcpyAsync(h, res.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_wait(st));
        if (h[0] != 0) { set_error("sort_selftest: %llu order/stability violations (n=%lld bits [%d,%d) vals=%d)", h[0], (long long)n, lo_bit, hi_bit, has_vals); rc = 2; }
        else if (h[1] != h[3]) { set_error("sort_selftest: the output is not a permutation of the input (n=%lld)", (long long)n); rc = 3; }
    }
    if (ms_out) *ms_out = reps > 0 ? total / reps : 0;
    return rc;
}

void radix_sort_pairs(uint64_t *&keys, uint32_t *&vals, uint64_t *&keys_alt, uint32_t *&vals_alt,
                      int64_t n, int nbits, cudaStream_t st) {
    if (radix_sort(keys, vals, keys_alt, vals_alt, n, 0, nbits, st) != 0) fatal("radix_sort_pairs failed");
}

}  // namespace damar
