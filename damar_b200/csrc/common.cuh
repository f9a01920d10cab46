// common.cuh -- shared declarations of the damar_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/damar_b200.h"

namespace damar {

// ---- error handling: the reference prints to stderr and exit(1)s (db/DB.h:84-86); so do we ----
void set_error(const char *fmt, ...);
[[noreturn]] void fatal(const char *fmt, ...);

#define CUDA_CHECK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            ::damar::fatal("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,      \
                           __LINE__, cudaGetErrorString(_e));                                  \
    } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

constexpr int NUM_SMS_B200 = 148;
constexpr int PACK_FRONT_PAD = 4;     // words of zero padding in front of every packed base array
constexpr int RTAB_SHIFT = 9;         // one read-lookup entry per 512 base slots (reads are mostly longer than that)

// ---- device memory: caching allocator (api.cu); all work is ordered on the context stream ----
// Device memory that the driver will not give even after every worker's cache was handed back and a wait: thrown by
// dev_alloc_bytes, caught at the C-ABI (api.cu: the call is run again alone on the device, then reported as an error).
struct OutOfDeviceMemory { size_t bytes; };
void *dev_alloc_bytes(size_t bytes);
void dev_free(void *p);
// raw device blocks a function holds outside a DevBuf: released if the function is left by an exception
struct FreeOnUnwind {
    void *slot[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; int n = 0; bool armed = true;
    template <typename T> T *add(T *p) { slot[n++] = p; return p; }
    void dismiss() { armed = false; }
    ~FreeOnUnwind() { if (armed) for (int i = 0; i < n; i++) if (slot[i]) dev_free(slot[i]); }
};
// runs fn() -> int under the library's admission rule: normally alongside the other workers' calls; if it runs out of device
// memory, once more ALONE (after the calls in flight have finished and every cache was released); then `oom_rc` with the error set
int run_with_memory_retry(const char *what, int oom_rc, const std::function<int()> &fn);
template <typename T> inline T *dev_alloc(size_t n) { return reinterpret_cast<T *>(dev_alloc_bytes(n * sizeof(T))); }

// ---- simple owning device buffer ----
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;      // capacity in elements
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) dev_free(p);
        p = nullptr; n = 0;
    }
    // grow-only (contents are NOT preserved)
    void reserve(size_t want) {
        if (want <= n) return;
        release();
        size_t cap = want + want / 8 + 256;
        p = dev_alloc<T>(cap);
        n = cap;
    }
    T *steal() { T *q = p; p = nullptr; n = 0; return q; }
    void take(DevBuf &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
};

// ---- grow-only pinned host buffer (D2H / H2D staging at full PCIe rate) ----
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t n = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    void reserve(size_t want) {
        if (want <= n) return;
        if (p) cudaFreeHost(p);
        size_t cap = want + want / 4 + 4096;
        CUDA_CHECK(cudaHostAlloc((void **)&p, cap * sizeof(T), cudaHostAllocDefault));
        n = cap;
    }
};

// ---- filter parameters (Set_Filter_Params, filter.c:171-201) ----
struct FilterParams {
    int kmer = 14, binshift = 6, suppress = 0, hitmin = 35, nthreads = 4;
    int nthreads_pow2 = 4, nshift = 2;
    uint64_t kmask = (1ull << 28) - 1;
    bool set = false;
    bool biased = false;                        // BIASED (-b, daligner.c:729)
};
// LogBase of filter.c:774-789: fixed by the FIRST Sort_Kmers call after Set_Filter_Params (api.cu)
void bias_log_base(const float freq[4], int out[4]);

// Host waits.  With several ranks per host (one process per GPU, a few worker threads each) spinning waits would
// oversubscribe the cores, and even one rank's four spinning workers slowed the end-to-end pass (1220 vs 1120 ms on
// C2, the main thread copies blocks meanwhile), so the waits block: DAMAR_BLOCKING_SYNC=0 spins instead.
bool blocking_sync();
cudaError_t stream_wait(cudaStream_t st);
inline unsigned timing_event_flags() { return blocking_sync() ? cudaEventBlockingSync : cudaEventDefault; }

struct Timer {
    cudaEvent_t a, b;
    cudaStream_t s;
    explicit Timer(cudaStream_t st) : s(st) {
        CUDA_CHECK(cudaEventCreateWithFlags(&a, timing_event_flags())); CUDA_CHECK(cudaEventCreateWithFlags(&b, timing_event_flags()));
        CUDA_CHECK(cudaEventRecord(a, s));
    }
    float lap() { return stop(); }      // elapsed so far; the timer keeps its start
    float stop() {
        float ms = 0;
        CUDA_CHECK(cudaEventRecord(b, s)); CUDA_CHECK(cudaEventSynchronize(b));
        CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
        return ms;
    }
    ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

// ---- host wall-clock profile (DAMAR_PROFILE=1 prints it at damar_shutdown / damar_dump_profile) ----
struct HostProf {
    const char *name; double t0;
    explicit HostProf(const char *n);
    ~HostProf();
};
void dump_profile();

// ---- per-worker context ----
// Independent comparisons (block pair x orientation) may run concurrently: each host thread binds one
// worker context (damar_worker_bind) with its own stream, device-memory cache, pinned staging and result
// slots; context 0 is the default one.  Filter parameters and options are process-wide and copied into
// the context at every API entry.
struct SeedHits;
struct HostOverlaps;
struct WorkerScratch;
struct Context {
    int index = 0;
    int device = -1;
    cudaStream_t stream = nullptr;
    FilterParams fp;
    damar_options opt{};
    damar_timings tm{};
    int64_t launches = 0;
    bool ready = false;
    cudaEvent_t wait_ev = nullptr;               // blocking-sync event of stream_wait()
    cudaMemPool_t pool = nullptr;                // this worker's stream-ordered device memory pool (api.cu)
    std::mutex cache_mu;                         // guards free_blocks: another worker may trim it under memory pressure
    std::multimap<size_t, void *> free_blocks;   // caching allocator (api.cu)
    WorkerScratch *ws = nullptr;                 // buffers that outlive a call (api.cu)
};
Context &ctx();
void ensure_init();
inline void count_launch(int n = 1) { ctx().launches += n; }

// ---- device-side block (stage 0) ----
// gpos = index into the reference's unpacked `bases` array: read i occupies
// [boff[i], boff[i]+rlen[i]) and position boff[i]+rlen[i] is its terminator slot.
struct DeviceBlock {
    int32_t nreads = 0, maxlen = 0, ufirst = 0, part = 0;
    int64_t totlen = 0, nbases = 0;             // nbases = boff[nreads]
    int64_t sizeof_db = 0;
    float freq[4] = {0, 0, 0, 0};
    uint32_t *packed[2] = {nullptr, nullptr};   // 2 bit/base, 16 bases per word, [0]=forward [1]=revcomp
    int64_t npacked_words = 0;
    uint32_t *boff = nullptr;                   // [nreads+1] (blocks are < 2^31 bases, filter.c:794)
    int32_t *rlen = nullptr;                    // [nreads]
    uint32_t *rtab = nullptr;                   // read holding gpos (i << RTAB_SHIFT): starting point of read_of_tab()
    bool has_mask = false;
    bool as_is = false;                         // host already complemented the block (DAMAR_BLOCK_AS_IS)
    int64_t *mask_anno[2] = {nullptr, nullptr}; // [nreads+1], forward / mirrored
    int32_t *mask_data[2] = {nullptr, nullptr};
    int64_t mask_len = 0;
    std::vector<int64_t> h_boff;                // host copies (small)
    std::vector<int32_t> h_rlen;
    ~DeviceBlock();
};

// ---- device-side sorted k-mer list (stage 1 result) ----
struct DeviceIndex {
    DeviceBlock *blk = nullptr;
    int comp = 0;
    int64_t len = 0;
    // One 64-bit word per k-mer, code << gbits | gpos, whenever 2k + gbits <= 64 (gpos == nullptr then); else the
    // code in `codes` and the position in `gpos`.  Ascending by code, stable: ties in gpos order.
    uint64_t *codes = nullptr;
    uint32_t *gpos = nullptr;    // position of the k-mer's LAST base (rpos = gpos - boff[read])
    int gbits = 0;               // width of the gpos field of a packed entry
    ~DeviceIndex();
};

// ---- stage entry points (implemented in the .cu files) ----
// radix_sort.cu : stable LSD radix sort of (u64 key, u32 value) on bits [0, nbits)
struct SortScratch;
void radix_sort_pairs(uint64_t *&keys, uint32_t *&vals, uint64_t *&keys_alt, uint32_t *&vals_alt,
                      int64_t n, int nbits, cudaStream_t st);
// general form: key bits [lo_bit, hi_bit), vals / vals_alt may be null (keys-only); returns non-zero with the error set
int radix_sort(uint64_t *&keys, uint32_t *&vals, uint64_t *&keys_alt, uint32_t *&vals_alt, int64_t n, int lo_bit, int hi_bit,
               cudaStream_t st);
int sort_selftest(int64_t n, int lo_bit, int hi_bit, int has_vals, int reps, uint64_t seed, int skew, double *ms_out);
// scan.cu
void exclusive_scan_u32_to_u64(const uint32_t *in, uint64_t *out, int64_t n, uint64_t *total_dev, cudaStream_t st);
void exclusive_scan_u32_inplace(uint32_t *data, int64_t n, uint32_t *total_dev, cudaStream_t st);

// kmer_index.cu
DeviceBlock *upload_block(const damar_block_desc *d);
void block_view(DeviceBlock *b, damar_block_view *v);
DeviceBlock *block_from_view(const damar_block_view *v);
void ensure_revcomp(DeviceBlock *b);
DeviceIndex *build_index(DeviceBlock *b, int comp);

// seed_merge.cu
struct SeedHits {
    // key = (bread << (abits+pbits)) | (aread << pbits) | apos with bpos beside it, or -- bpos == nullptr -- ONE word
    // per hit: key << qbits | bpos
    uint64_t *keys = nullptr;
    uint32_t *bpos = nullptr;
    int64_t n = 0;
    int pbits = 0, abits = 0, bbits = 0, qbits = 0;
    int32_t limit = 0;
    int64_t total = 0;           // hits counted (== n unless the call stopped at the count: too many for one sort)
    void release();
};
struct HitChunk { uint32_t g_lo, g_hi; int32_t limit; };      // B positions [g_lo, g_hi) of a comparison cut into chunks of B reads
// 0: ok; 1: error set; 2: more than max_hits hits (nothing emitted, out.total / out.limit say what was counted)
int seed_hits(const DeviceIndex *ai, const DeviceIndex *bi, int self, int comp, SeedHits &out, const HitChunk *chunk = nullptr,
              int64_t max_hits = INT64_MAX);

// tandem.cu : datander's seed stage -- links between equal k-mers of one read as a hit list (scrub/tandem.c:556-590, 1281)
int self_hits(const DeviceIndex *idx, SeedHits &out);

// pair_align.cu
struct DeviceResult;
struct HostOverlaps {                 // views of pinned staging buffers, valid until the next call
    const damar_overlap *ovl = nullptr;
    const uint64_t *toff = nullptr, *seq = nullptr;
    const void *trace = nullptr;
    int64_t novl = 0, ntrace = 0;
    int tbytes = 2;
    int64_t nfilt = 0, multi = 0, cells = 0, pairs = 0, redo = 0;
    DeviceResult *keep = nullptr;     // set by the caller: leave the records on the device instead of downloading them
};
// One Match_Filter call's records left in HBM for the device-side writer (las_device.cu)
struct DeviceResult {
    DevBuf<damar_overlap> recs; DevBuf<uint64_t> toff, seq; DevBuf<uint8_t> trace, dead;
    uint64_t nrec = 0, ntrace = 0; int tbytes = 0; bool have_dead = false; int ordinal = 0;
};
struct DeviceSink;
void dsink_add(DeviceSink *s, DeviceResult *r);

// buffers that used to be function statics, one set per worker context
struct WorkerScratch {
    DevBuf<uint32_t> tile_hist;                  // radix_sort.cu
    cudaEvent_t rs_ev[16][2]; bool rs_ev_init = false;
    ~WorkerScratch() {
        if (rs_ev_init) for (auto &e : rs_ev) { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); }
        for (auto &e : up_ev) if (e) cudaEventDestroy(e);
    }
    DevBuf<uint64_t> scan_tot, scan_sums;        // scan.cu
    DevBuf<uint32_t> tot32_index, tot32_merge;   // kmer_index.cu, seed_merge.cu
    PinnedBuf<uint8_t> h_up[2]; cudaEvent_t up_ev[2] = {nullptr, nullptr};   // upload_block: pinned double buffer
    PinnedBuf<damar_overlap> h_recs; PinnedBuf<uint64_t> h_toff, h_seq; PinnedBuf<uint8_t> h_trace;   // pair_align.cu
    HostOverlaps out;                            // results of the last match_filter call
    std::vector<damar_overlap> acc_recs; std::vector<uint64_t> acc_toff, acc_seq; std::vector<uint8_t> acc_trace;   // ... of one cut into chunks (api.cu)
    SeedHits hits;
    const DeviceBlock *hits_a = nullptr, *hits_b = nullptr;
};
inline WorkerScratch &scratch() { return *ctx().ws; }

// trace_pts.cu
int trace_point_diffs_host(DeviceBlock *ab, DeviceBlock *bb, const damar_overlap *ovl, const uint64_t *toff, const void *trace,
                           int tbytes, int64_t novl, int64_t ntrace, int TS, int32_t *rec_diffs, uint16_t *seg_diffs,
                           int64_t *script_off = nullptr, int32_t **script_out = nullptr, int64_t *script_len = nullptr,
                           int flavour = 0);   // 0: Compute_Trace_PTS, 1: Compute_Trace_MID

// span: the hits are chunk [base, base + hits.n) of a comparison with `total` hits whose earlier chunks held pair_base work pairs
struct HitSpan { int64_t base = 0, total = 0; uint64_t pair_base = 0; };
int report_and_align(const DeviceBlock *ab, const DeviceBlock *bb, int comp, int self, const SeedHits &hits,
                     HostOverlaps &out, bool tandem = false, const HitSpan *span = nullptr);   // non-zero: error set (capacity limits are errors, not exits)

}  // namespace damar
