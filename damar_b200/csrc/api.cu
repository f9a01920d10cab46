// api.cu -- the C-ABI of include/damar_b200.h on top of the stage implementations.
#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>

#include "common.cuh"

namespace damar {

// one message per host thread: workers fail independently, and damar_last_error() hands out a pointer
static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_last_error = buf;
    fprintf(stderr, "damar_b200: %s\n", buf);
}
void fatal(const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    fprintf(stderr, "damar_b200: fatal: %s\n", buf);
    fflush(stderr);
    exit(1);                                   // reference behaviour: message + exit(1), db/DB.h:84-86
}

// ---- worker contexts ----
constexpr int MAX_WORKERS = 13;
static Context g_ctx[MAX_WORKERS];
static thread_local Context *tl_ctx = &g_ctx[0];
static std::mutex g_mu;                           // guards the process-wide settings, size map and profile
static FilterParams g_fp;
static damar_options g_opt{};
static bool g_opt_set = false;
static bool g_bias_latched = false;
static int g_log_base[4];
static int g_device = -1;
Context &ctx() { return *tl_ctx; }

static std::map<std::string, std::pair<double, long>> g_prof;
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
HostProf::HostProf(const char *n) : name(n), t0(now_s()) {}
// DAMAR_TRACE=<file>: every profiled scope also leaves a line "thread name start_ms end_ms" (written at exit): the
// poor man's timeline of which worker was where
struct TraceRec { int tid; const char *name; double t0, t1; };
static std::vector<TraceRec> g_trace;
static const char *g_trace_path = nullptr;
static thread_local int tl_tid = -1;
static int g_next_tid = 0;
HostProf::~HostProf() {
    const double t1 = now_s();
    std::lock_guard<std::mutex> lock(g_mu);
    auto &e = g_prof[name]; e.first += t1 - t0; e.second++;
    static const bool tracing = [] {
        g_trace_path = getenv("DAMAR_TRACE");
        if (!g_trace_path || !*g_trace_path) return false;
        atexit([] {
            if (FILE *f = fopen(g_trace_path, "w")) {
                for (auto &r : g_trace) fprintf(f, "%d %s %.3f %.3f\n", r.tid, r.name, r.t0 * 1e3, r.t1 * 1e3);
                fclose(f);
            }
        });
        return true;
    }();
    if (tracing) { if (tl_tid < 0) tl_tid = g_next_tid++; g_trace.push_back(TraceRec{tl_tid, name, t0, t1}); }
}
void dump_profile() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &kv : g_prof) fprintf(stderr, "[damar profile] %-28s %9.1f ms  (%ld calls)\n", kv.first.c_str(), kv.second.first * 1e3, kv.second.second);
    g_prof.clear();
}

static bool pool_alloc();
static cudaError_t make_pool(Context &c);
static int pick_device() {
    const char *e = getenv("DAMAR_DEVICE");
    if (e && *e) return atoi(e);
    e = getenv("LOCAL_RANK");
    if (e && *e) return atoi(e);
    return 0;
}

static int do_init(int device) {
    Context &c = ctx();
    if (c.ready) return 0;
    const double t_init = now_s();
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_device < 0) {
        static bool hooked = false;
        if (!hooked && getenv("DAMAR_PROFILE")) { hooked = true; atexit([] { dump_profile(); }); }   // hosts that never call damar_shutdown
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0) {
            set_error("no usable CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
            return 1;
        }
        if (device < 0) device = pick_device();
        if (device >= n) device = device % n;
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) {
            set_error("device %d (%s, sm_%d%d) is not a Blackwell sm_100 part", device, prop.name, prop.major, prop.minor);
            return 1;
        }
        g_device = device;
        if (!g_opt_set) {                         // defaults of daligner.c:688-703
            g_opt.minover = 2000; g_opt.symmetric = 1; g_opt.trace_spacing = 100; g_opt.ave_corr = .70;
            g_opt.freq[0] = g_opt.freq[1] = g_opt.freq[2] = g_opt.freq[3] = .25f;
        }
    }
    CUDA_CHECK(cudaSetDevice(g_device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    if (pool_alloc()) CUDA_CHECK(make_pool(c));
    c.index = (int)(&c - g_ctx);
    c.device = g_device;
    if (!c.ws) c.ws = new WorkerScratch();
    c.fp = g_fp; c.opt = g_opt;
    c.ready = true;
    { auto &e = g_prof["init (CUDA context, stream)"]; e.first += now_s() - t_init; e.second++; }
    return 0;
}

// ---- caching device allocator --------------------------------------------------------------------
// A block pair allocates and frees tens of GB of tuple / hit lists of data-dependent sizes.  cudaMalloc costs ~70 ms per
// GB and cudaFree waits for every running kernel (0.3 s behind a wave kernel, measured: tools/exp/malloc_stall.cu).  All
// work of a worker context is ordered on its stream, so a freed block can be handed to the context's next request
// immediately: requests are rounded up to a coarse size class and served best-fit from the context's free list (at most
// 1.5x the rounded size).  A block freed by ANOTHER context than the one that allocated it (the collector releasing a
// job's records, a plan dropping its k-mer indexes) comes with the caller's guarantee that no stream still uses it: it
// goes to one shared pool that every context draws from before it asks the driver -- the lists of a plan circulate
// between the workers (an index is built by whichever worker needs it first) instead of piling up where they were last
// freed.  DAMAR_SHARED_POOL=0 returns such a block to the cache of the context that allocated it instead.
// DAMAR_ALLOC=pool selects per-worker stream-ordered pools (cudaMemPool) instead: measured slower -- while the pools grow
// every request waits for the driver to map memory (a first pass of 9 s instead of 0.9 s) and the steady state is the same.
struct BlockInfo { size_t size; int owner; };          // owner = the worker context whose request allocated the block
static std::map<void *, BlockInfo> g_block_size;
static std::mutex g_shared_mu;
static std::multimap<size_t, void *> g_shared_blocks;  // quiescent blocks any context may take
static bool shared_pool() {
    static const bool on = !(getenv("DAMAR_SHARED_POOL") && atoi(getenv("DAMAR_SHARED_POOL")) == 0);
    return on;
}
static bool pool_alloc() {
    static const bool on = [] { const char *e = getenv("DAMAR_ALLOC"); return e && strcmp(e, "pool") == 0; }();
    return on;
}
static size_t round_class(size_t b) {
    size_t q = b >= (256u << 20) ? (64u << 20) : b >= (16u << 20) ? (8u << 20) : b >= (1u << 20) ? (1u << 20) : (64u << 10);
    return (b + q - 1) / q * q;
}
static cudaError_t make_pool(Context &c) {
    cudaMemPoolProps pr{};
    pr.allocType = cudaMemAllocationTypePinned; pr.handleTypes = cudaMemHandleTypeNone;
    pr.location.type = cudaMemLocationTypeDevice; pr.location.id = g_device;
    cudaError_t e = cudaMemPoolCreate(&c.pool, &pr);
    if (e != cudaSuccess) return e;
    unsigned long long keep = ~0ull;
    int off = 0;
    cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &keep);
    // never make this stream wait for another worker's stream in order to reuse a block freed there
    cudaMemPoolSetAttribute(c.pool, cudaMemPoolReuseAllowInternalDependencies, &off);
    return cudaSuccess;
}
static void release_cached_blocks(Context &c) {
    if (c.pool) { cudaMemPoolTrimTo(c.pool, 0); return; }
    std::lock_guard<std::mutex> lk(c.cache_mu);
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &kv : c.free_blocks) { g_block_size.erase(kv.second); cudaFree(kv.second); }
    c.free_blocks.clear();
}
// Memory pressure between workers: a worker whose request fails hands the unused memory of EVERY worker's pool back to the
// driver and then retries for a while -- the other comparisons in flight finish and release their lists.
static void release_shared_blocks() {
    std::lock_guard<std::mutex> lk(g_shared_mu);
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &kv : g_shared_blocks) { g_block_size.erase(kv.second); cudaFree(kv.second); }
    g_shared_blocks.clear();
}
static void release_all_caches() {
    for (int i = 0; i < MAX_WORKERS; i++) release_cached_blocks(g_ctx[i]);
    release_shared_blocks();
}
void *dev_alloc_bytes(size_t bytes) {
    Context &c = ctx();
    void *p = nullptr;
    if (pool_alloc() && c.pool) {
        const size_t r = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
        cudaError_t e = cudaMallocFromPoolAsync(&p, r, c.pool, c.stream);
        if (e != cudaSuccess) {
            HostProf hp("alloc.retry");
            cudaGetLastError();
            CUDA_CHECK(stream_wait(c.stream));
            for (int attempt = 0;; attempt++) {
                release_all_caches();
                e = cudaMallocFromPoolAsync(&p, r, c.pool, c.stream);
                if (e == cudaSuccess) break;
                cudaGetLastError();
                if (attempt >= 40) throw OutOfDeviceMemory{r};
                std::this_thread::sleep_for(std::chrono::milliseconds(50));
            }
        }
        return p;
    }
    const size_t r = round_class(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> lk(c.cache_mu);
        auto it = c.free_blocks.lower_bound(r);
        if (it != c.free_blocks.end() && it->first <= r + r / 2) {
            p = it->second;
            c.free_blocks.erase(it);
            return p;
        }
    }
    {
        std::lock_guard<std::mutex> lk(g_shared_mu);
        auto it = g_shared_blocks.lower_bound(r);
        if (it != g_shared_blocks.end() && it->first <= r + r / 2) {
            p = it->second;
            g_shared_blocks.erase(it);
            std::lock_guard<std::mutex> lock(g_mu);
            g_block_size[p].owner = c.index;
            return p;
        }
    }
    // DAMAR_FAKE_OOM=n (tests): the n-th request of the process that reaches the driver fails once
    static const long fake_at = getenv("DAMAR_FAKE_OOM") ? atol(getenv("DAMAR_FAKE_OOM")) : 0;
    static std::atomic<long> misses{0};
    if (fake_at > 0 && ++misses == fake_at) throw OutOfDeviceMemory{r};
    HostProf hp("cudaMalloc");
    cudaError_t e = cudaMalloc(&p, r);
    if (e != cudaSuccess) {
        // the other workers' cached blocks first, then a short wait for comparisons in flight to hand their lists back; after
        // that the caller is unwound (its own buffers return to the cache) and runs again alone: run_with_memory_retry()
        cudaGetLastError();
        CUDA_CHECK(stream_wait(c.stream));
        for (int attempt = 0;; attempt++) {
            release_all_caches();
            e = cudaMalloc(&p, r);
            if (e == cudaSuccess) break;
            cudaGetLastError();
            if (attempt >= 40) throw OutOfDeviceMemory{r};
            std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
    }
    std::lock_guard<std::mutex> lock(g_mu);
    g_block_size[p] = BlockInfo{r, c.index};
    return p;
}
void dev_free(void *p) {
    Context &c = ctx();
    if (!p) return;
    if (!c.ready) {                                // after damar_shutdown: no stream / cache to return it to
        { std::lock_guard<std::mutex> lock(g_mu); g_block_size.erase(p); }
        cudaFree(p);
        return;
    }
    if (pool_alloc()) { cudaFreeAsync(p, c.stream); return; }      // goes back to the pool it was carved from
    size_t sz = 0; int owner = c.index;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = g_block_size.find(p);
        if (it != g_block_size.end()) { sz = it->second.size; owner = it->second.owner; }
    }
    if (sz == 0) { cudaFree(p); return; }
    if (owner != c.index && shared_pool()) {       // freed by another context: quiescent by the caller's guarantee
        std::lock_guard<std::mutex> lk(g_shared_mu);
        g_shared_blocks.emplace(sz, p);
        return;
    }
    Context &o = (owner >= 0 && owner < MAX_WORKERS && g_ctx[owner].ready) ? g_ctx[owner] : c;
    std::lock_guard<std::mutex> lk(o.cache_mu);
    o.free_blocks.emplace(sz, p);
}

// Admission under memory pressure.  Calls that allocate hold g_solo shared; one that ran out of memory has been unwound
// (everything it held is back in its worker's cache), takes g_solo exclusively -- i.e. waits for the calls in flight --
// gives every cache back to the driver and runs again with the device to itself.  Workers that wait here hold no device
// memory, so the retry cannot deadlock; only a call that does not fit the device ALONE is an error.
static std::shared_mutex g_solo;
int run_with_memory_retry(const char *what, int oom_rc, const std::function<int()> &fn) {
    try {
        std::shared_lock<std::shared_mutex> lk(g_solo);
        return fn();
    } catch (const OutOfDeviceMemory &) {
    }
    std::unique_lock<std::shared_mutex> lk(g_solo);
    HostProf hp("memory.solo_rerun");
    cudaDeviceSynchronize();
    release_all_caches();
    try {
        return fn();
    } catch (const OutOfDeviceMemory &e) {
        set_error("%s: out of device memory (%zu bytes requested with the device to itself)", what, e.bytes);
        return oom_rc;
    }
}

void ensure_init() {
    Context &c = ctx();
    if (!c.ready && do_init(-1) != 0) fatal("%s", g_last_error.c_str());
    std::lock_guard<std::mutex> lock(g_mu);
    c.fp = g_fp; c.opt = g_opt;
}

bool blocking_sync() {
    static const bool on = [] {
        const char *e = getenv("DAMAR_BLOCKING_SYNC");
        return (e && *e) ? atoi(e) != 0 : true;
    }();
    return on;
}
cudaError_t stream_wait(cudaStream_t st) {
    HostProf hp("wait");
    if (!blocking_sync()) return cudaStreamSynchronize(st);
    Context &c = ctx();
    if (!c.wait_ev) { cudaError_t e = cudaEventCreateWithFlags(&c.wait_ev, cudaEventBlockingSync | cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    cudaError_t e = cudaEventRecord(c.wait_ev, st);
    return e != cudaSuccess ? e : cudaEventSynchronize(c.wait_ev);
}

// filter.c:774-789: NormShift/LogBase are built once per process, from the freq of the block of the first
// Sort_Kmers call (the forward A block, daligner.c:982) and reused for every later block and orientation.
void bias_log_base(const float freq[4], int out[4]) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_bias_latched) {
        const double scale = -10000. / log(4.);
        for (int i = 0; i < 4; i++) g_log_base[i] = (int)ceil(scale * log(freq[i]));
        g_bias_latched = true;
    }
    for (int i = 0; i < 4; i++) out[i] = g_log_base[i];
}

// ---- results of the last match_filter call live in the worker's scratch ----
#define g_out (scratch().out)
#define g_hits (scratch().hits)
#define g_hits_a (scratch().hits_a)
#define g_hits_b (scratch().hits_b)

void SeedHits::release() {
    if (keys) dev_free(keys);
    if (bpos) dev_free(bpos);
    keys = nullptr; bpos = nullptr; n = 0;
}

}  // namespace damar

using namespace damar;

extern "C" {

int damar_abi_version(void) { return DAMAR_B200_ABI_VERSION; }
int damar_init(int device) { return do_init(device); }
const char *damar_last_error(void) { return g_last_error.c_str(); }

void damar_shutdown(void) {
    Context *saved = tl_ctx;
    for (int i = 0; i < MAX_WORKERS; i++) {
        Context &c = g_ctx[i];
        if (!c.ready) continue;
        tl_ctx = &c;
        c.ws->hits.release();
        cudaStreamSynchronize(c.stream);
        delete c.ws; c.ws = nullptr;                 // scratch buffers go back to the cache, pinned staging and events are released
        release_cached_blocks(c);
        if (c.pool) { cudaMemPoolDestroy(c.pool); c.pool = nullptr; }     // blocks still held by the host stay valid until freed
        if (c.wait_ev) { cudaEventDestroy(c.wait_ev); c.wait_ev = nullptr; }
        cudaStreamDestroy(c.stream);
        c.stream = nullptr;
        c.ready = false;
    }
    tl_ctx = saved;
    release_shared_blocks();
    g_device = -1;
}

/* Binds the calling host thread to worker context `index` (0 = the default context).  Every context has its
 * own stream, device-memory cache and result slots, so independent comparisons can be in flight at once. */
int damar_worker_bind(int index) {
    if (index < 0 || index >= MAX_WORKERS) return 1;
    tl_ctx = &g_ctx[index];
    const int rc = do_init(-1);
    if (rc == 0) CUDA_CHECK(cudaSetDevice(g_device));        // the context may have been created by another thread
    return rc;
}
int damar_worker_count(void) { return MAX_WORKERS; }

int damar_set_filter_params(int kmer, int binshift, int suppress, int hitmin, int nthreads) {
    if (kmer <= 1 || kmer > 32) return 1;                     // filter.c:173-174 (+ u64 code limit)
    std::lock_guard<std::mutex> lock(g_mu);
    FilterParams &fp = g_fp;
    fp.kmer = kmer; fp.binshift = binshift; fp.suppress = suppress; fp.hitmin = hitmin; fp.nthreads = nthreads;
    fp.kmask = (kmer == 32) ? ~0ull : ((1ull << (2 * kmer)) - 1);
    fp.nthreads_pow2 = 1; fp.nshift = 0;
    while (2 * fp.nthreads_pow2 <= nthreads) { fp.nthreads_pow2 *= 2; fp.nshift++; }   // filter.c:192-198
    fp.set = true;
    g_bias_latched = false;
    ctx().fp = fp;
    return 0;
}

/* BIASED (daligner.c:132, set by -b).  Call before the first damar_sort_kmers of a job. */
void damar_set_biased(int biased) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_fp.biased = biased != 0;
    g_bias_latched = false;
    ctx().fp.biased = g_fp.biased;
}

void damar_set_options(const damar_options *opt) { std::lock_guard<std::mutex> lock(g_mu); g_opt = *opt; g_opt_set = true; ctx().opt = *opt; }

damar_block *damar_block_upload(const damar_block_desc *desc) {
    ensure_init();
    HostProf hp("block_upload");
    DeviceBlock *b = nullptr;
    if (run_with_memory_retry("damar_block_upload", 1, [&] { b = upload_block(desc); return 0; }) != 0) return nullptr;
    return reinterpret_cast<damar_block *>(b);
}
void damar_block_free(damar_block *blk) { delete reinterpret_cast<DeviceBlock *>(blk); }
int damar_block_view_get(damar_block *blk, damar_block_view *view) {
    if (blk == nullptr || view == nullptr) return 1;
    block_view(reinterpret_cast<DeviceBlock *>(blk), view);
    return 0;
}
damar_block *damar_block_from_view(const damar_block_view *view) {
    ensure_init();
    if (view == nullptr || view->packed == nullptr || view->boff == nullptr || view->rlen == nullptr) { set_error("damar_block_from_view: incomplete view"); return nullptr; }
    HostProf hp("block_from_view");
    DeviceBlock *b = nullptr;
    if (run_with_memory_retry("damar_block_from_view", 1, [&] { b = block_from_view(view); return 0; }) != 0) return nullptr;
    return reinterpret_cast<damar_block *>(b);
}

damar_index *damar_sort_kmers(damar_block *blk, int comp, int64_t *len) {
    ensure_init();
    HostProf hp("sort_kmers");
    DeviceIndex *idx = nullptr;
    if (run_with_memory_retry("damar_sort_kmers", 1, [&] { idx = build_index(reinterpret_cast<DeviceBlock *>(blk), comp); return 0; }) != 0) {
        if (len) *len = -1;                                   // NULL with *len == 0 means "no k-mers" (filter.c:990-993); -1: failed, error set
        return nullptr;
    }
    if (len) *len = idx->len;
    ctx().tm.kernel_launches = ctx().launches;
    if (idx->len == 0) { delete idx; return nullptr; }        // filter.c:990-993
    return reinterpret_cast<damar_index *>(idx);
}
void damar_index_free(damar_index *idx) { delete reinterpret_cast<DeviceIndex *>(idx); }

// One sort takes fewer than 2^30 elements and the alignment stage indexes hits with 32 bits; the reference's hit array has no
// such bound (int64 counts, filter.c:2529).  A comparison with more hits is cut into chunks of B reads: hits are ordered by
// B read first, so the chunks are consecutive ranges of the reference's hit array and every read pair lies in one chunk.
// DAMAR_MAX_HITS lowers the bound (the tests force dozens of chunks on small fixtures).
static int64_t max_hits_per_sort() {
    static const int64_t v = [] { const char *e = getenv("DAMAR_MAX_HITS"); const int64_t x = (e && *e) ? atoll(e) : 0; return x > 0 ? x : ((int64_t)1 << 30) - 4096; }();
    return v;
}

static int match_filter_impl(damar_block *ablk, damar_block *bblk, damar_index *aidx, damar_index *bidx,
                             int self, int comp, damar_overlaps *out, DeviceSink *sink, int ordinal) {
    ensure_init();
    Context &c = ctx();
    DeviceBlock *ab = reinterpret_cast<DeviceBlock *>(ablk), *bb = reinterpret_cast<DeviceBlock *>(bblk);
    DeviceIndex *ai = reinterpret_cast<DeviceIndex *>(aidx), *bi = reinterpret_cast<DeviceIndex *>(bidx);
    g_out = HostOverlaps();
    g_hits.release();
    g_hits_a = ab; g_hits_b = bb;
    memset(out, 0, sizeof(*out));
    c.tm.count_ms = c.tm.merge_ms = c.tm.hsort_ms = c.tm.pairs_ms = c.tm.align_ms = c.tm.download_ms = 0;
    c.tm.hits = c.tm.pairs = c.tm.la_calls = c.tm.wave_cells = 0; c.tm.wide_passes = 0;
    int64_t nhits = 0;
    // stages 3 + 4 on the hits in g_hits; with a sink every call (chunk) leaves its own result, handed over when the whole
    // comparison has succeeded (a comparison that ran out of memory is run again from the start)
    struct Pending { std::vector<DeviceResult *> v; ~Pending() { for (auto *r : v) delete r; } } pending;
    auto align = [&](const HitSpan *span) -> int {
        DeviceResult *r = nullptr;
        if (sink) { r = new DeviceResult(); r->ordinal = ordinal; pending.v.push_back(r); }
        g_out.keep = r;
        int rc = 0;
        if (g_hits.n > 0) { HostProf hp("mf.report_and_align"); rc = report_and_align(ab, bb, comp, self, g_hits, g_out, false, span); }
        g_out.keep = nullptr;
        return rc;
    };
    if (ai != nullptr && bi != nullptr && ai->len > 0 && bi->len > 0) {   // filter.c:2592-2593
        ensure_revcomp(ab); ensure_revcomp(bb);               // the reverse wave pass runs on the reverse-complemented reads
        int rc;
        { HostProf hp("mf.seed_hits"); rc = seed_hits(ai, bi, self, comp, g_hits, nullptr, max_hits_per_sort()); }
        if (rc == 1) return 1;
        nhits = g_hits.total;
        if (rc == 0) { if (align(nullptr) != 0) return 1; }
        else {
            // ---- too many hits for one sort: chunks of B reads, in order ----
            HostProf hp("mf.chunked");
            const int64_t total = g_hits.total;
            const int32_t limit = g_hits.limit;
            const int nreads = bb->nreads;
            const std::vector<int64_t> &boff = bb->h_boff;
            std::vector<std::pair<int, int>> todo;            // ranges of B reads still to do, last = next
            {
                int64_t want = (total + max_hits_per_sort() / 2 - 1) / std::max<int64_t>(1, max_hits_per_sort() / 2);
                const int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>(nreads, want));
                std::vector<int> cut(1, 0);
                for (int k = 1; k < nchunk; k++) {            // equal shares of the bases, at read boundaries
                    const int64_t g = boff[nreads] / nchunk * k;
                    int r = (int)(std::upper_bound(boff.begin(), boff.begin() + nreads + 1, g) - boff.begin()) - 1;
                    if (r > cut.back()) cut.push_back(r);
                }
                cut.push_back(nreads);
                for (size_t k = cut.size() - 1; k >= 1; k--) todo.emplace_back(cut[k - 1], cut[k]);
            }
            WorkerScratch &ws = scratch();
            ws.acc_recs.clear(); ws.acc_toff.clear(); ws.acc_seq.clear(); ws.acc_trace.clear();
            HostOverlaps sum;
            HitSpan span; span.total = total;
            while (!todo.empty()) {
                const std::pair<int, int> rg = todo.back(); todo.pop_back();
                HitChunk ch{(uint32_t)boff[rg.first], (uint32_t)boff[rg.second], limit};
                { HostProf hp2("mf.seed_hits"); rc = seed_hits(ai, bi, self, comp, g_hits, &ch, max_hits_per_sort()); }
                if (rc == 1) return 1;
                if (rc == 2) {                                // still too many: halve the range of reads
                    if (rg.second - rg.first < 2) { set_error("one B read alone has %lld seed hits, more than one sort takes: lower -M or -t", (long long)g_hits.total); return 1; }
                    const int mid = rg.first + (rg.second - rg.first) / 2;
                    todo.emplace_back(mid, rg.second); todo.emplace_back(rg.first, mid);
                    continue;
                }
                g_out = HostOverlaps();
                if (align(&span) != 0) return 1;
                span.base += g_hits.n; span.pair_base += (uint64_t)g_out.pairs;
                sum.nfilt += g_out.nfilt; sum.multi += g_out.multi; sum.cells += g_out.cells; sum.pairs += g_out.pairs; sum.redo += g_out.redo;
                sum.novl += g_out.novl; sum.tbytes = g_out.tbytes;
                if (!sink && g_out.novl > 0) {                // the host path returns ONE set of arrays: append this chunk's
                    const uint64_t tbase = (uint64_t)sum.ntrace;
                    ws.acc_recs.insert(ws.acc_recs.end(), g_out.ovl, g_out.ovl + g_out.novl);
                    for (int64_t k = 0; k < g_out.novl; k++) ws.acc_toff.push_back(g_out.toff[k] + tbase);
                    ws.acc_seq.insert(ws.acc_seq.end(), g_out.seq, g_out.seq + g_out.novl);
                    const uint8_t *tp = (const uint8_t *)g_out.trace;
                    ws.acc_trace.insert(ws.acc_trace.end(), tp, tp + (size_t)g_out.ntrace * g_out.tbytes);
                }
                sum.ntrace += g_out.ntrace;
                g_hits.release();
            }
            if (span.base != total) { set_error("chunked comparison: %lld hits emitted, %lld counted", (long long)span.base, (long long)total); return 1; }
            if (!sink) { sum.ovl = ws.acc_recs.data(); sum.toff = ws.acc_toff.data(); sum.seq = ws.acc_seq.data(); sum.trace = ws.acc_trace.data(); }
            g_out = sum;
            g_hits.limit = limit;
        }
    }
    out->novl = g_out.novl;
    out->ovl = g_out.ovl;
    out->trace_off = g_out.toff;
    out->seq = g_out.seq;
    out->trace = g_out.trace;
    out->ntrace = g_out.ntrace;
    out->trace_bytes = g_out.tbytes;
    out->nhits = nhits;
    out->nfilt = g_out.nfilt;
    out->ncheck = out->novl;
    out->limit = g_hits.limit;
    out->multi_la_pairs = (int32_t)g_out.multi;
    c.tm.cells_total += g_out.cells; c.tm.wide_total += c.tm.wide_passes; c.tm.la_total += g_out.nfilt;
    c.tm.hits = nhits; c.tm.la_calls = g_out.nfilt; c.tm.wave_cells = g_out.cells; c.tm.pairs = g_out.pairs; c.tm.redo_pairs = g_out.redo;
    c.tm.kernel_launches = c.launches;
    for (auto *r : pending.v) dsink_add(sink, r);
    pending.v.clear();
    return 0;
}

int damar_match_filter(damar_block *ablk, damar_block *bblk, damar_index *aidx, damar_index *bidx,
                       int self, int comp, damar_overlaps *out) {
    return run_with_memory_retry("damar_match_filter", 1, [&] { return match_filter_impl(ablk, bblk, aidx, bidx, self, comp, out, nullptr, 0); });
}
/* Match_Filter whose records stay in HBM: they are appended to `sink` as call number `ordinal` of the job (the order in
 * which the reference would have buffered the calls breaks full SORT_OVL ties); `out` receives the counters only. */
int damar_match_filter_sink(damar_block *ablk, damar_block *bblk, damar_index *aidx, damar_index *bidx, int self, int comp,
                            damar_device_sink *sink, int ordinal, damar_overlaps *out) {
    if (sink == nullptr) { set_error("damar_match_filter_sink: no sink"); return 1; }
    return run_with_memory_retry("damar_match_filter_sink", 1, [&] {
        return match_filter_impl(ablk, bblk, aidx, bidx, self, comp, out, reinterpret_cast<DeviceSink *>(sink), ordinal);
    });
}

/* datander's Match_Self (scrub/tandem.c:1182): every read of the block against itself. */
static int match_self_impl(damar_block *blk, damar_overlaps *out) {
    ensure_init();
    Context &c = ctx();
    DeviceBlock *b = reinterpret_cast<DeviceBlock *>(blk);
    g_out = HostOverlaps();
    g_hits.release();
    g_hits_a = b; g_hits_b = b;
    memset(out, 0, sizeof(*out));
    c.tm.count_ms = c.tm.merge_ms = c.tm.hsort_ms = c.tm.pairs_ms = c.tm.align_ms = c.tm.download_ms = 0;
    c.tm.hits = c.tm.pairs = c.tm.la_calls = c.tm.wave_cells = 0; c.tm.wide_passes = 0;
    if (b == nullptr) { set_error("damar_match_self: no block"); return 1; }
    if (b->has_mask) { set_error("damar_match_self: datander takes no mask tracks"); return 1; }
    if (c.fp.suppress != 0 || c.fp.biased) { set_error("damar_match_self: datander has neither -t nor -b (scrub/tandem.c:137)"); return 1; }
    std::unique_ptr<DeviceIndex> idx;
    { HostProf hp("ms.sort_kmers"); idx.reset(build_index(b, 0)); }
    int rc = 0;
    if (idx->len > 0) {
        ensure_revcomp(b);                                    // the reverse wave pass runs on the reverse-complemented reads
        { HostProf hp("ms.self_hits"); rc = self_hits(idx.get(), g_hits); }
        if (rc == 0 && g_hits.n > 0) { HostProf hp("ms.report_and_align"); rc = report_and_align(b, b, 0, 1, g_hits, g_out, true); }
    }
    idx.reset();
    if (rc != 0) return rc;
    out->novl = g_out.novl; out->ovl = g_out.ovl; out->trace_off = g_out.toff; out->seq = g_out.seq; out->trace = g_out.trace;
    out->ntrace = g_out.ntrace; out->trace_bytes = g_out.tbytes; out->nhits = g_hits.n; out->nfilt = g_out.nfilt;
    out->ncheck = out->novl; out->limit = g_hits.limit; out->multi_la_pairs = (int32_t)g_out.multi;
    c.tm.cells_total += g_out.cells; c.tm.wide_total += c.tm.wide_passes; c.tm.la_total += g_out.nfilt;
    c.tm.hits = g_hits.n; c.tm.la_calls = g_out.nfilt; c.tm.wave_cells = g_out.cells; c.tm.pairs = g_out.pairs; c.tm.redo_pairs = g_out.redo;
    c.tm.kernel_launches = c.launches;
    return 0;
}

int damar_match_self(damar_block *blk, damar_overlaps *out) {
    return run_with_memory_retry("damar_match_self", 1, [&] { return match_self_impl(blk, out); });
}

int damar_trace_point_diffs(damar_block *ablk, damar_block *bblk, const damar_overlap *ovl, const uint64_t *trace_off,
                            const void *trace, int trace_bytes, int64_t novl, int64_t ntrace, int trace_spacing,
                            int32_t *rec_diffs, uint16_t *seg_diffs) {
    ensure_init();
    if (trace_spacing <= 0 || (trace_bytes != 1 && trace_bytes != 2)) { set_error("damar_trace_point_diffs: bad trace spacing / width"); return 1; }
    return run_with_memory_retry("damar_trace_point_diffs", 1, [&] {
        return trace_point_diffs_host(reinterpret_cast<DeviceBlock *>(ablk), reinterpret_cast<DeviceBlock *>(bblk), ovl, trace_off, trace,
                                      trace_bytes, novl, ntrace, trace_spacing, rec_diffs, seg_diffs);
    });
}

int damar_trace_point_scripts(damar_block *ablk, damar_block *bblk, const damar_overlap *ovl, const uint64_t *trace_off,
                              const void *trace, int trace_bytes, int64_t novl, int64_t ntrace, int trace_spacing,
                              int32_t *rec_diffs, int64_t *script_off, int32_t **script, int64_t *script_len) {
    ensure_init();
    if (trace_spacing <= 0 || (trace_bytes != 1 && trace_bytes != 2) || !rec_diffs || !script_off || !script || !script_len) {
        set_error("damar_trace_point_scripts: bad arguments"); return 2;
    }
    return run_with_memory_retry("damar_trace_point_scripts", 2, [&] {
        return trace_point_diffs_host(reinterpret_cast<DeviceBlock *>(ablk), reinterpret_cast<DeviceBlock *>(bblk), ovl, trace_off, trace,
                                      trace_bytes, novl, ntrace, trace_spacing, rec_diffs, nullptr, script_off, script, script_len);
    });
}

int damar_trace_mid_scripts(damar_block *ablk, damar_block *bblk, const damar_overlap *ovl, const uint64_t *trace_off,
                            const void *trace, int trace_bytes, int64_t novl, int64_t ntrace, int trace_spacing,
                            int32_t *rec_diffs, int64_t *script_off, int32_t **script, int64_t *script_len) {
    ensure_init();
    if (trace_spacing <= 0 || (trace_bytes != 1 && trace_bytes != 2) || !rec_diffs || !script_off || !script || !script_len) {
        set_error("damar_trace_mid_scripts: bad arguments"); return 2;
    }
    return run_with_memory_retry("damar_trace_mid_scripts", 2, [&] {
        return trace_point_diffs_host(reinterpret_cast<DeviceBlock *>(ablk), reinterpret_cast<DeviceBlock *>(bblk), ovl, trace_off, trace,
                                      trace_bytes, novl, ntrace, trace_spacing, rec_diffs, nullptr, script_off, script, script_len, 1);
    });
}

void damar_host_free(void *p) { free(p); }

void damar_emission_order(const damar_overlaps *o, int64_t *order) {
    std::vector<int64_t> idx((size_t)o->novl);
    for (int64_t i = 0; i < o->novl; i++) idx[i] = i;
    if (o->seq) std::sort(idx.begin(), idx.end(), [&](int64_t x, int64_t y) { return o->seq[x] < o->seq[y]; });
    memcpy(order, idx.data(), sizeof(int64_t) * (size_t)o->novl);
}

// ---- stage read-backs for the parity tests ----
namespace {
__global__ void to_kmerpos(const uint64_t *__restrict__ codes, const uint32_t *__restrict__ gpos, int gbits, int64_t n,
                           const uint32_t *__restrict__ boff, int nreads, uint4 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = gpos ? gpos[i] : (uint32_t)(codes[i] & ((1ull << gbits) - 1ull));
    int lo = 0, hi = nreads;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (boff[mid] <= p) lo = mid; else hi = mid; }
    uint64_t c = gpos ? codes[i] : (codes[i] >> gbits);
    out[i] = make_uint4((uint32_t)c, (uint32_t)(c >> 32), p - boff[lo], (uint32_t)lo);
}
__global__ void to_seedpair(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ bpos, int64_t n, int pbits, int abits, int qbits,
                            int4 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = bpos ? keys[i] : (keys[i] >> qbits);
    const int bp = bpos ? (int)bpos[i] : (int)(keys[i] & ((1ull << qbits) - 1ull));
    int apos = (int)(k & ((1ull << pbits) - 1));
    int aread = (int)((k >> pbits) & ((1ull << abits) - 1));
    int bread = (int)(k >> (pbits + abits));
    out[i] = make_int4(apos - bp, apos, aread, bread);
}
}  // namespace

int damar_index_download(const damar_index *idx_, void *kmerpos_out, int64_t capacity) {
    const DeviceIndex *idx = reinterpret_cast<const DeviceIndex *>(idx_);
    if (idx == nullptr || idx->len > capacity) return 1;
    Context &c = ctx();
    DevBuf<uint4> tmp; tmp.reserve((size_t)idx->len);
    to_kmerpos<<<(unsigned)((idx->len + 255) / 256), 256, 0, c.stream>>>(idx->codes, idx->gpos, idx->gbits, idx->len, idx->blk->boff,
                                                                         idx->blk->nreads, tmp.p);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(kmerpos_out, tmp.p, sizeof(uint4) * idx->len, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(stream_wait(c.stream));
    return 0;
}

int64_t damar_last_hits_download(void *seedpair_out, int64_t capacity) {
    if (g_hits.n == 0) return 0;
    if (seedpair_out == nullptr || g_hits.n > capacity) return g_hits.n;
    Context &c = ctx();
    DevBuf<int4> tmp; tmp.reserve((size_t)g_hits.n);
    to_seedpair<<<(unsigned)((g_hits.n + 255) / 256), 256, 0, c.stream>>>(g_hits.keys, g_hits.bpos, g_hits.n, g_hits.pbits,
                                                                         g_hits.abits, g_hits.qbits, tmp.p);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(seedpair_out, tmp.p, sizeof(int4) * g_hits.n, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(stream_wait(c.stream));
    return g_hits.n;
}

int damar_sort_selftest(int64_t n, int lo_bit, int hi_bit, int has_vals, int reps, uint64_t seed, int skew, double *ms_out) {
    ensure_init();
    if (n < 0 || lo_bit < 0 || hi_bit > 64 || hi_bit < lo_bit) { set_error("damar_sort_selftest: bad arguments"); return 1; }
    return sort_selftest(n, lo_bit, hi_bit, has_vals, reps, seed, skew, ms_out);
}

void damar_dump_profile(void) { dump_profile(); }
void damar_get_timings(damar_timings *t) {
    *t = ctx().tm;                                            // per-call fields of the calling worker ...
    t->kernel_launches = 0; t->scatter_ms = 0; t->scatter_elems = 0; t->scatter_launches = 0; t->scatter_bytes = 0;
    t->sort_ms = 0; t->sort_contract_bytes = 0; t->sort_elems = 0; t->cells_total = 0; t->wide_total = 0; t->la_total = 0;
    t->align_kernel_ms = 0; t->align_launches = 0; t->redo_kernel_ms = 0;
    for (int i = 0; i < MAX_WORKERS; i++) {                   // ... accumulators summed over all workers
        const Context &c = g_ctx[i];
        t->kernel_launches += c.launches;
        t->scatter_ms += c.tm.scatter_ms; t->scatter_elems += c.tm.scatter_elems; t->scatter_launches += c.tm.scatter_launches; t->scatter_bytes += c.tm.scatter_bytes;
        t->sort_ms += c.tm.sort_ms; t->sort_contract_bytes += c.tm.sort_contract_bytes; t->sort_elems += c.tm.sort_elems;
        t->cells_total += c.tm.cells_total; t->wide_total += c.tm.wide_total; t->la_total += c.tm.la_total;
        t->align_kernel_ms += c.tm.align_kernel_ms; t->align_launches += c.tm.align_launches; t->redo_kernel_ms += c.tm.redo_kernel_ms;
    }
}
void damar_reset_kernel_timers(void) {
    for (int i = 0; i < MAX_WORKERS; i++) {
        damar_timings &t = g_ctx[i].tm;
        t.scatter_ms = 0; t.scatter_elems = 0; t.scatter_launches = 0; t.scatter_bytes = 0; t.sort_ms = 0; t.sort_contract_bytes = 0; t.sort_elems = 0; t.cells_total = 0; t.wide_total = 0; t.la_total = 0; t.align_kernel_ms = 0; t.align_launches = 0; t.redo_kernel_ms = 0;
    }
}

}  // extern "C"
