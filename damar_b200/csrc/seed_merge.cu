// seed_merge.cu -- stage 2: A-versus-B merge of two sorted k-mer lists into seed hits, then the
// (bread, aread, apos) sort.
//
// Replaces count_thread (reference dalign/filter.c:1039-1165), the -M limit choice
// (filter.c:2634-2702), merge_thread (filter.c:1170-1358) and the seed-hit lex_sort
// (filter.c:2561-2580, 2776-2781).
//
// Shape of the computation on the device:
//   1. run-length encode the A list by code  -> group id per entry, start per group
//   2. one thread per A group: locate the equal-code run in B (two binary searches; the
//      searches of neighbouring threads walk neighbouring cache lines because both lists are
//      sorted), compute the group's match count ct, optionally histogram it (-M rule)
//   3. host: pick `limit` from the 10000-bin histogram exactly as filter.c:2645-2663 does
//   4. one thread per A entry: number of B partners it emits (0 when ct >= limit); scan
//   5. one thread per A entry: write its hits; hit = (key, bpos) with
//         key = bread << (abits+pbits) | aread << pbits | apos
//   6. stable radix sort on the key: ties keep merge order (= bpos ascending), which is the
//      order report_thread depends on.
#include <type_traits>

#include <mutex>

#include "common.cuh"

namespace damar {
namespace {

constexpr int MAXGRAM = 10000;   // filter.c:71

__global__ void mark_heads(const uint64_t *__restrict__ codes, int64_t n, uint32_t *__restrict__ head) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || codes[i] != codes[i - 1]) ? 1u : 0u;
}
// gid[] holds the exclusive scan of the head flags on entry, the group id on exit
__global__ void finish_groups(const uint64_t *__restrict__ codes, int64_t n, uint32_t *__restrict__ gid, uint32_t *__restrict__ gstart) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool h = (i == 0 || codes[i] != codes[i - 1]);
    uint32_t g = gid[i] + (h ? 1u : 0u) - 1u;
    gid[i] = g;
    if (h) gstart[g] = (uint32_t)i;
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *__restrict__ a, int64_t n, uint64_t x) {
    int64_t l = 0, r = n;
    while (l < r) { int64_t m = (l + r) >> 1; if (a[m] < x) l = m + 1; else r = m; }
    return l;
}
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *__restrict__ a, uint32_t l, uint32_t r, uint32_t x) {
    while (l < r) { uint32_t m = (l + r) >> 1; if (a[m] < x) l = m + 1; else r = m; }
    return l;
}
// the read whose slot range [boff[r], boff[r+1]) holds gpos p, from the block's lookup table (kmer_index.cu::build_rtab)
struct ReadMap { const uint32_t *boff, *rtab; };
__device__ __forceinline__ int read_of(const ReadMap m, uint32_t p) {
    int r = (int)m.rtab[p >> RTAB_SHIFT];
    while (m.boff[r + 1] <= p) r++;
    return r;
}

// Number of B-group entries admitted for A entry (gpos ga) in a self comparison:
//   default      : b.read <  a.read                          (filter.c:1108-1112)
//   -I, N        : (b.read,b.rpos) < (a.read,a.rpos)         (filter.c:1097-1104)
//   -I, C        : b.read <= a.read                          (filter.c:1092-1096)
// All three are "gposB < threshold" because the group is in gpos order and both blocks share boff.
__device__ __forceinline__ uint32_t self_partners(const uint32_t *__restrict__ bgpos, uint32_t jb, uint32_t ib, uint32_t ga,
                                                  const ReadMap m, int identity, int comp) {
    uint32_t thr;
    if (identity && !comp) thr = ga;
    else {
        int r = read_of(m, ga);
        thr = (identity && comp) ? m.boff[r + 1] : m.boff[r];
    }
    return lower_bound_u32(bgpos, jb, ib, thr) - jb;
}

struct GroupArgs {
    const uint64_t *acodes; const uint32_t *agpos; int64_t alen;
    const uint64_t *bcodes; const uint32_t *bgpos; int64_t blen;
    const uint32_t *gstart; uint32_t ngroups;
    ReadMap amap;                                // of the A block (== B block when self)
    int self, comp, identity, same_list;
    uint32_t *gjb, *gnb, *gct;                   // per group: B run start, B run length, min(ct, 2^32-1)
    unsigned long long *gram;                    // MAXGRAM bins or nullptr
};

// [jb, ib) = the run of `code` inside the sorted window w[0, n)
__device__ __forceinline__ void equal_run(const uint64_t *w, int64_t n, uint64_t code, int64_t &jb, int64_t &ib) {
    jb = lower_bound_u64(w, n, code);
    ib = jb;
    if (jb < n && w[jb] == code) {                               // equal-code runs are short: gallop, then bisect
        int64_t step = 1, e = jb + 1;
        while (e < n && w[e] == code) { e += step; step <<= 1; }
        const int64_t from = e - (step >> 1), to = (e < n) ? e : n;
        ib = (code == ~0ull) ? n : from + lower_bound_u64(w + from, to - from, code + 1);
    }
}

// Each CTA owns a contiguous chunk of A groups and walks it in tiles of MG_TILE groups.  Because both
// lists are sorted, the B entries that can match a tile lie in one window [lo, hi): two threads
// locate it with a full binary search; the window's codes are then staged in shared memory (coalesced
// stream of the B list, read once) and every thread bisects there.  Windows larger than MG_WIN entries
// (a tile that spans a very frequent code) are searched in place.
// Self comparison of a list with itself (N orientation): the B run IS the A run and the partners of an entry
// are the entries of the run that belong to earlier reads -- a run of one k-mer (most of them: any k-mer with
// a sequencing error in it) has none and needs no search at all.
constexpr int MG_TILE = 2048;
constexpr int MG_WIN = 4096;
// window of B entries that can match tile t of A groups: one thread per tile, so the two 27-step bisections of the whole
// B list (DRAM latency each step) of all tiles run side by side instead of one tile after the other inside a CTA
__global__ void tile_windows(GroupArgs a, int64_t *__restrict__ wlo, int64_t *__restrict__ whi, uint32_t ntiles) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint64_t t0 = (uint64_t)t * MG_TILE, t1 = (t0 + MG_TILE < a.ngroups) ? t0 + MG_TILE : a.ngroups;
    const uint64_t first = a.acodes[a.gstart[t0]], last = a.acodes[a.gstart[t1 - 1]];
    wlo[t] = lower_bound_u64(a.bcodes, a.blen, first);
    whi[t] = (last == ~0ull) ? a.blen : lower_bound_u64(a.bcodes, a.blen, last + 1);
}

__global__ void __launch_bounds__(256) match_groups(GroupArgs a, const int64_t *__restrict__ wlo, const int64_t *__restrict__ whi, uint32_t ntiles) {
    extern __shared__ uint32_t sgram[];
    __shared__ uint64_t swin[MG_WIN];
    const bool histo = (a.gram != nullptr);
    if (histo) { for (int i = threadIdx.x; i < MAXGRAM; i += blockDim.x) sgram[i] = 0; }
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t t0 = (uint64_t)tile * MG_TILE, t1 = (t0 + MG_TILE < a.ngroups) ? t0 + MG_TILE : a.ngroups;
        __syncthreads();
        const int64_t lo = a.same_list ? 0 : wlo[tile], hi = a.same_list ? 0 : whi[tile];
        const bool staged = !a.same_list && (hi - lo <= MG_WIN);
        if (staged) {
            for (int i = threadIdx.x; i < (int)(hi - lo); i += blockDim.x) swin[i] = a.bcodes[lo + i];
            __syncthreads();
        }
        for (uint64_t gg = t0 + threadIdx.x; gg < t1; gg += blockDim.x) {
            const uint32_t g = (uint32_t)gg;
            const uint32_t ja = a.gstart[g];
            const uint32_t ia = (g + 1 < a.ngroups) ? a.gstart[g + 1] : (uint32_t)a.alen;
            uint32_t jb, ib;
            unsigned long long ct = 0;
            if (a.same_list) {
                jb = ja; ib = ia;
                const unsigned long long n = ia - ja;
                if (!a.self) ct = n * n;
                else if (a.identity) ct = n * (n - 1) / 2;          // partners of the x-th entry: the x before it (filter.c:1097-1104)
                else if (n > 1) {                                    // entries of earlier reads (filter.c:1108-1112)
                    uint32_t rs = ja, bhi = 0;
                    for (uint32_t x = ja; x < ia; x++) {
                        const uint32_t gp = a.agpos[x];
                        if (gp >= bhi) { rs = x; bhi = a.amap.boff[read_of(a.amap, gp) + 1]; }
                        ct += rs - ja;
                    }
                }
            } else {
                const uint64_t code = a.acodes[ja];
                int64_t rj, ri;
                if (staged) equal_run(swin, hi - lo, code, rj, ri);
                else equal_run(a.bcodes + lo, hi - lo, code, rj, ri);
                jb = (uint32_t)(lo + rj); ib = (uint32_t)(lo + ri);
                if (ib > jb) {
                    if (a.self) {
                        for (uint32_t x = ja; x < ia; x++)
                            ct += self_partners(a.bgpos, jb, ib, a.agpos[x], a.amap, a.identity, a.comp);
                    } else
                        ct = (unsigned long long)(ia - ja) * (unsigned long long)(ib - jb);
                }
            }
            if (histo && ib > jb && ct < MAXGRAM) atomicAdd(&sgram[(uint32_t)ct], 1u);
            a.gjb[g] = jb; a.gnb[g] = ib - jb;
            a.gct[g] = (ct > 0xffffffffull) ? 0xffffffffu : (uint32_t)ct;
        }
    }
    if (histo) {
        __syncthreads();
        for (int i = threadIdx.x; i < MAXGRAM; i += blockDim.x)
            if (sgram[i]) atomicAdd(&a.gram[i], (unsigned long long)sgram[i]);
    }
}

struct EntryArgs {
    const uint32_t *agpos; int64_t alen; const uint32_t *bgpos;
    const uint32_t *gid, *gjb, *gnb, *gct;
    ReadMap amap, bmap;
    int self, comp, identity, same_list; uint32_t limit;
    uint32_t *cnt; const uint64_t *off;
    uint64_t *keys; uint32_t *bpos; int pbits, abits;
};

__global__ void __launch_bounds__(256) count_entries(EntryArgs e) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.alen) return;
    const uint32_t g = e.gid[i];
    uint32_t c = 0;
    const uint32_t nb = e.gnb[g];
    if (nb > 0 && e.gct[g] < e.limit && !(e.same_list && e.self && nb == 1)) {   // a run of one k-mer of the list itself has no partner
        if (e.self) c = self_partners(e.bgpos, e.gjb[g], e.gjb[g] + nb, e.agpos[i], e.amap, e.identity, e.comp);
        else c = nb;
    }
    e.cnt[i] = c;
}

__global__ void __launch_bounds__(256) emit_hits(EntryArgs e) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.alen) return;
    const uint32_t c = e.cnt[i];
    if (c == 0) return;
    const uint32_t g = e.gid[i];
    const uint32_t jb = e.gjb[g];
    const uint32_t ga = e.agpos[i];
    const int ar = read_of(e.amap, ga);
    const uint32_t apos = ga - e.amap.boff[ar];
    const uint64_t klow = ((uint64_t)ar << e.pbits) | apos;
    uint64_t o = e.off[i];
    int br = -1; uint32_t blo = 0, bhi = 0;
    for (uint32_t x = 0; x < c; x++, o++) {
        const uint32_t gb = e.bgpos[jb + x];
        if (br < 0 || gb < blo || gb >= bhi) { br = read_of(e.bmap, gb); blo = e.bmap.boff[br]; bhi = e.bmap.boff[br + 1]; }
        e.keys[o] = ((uint64_t)br << (e.pbits + e.abits)) | klow;
        e.bpos[o] = gb - blo;
    }
}

int bits_for(int64_t maxval_exclusive) {   // smallest b with 2^b >= maxval_exclusive (>= 1)
    int b = 1;
    while ((1ll << b) < maxval_exclusive) b++;
    return b;
}

}  // namespace

void seed_hits(const DeviceIndex *ai, const DeviceIndex *bi, int self, int comp, SeedHits &out) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const DeviceBlock *ab = ai->blk, *bb = bi->blk;
    const int64_t alen = ai->len, blen = bi->len;
    const int same_list = (ai == bi) ? 1 : 0;
    out.release();
    out.limit = INT32_MAX;

    HostProf *hp1 = new HostProf("sh.count");
    Timer t_count(st);
    // 1. groups of A
    DevBuf<uint32_t> gid, gstart, gjb, gnb, gct, cnt;
    DevBuf<uint64_t> off, total;
    DevBuf<unsigned long long> gram;
    gid.reserve((size_t)alen); total.reserve(1);
    DevBuf<uint32_t> &tot32 = scratch().tot32_merge; tot32.reserve(1);
    const unsigned eb = (unsigned)((alen + 255) / 256);
    mark_heads<<<eb, 256, 0, st>>>(ai->codes, alen, gid.p);
    KERNEL_CHECK();
    exclusive_scan_u32_inplace(gid.p, alen, tot32.p, st);
    uint32_t ngroups = 0;
    CUDA_CHECK(cudaMemcpyAsync(&ngroups, tot32.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    gstart.reserve((size_t)ngroups + 1); gjb.reserve(ngroups); gnb.reserve(ngroups); gct.reserve(ngroups);
    finish_groups<<<eb, 256, 0, st>>>(ai->codes, alen, gid.p, gstart.p);
    KERNEL_CHECK();
    count_launch(2);

    // 2. per-group match
    const bool use_limit = (c.opt.mem_limit > 0);
    if (use_limit) { gram.reserve(MAXGRAM); CUDA_CHECK(cudaMemsetAsync(gram.p, 0, sizeof(unsigned long long) * MAXGRAM, st)); }
    GroupArgs ga{};
    ga.acodes = ai->codes; ga.agpos = ai->gpos; ga.alen = alen;
    ga.bcodes = bi->codes; ga.bgpos = bi->gpos; ga.blen = blen;
    ga.gstart = gstart.p; ga.ngroups = ngroups; ga.amap = ReadMap{ab->boff, ab->rtab};
    ga.self = self; ga.comp = comp; ga.identity = c.opt.identity; ga.same_list = same_list;
    ga.gjb = gjb.p; ga.gnb = gnb.p; ga.gct = gct.p; ga.gram = use_limit ? gram.p : nullptr;
    {
        static std::once_flag attr;
        std::call_once(attr, [] { CUDA_CHECK(cudaFuncSetAttribute(match_groups, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXGRAM * 4)); });
        const unsigned ntiles = (ngroups + MG_TILE - 1) / MG_TILE;
        unsigned blocks = ntiles;
        if (blocks > NUM_SMS_B200 * 8) blocks = NUM_SMS_B200 * 8;
        if (blocks == 0) blocks = 1;
        DevBuf<int64_t> wlo, whi;
        if (!same_list && ntiles > 0) {
            wlo.reserve(ntiles); whi.reserve(ntiles);
            tile_windows<<<(ntiles + 127) / 128, 128, 0, st>>>(ga, wlo.p, whi.p, ntiles);
            KERNEL_CHECK(); count_launch();
        }
        match_groups<<<blocks, 256, use_limit ? MAXGRAM * 4 : 0, st>>>(ga, wlo.p, whi.p, ntiles);
        KERNEL_CHECK(); count_launch();
    }

    // 3. limit (filter.c:2634-2699), evaluated on the host from the histogram like the reference's main thread
    int64_t limit = INT32_MAX;
    if (use_limit) {
        std::vector<unsigned long long> histo(MAXGRAM);
        CUDA_CHECK(cudaMemcpyAsync(histo.data(), gram.p, sizeof(unsigned long long) * MAXGRAM, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_wait(st));
        int64_t avail = (int64_t)(((int64_t)(c.opt.mem_limit - (uint64_t)(ab->sizeof_db + bb->sizeof_db))) / (uint64_t)16);
        if (same_list || avail > alen + 2 * blen) avail = (avail - alen) / 2;
        else avail = avail - (alen + blen);
        avail *= .98;
        int64_t tom = 0; int j;
        for (j = 0; j < MAXGRAM; j++) { tom += j * (int64_t)histo[j]; if (tom > avail) break; }
        limit = j;
        if (limit <= 1) fatal("Insufficient memory allocation (%.1fGb), reduce block size or increase allocation", (1. * c.opt.mem_limit) / 0x40000000ll);
        if (limit < 10) fprintf(stderr, "\nWarning: Sensitivity hampered by low memory allocation (%.1fGb), reduce block size or increase allocation\n", (1. * c.opt.mem_limit) / 0x40000000ll);
        if (c.opt.verbose) printf("   Capping mutual k-mer matches over %d (effectively -t%d)\n", (int)limit, (int)__builtin_sqrt(1. * limit));
    }
    out.limit = (int32_t)limit;
    c.tm.count_ms = t_count.stop();

    delete hp1;
    // 4. per-entry partner counts + scan
    HostProf *hp2 = new HostProf("sh.merge");
    Timer t_merge(st);
    cnt.reserve((size_t)alen); off.reserve((size_t)alen);
    EntryArgs ea{};
    ea.agpos = ai->gpos; ea.alen = alen; ea.bgpos = bi->gpos;
    ea.gid = gid.p; ea.gjb = gjb.p; ea.gnb = gnb.p; ea.gct = gct.p;
    ea.amap = ReadMap{ab->boff, ab->rtab}; ea.bmap = ReadMap{bb->boff, bb->rtab};
    ea.self = self; ea.comp = comp; ea.identity = c.opt.identity; ea.same_list = same_list; ea.limit = (uint32_t)limit;
    ea.cnt = cnt.p; ea.off = off.p;
    count_entries<<<eb, 256, 0, st>>>(ea);
    KERNEL_CHECK(); count_launch();
    exclusive_scan_u32_to_u64(cnt.p, off.p, alen, total.p, st);
    uint64_t nhits = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nhits, total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    if (c.opt.verbose) { printf("   Hit count = %lld\n", (long long)nhits); fflush(stdout); }
    if (nhits == 0) { c.tm.merge_ms = t_merge.stop(); delete hp2; return; }

    out.pbits = bits_for(ab->maxlen);
    out.abits = bits_for(ab->nreads);
    out.bbits = bits_for(bb->nreads);
    if (out.pbits + out.abits + out.bbits > 64)
        fatal("seed-hit key needs %d bits (> 64): block too large for this build", out.pbits + out.abits + out.bbits);

    uint64_t *k0, *k1; uint32_t *v0, *v1;
    k0 = (std::remove_reference<decltype(k0)>::type)dev_alloc_bytes(sizeof(uint64_t) * (nhits + 1));
    k1 = (std::remove_reference<decltype(k1)>::type)dev_alloc_bytes(sizeof(uint64_t) * (nhits + 1));
    v0 = (std::remove_reference<decltype(v0)>::type)dev_alloc_bytes(sizeof(uint32_t) * (nhits + 1));
    v1 = (std::remove_reference<decltype(v1)>::type)dev_alloc_bytes(sizeof(uint32_t) * (nhits + 1));
    ea.keys = k0; ea.bpos = v0; ea.pbits = out.pbits; ea.abits = out.abits;
    emit_hits<<<eb, 256, 0, st>>>(ea);
    KERNEL_CHECK(); count_launch();
    c.tm.merge_ms = t_merge.stop();

    delete hp2;
    HostProf hp3("sh.sort+free");
    Timer t_sort(st);
    radix_sort_pairs(k0, v0, k1, v1, (int64_t)nhits, out.pbits + out.abits + out.bbits, st);
    c.tm.hsort_ms = t_sort.stop();
    dev_free(k1); dev_free(v1);
    out.keys = k0; out.bpos = v0; out.n = (int64_t)nhits;
}

}  // namespace damar
