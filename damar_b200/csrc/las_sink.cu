// las_sink.cu -- host side of the output path: the per-job overlap buffer and the .las writer.
// Mirrors Overlap_IO_Buffer / AddOverlapToBuffer / Write_Overlap_Buffer / Reset_Overlap_Buffer
// (reference dalign/align.c:5969-6380), Write_Overlap (align.c:3365-3373) and Compress_TraceTo8
// (align.c:3375-3396).  Host-only code; nothing here touches the device.  The final sort and the
// image build are spread over host threads (the reference does both single-threaded).
#include <algorithm>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace damar {

struct LasSink {
    std::vector<damar_overlap> ovl;
    std::vector<uint64_t> toff;      // byte offsets into `trace`
    std::vector<uint64_t> seq;       // (call ordinal << 48) | emission rank inside the call
    std::vector<uint8_t> trace;      // raw trace bytes (uint8 or uint16 values)
    int tbytes = 0;
    uint64_t calls = 0;
    std::mutex mu;                   // damar_sink_add_at may be called from several worker threads
};

struct SortKey {                     // SORT_OVL (align.c:6104-6164) flattened into integers
    uint64_t k0;                     // aread, bread
    uint64_t k1;                     // comp, abpos, aepos
    uint64_t k2;                     // bbpos, bepos
    uint64_t seq;                    // the reference falls back to the record address = emission order
    uint32_t idx;
    bool operator<(const SortKey &o) const {
        if (k0 != o.k0) return k0 < o.k0;
        if (k1 != o.k1) return k1 < o.k1;
        if (k2 != o.k2) return k2 < o.k2;
        return seq < o.seq;
    }
};

// .las images are tens to hundreds of MB; glibc serves such requests with fresh mmaps (page faults on
// first touch dominate the build).  Freed images are kept on a small free list and reused.
struct ImgHdr { size_t cap; size_t magic; };          // magic's low bit: the block is pinned (cudaHostAlloc)
static std::vector<ImgHdr *> g_img_free;
static std::mutex g_img_mu;
static size_t g_img_pinned = 0;                        // bytes of pinned images alive (in use or pooled)
static const size_t IMG_MAGIC = 0x64616d61726c6170ull;
static const size_t IMG_PIN_CAP = (size_t)6 << 30;     // beyond this images are ordinary memory (the copy is then staged by the driver)
static void img_release(ImgHdr *h) {
    if (h->magic & 1) { g_img_pinned -= h->cap; cudaFreeHost(h); } else free(h);
}
static uint8_t *img_alloc(size_t n) {
    std::lock_guard<std::mutex> lock(g_img_mu);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < g_img_free.size(); i++)
        if (g_img_free[i]->cap >= n && (best == (size_t)-1 || g_img_free[i]->cap < g_img_free[best]->cap)) best = i;
    ImgHdr *h = nullptr;
    if (best != (size_t)-1 && g_img_free[best]->cap <= 2 * n + (1u << 20)) { h = g_img_free[best]; g_img_free.erase(g_img_free.begin() + best); }
    else {
        size_t cap = n + n / 8 + 4096;
        bool pinned = false;
        if (n >= (1u << 20) && g_img_pinned + cap <= IMG_PIN_CAP) {      // device-to-host copies run at full PCIe rate into pinned memory
            void *q = nullptr;
            if (cudaHostAlloc(&q, sizeof(ImgHdr) + cap, cudaHostAllocDefault) == cudaSuccess) { h = (ImgHdr *)q; pinned = true; g_img_pinned += cap; }
            else cudaGetLastError();
        }
        if (!h) h = (ImgHdr *)malloc(sizeof(ImgHdr) + cap);
        if (!h) fatal("out of host memory serialising %zu bytes of .las", n);
        h->cap = cap; h->magic = IMG_MAGIC | (pinned ? 1 : 0);
    }
    return reinterpret_cast<uint8_t *>(h + 1);
}
uint8_t *img_alloc_host(size_t n) { return img_alloc(n ? n : 1); }
static void img_free(void *p) {
    if (!p) return;
    ImgHdr *h = reinterpret_cast<ImgHdr *>(p) - 1;
    if ((h->magic & ~(size_t)1) != IMG_MAGIC) { free(p); return; }
    std::lock_guard<std::mutex> lock(g_img_mu);
    if (g_img_free.size() >= 24) {                        // keep the 24 largest
        size_t small = 0;
        for (size_t i = 1; i < g_img_free.size(); i++) if (g_img_free[i]->cap < g_img_free[small]->cap) small = i;
        if (g_img_free[small]->cap < h->cap) { img_release(g_img_free[small]); g_img_free[small] = h; } else img_release(h);
        return;
    }
    g_img_free.push_back(h);
}

static int host_threads() {
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 4;
    if (n > 32) n = 32;
    return (int)n;
}

// sorted order of the selected records
static void sorted_order(const LasSink &s, int only_identity, std::vector<SortKey> &keys) {
    keys.clear();
    keys.reserve(s.ovl.size());
    for (uint32_t i = 0; i < s.ovl.size(); i++) {
        const damar_overlap &r = s.ovl[i];
        if (only_identity && r.aread != r.bread) continue;                       // align.c:6185-6194
        SortKey k;
        k.k0 = ((uint64_t)(uint32_t)r.aread << 32) | (uint32_t)r.bread;          // ids and coordinates are >= 0
        k.k1 = ((uint64_t)(r.flags & 1u) << 62) | ((uint64_t)(uint32_t)r.abpos << 31) | (uint32_t)r.aepos;
        k.k2 = ((uint64_t)(uint32_t)r.bbpos << 32) | (uint32_t)r.bepos;
        k.seq = s.seq[i]; k.idx = i;
        keys.push_back(k);
    }
    const size_t n = keys.size();
    const int T = host_threads();
    if (n < 50000 || T < 2) { std::sort(keys.begin(), keys.end()); return; }
    // bucket by aread range, sort buckets in parallel, concatenate
    uint32_t amin = ~0u, amax = 0;
    for (const SortKey &k : keys) { uint32_t a = (uint32_t)(k.k0 >> 32); if (a < amin) amin = a; if (a > amax) amax = a; }
    const uint64_t span = (uint64_t)amax - amin + 1;
    std::vector<size_t> cnt(T + 1, 0);
    auto bucket = [&](const SortKey &k) { return (size_t)((((k.k0 >> 32) - amin) * (uint64_t)T) / span); };
    for (const SortKey &k : keys) cnt[bucket(k) + 1]++;
    for (int t = 0; t < T; t++) cnt[t + 1] += cnt[t];
    std::vector<SortKey> tmp(n);
    { std::vector<size_t> pos(cnt.begin(), cnt.end() - 1); for (const SortKey &k : keys) tmp[pos[bucket(k)]++] = k; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([&, t] { std::sort(tmp.begin() + cnt[t], tmp.begin() + cnt[t + 1]); });
    for (auto &x : th) x.join();
    keys.swap(tmp);
}

static int put_range(const LasSink &s, const std::vector<SortKey> &keys, size_t from, size_t to, int tspace, int no_trace,
                     uint8_t **out, int64_t *len) {
    const int tbytes = (tspace <= 125) ? 1 : 2;               // TRACE_XOVR, align.h:58
    if (!no_trace && to > from && s.tbytes != tbytes) { set_error("trace width of the buffered overlaps (%d) does not match -s%d", s.tbytes, tspace); return 1; }
    const size_t n = to - from;
    std::vector<size_t> offs(n + 1);
    offs[0] = 12;
    for (size_t i = 0; i < n; i++) offs[i + 1] = offs[i] + 40 + (no_trace ? 0 : (size_t)tbytes * s.ovl[keys[from + i].idx].tlen);
    uint8_t *buf = img_alloc(offs[n] ? offs[n] : 1);
    int64_t n64 = (int64_t)n; int32_t ts = tspace;
    memcpy(buf, &n64, 8); memcpy(buf + 8, &ts, 4);
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            const uint32_t id = keys[from + i].idx;
            damar_overlap r = s.ovl[id];
            const int tl = r.tlen;
            if (no_trace) r.tlen = 0;                         // align.c:6086-6090
            r.pad = 0;
            memcpy(buf + offs[i], &r, 40);
            if (!no_trace && tl > 0) memcpy(buf + offs[i] + 40, s.trace.data() + s.toff[id], (size_t)tbytes * tl);
        }
    };
    const int T = host_threads();
    if (n < 20000 || T < 2) work(0, n);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, n * t / T, n * (t + 1) / T);
        for (auto &x : th) x.join();
    }
    *out = buf; *len = (int64_t)offs[n];
    return 0;
}

}  // namespace damar

using namespace damar;

extern "C" {

damar_las_sink *damar_sink_new(void) { return reinterpret_cast<damar_las_sink *>(new LasSink()); }
void damar_sink_free(damar_las_sink *s) { delete reinterpret_cast<LasSink *>(s); }
void damar_sink_reset(damar_las_sink *s_) {
    LasSink *s = reinterpret_cast<LasSink *>(s_);
    s->ovl.clear(); s->toff.clear(); s->seq.clear(); s->trace.clear(); s->calls = 0; s->tbytes = 0;
}
int64_t damar_sink_count(const damar_las_sink *s) { return (int64_t) reinterpret_cast<const LasSink *>(s)->ovl.size(); }

// `ordinal` = position of this Match_Filter call in the job (0 = first B block normal, 1 = its complement, ...):
// the reference buffers the calls in that order, which breaks full SORT_OVL ties.
void damar_sink_add_at(damar_las_sink *s_, const damar_overlaps *o, int64_t ordinal) {
    HostProf hp("sink_add");
    LasSink *s = reinterpret_cast<LasSink *>(s_);
    std::lock_guard<std::mutex> lock(s->mu);
    if ((uint64_t)ordinal >= s->calls) s->calls = (uint64_t)ordinal + 1;
    if (o->novl == 0) return;
    if (s->tbytes == 0) s->tbytes = o->trace_bytes;
    if (s->tbytes != o->trace_bytes) fatal("damar_sink_add: mixed trace widths in one sink");
    const uint64_t base = s->trace.size();
    const size_t n0 = s->ovl.size();
    s->ovl.insert(s->ovl.end(), o->ovl, o->ovl + o->novl);
    s->toff.resize(n0 + o->novl); s->seq.resize(n0 + o->novl);
    for (int64_t i = 0; i < o->novl; i++) {
        s->toff[n0 + i] = base + o->trace_off[i] * (uint64_t)o->trace_bytes;
        s->seq[n0 + i] = ((uint64_t)ordinal << 48) | (o->seq ? o->seq[i] : (uint64_t)i);
    }
    const uint8_t *tb = reinterpret_cast<const uint8_t *>(o->trace);
    s->trace.insert(s->trace.end(), tb, tb + (size_t)o->ntrace * o->trace_bytes);
}
void damar_sink_add(damar_las_sink *s_, const damar_overlaps *o) {
    damar_sink_add_at(s_, o, (int64_t)reinterpret_cast<LasSink *>(s_)->calls);
}

int damar_sink_serialise(damar_las_sink *s_, int trace_spacing, int no_trace, int only_identity, int split_read,
                         uint8_t **buf1, int64_t *len1, uint8_t **buf2, int64_t *len2) {
    HostProf hp("sink_serialise");
    LasSink *s = reinterpret_cast<LasSink *>(s_);
    std::vector<SortKey> keys;
    sorted_order(*s, only_identity, keys);
    if (split_read < 0) {
        if (buf2) { *buf2 = nullptr; *len2 = 0; }
        return put_range(*s, keys, 0, keys.size(), trace_spacing, no_trace, buf1, len1);
    }
    // first index with aread > split_read (align.c:6330-6336); keys are sorted by aread first
    size_t lo = 0, hi = keys.size();
    while (lo < hi) { size_t m = (lo + hi) >> 1; if ((int32_t)(keys[m].k0 >> 32) <= split_read) lo = m + 1; else hi = m; }
    if (put_range(*s, keys, 0, lo, trace_spacing, no_trace, buf1, len1)) return 1;
    return put_range(*s, keys, lo, keys.size(), trace_spacing, no_trace, buf2, len2);
}

// LAmerge (utils/LAmerge.c, utils/LAmergeUtils.c:31-145): the k-way merge of sorted .las files into one -- what the pipeline
// runs right after daligner to turn a block's pair files R.i.R.1.las ... R.i.R.n.las into R.i.las.  The files of one block
// hold disjoint (aread, bread) pairs, so the merge is decided by the leading fields of SORT_OVL (aread, bread, comp, abpos);
// records are copied verbatim (header: summed novl, common trace spacing).  Host only, like Write_Overlap_Buffer.
int damar_las_merge(int n, const uint8_t *const *bufs, const int64_t *lens, uint8_t **out, int64_t *outlen) {
    struct Cur { const uint8_t *p, *end; int64_t left; };
    std::vector<Cur> cur((size_t)n);
    int64_t novl = 0, total = 12;
    int32_t tspace = 0;
    bool have = false;
    for (int i = 0; i < n; i++) {
        if (lens[i] < 12) return 1;
        int64_t k; int32_t ts;
        memcpy(&k, bufs[i], 8); memcpy(&ts, bufs[i] + 8, 4);
        if (k < 0) return 1;
        if (!have) { tspace = ts; have = true; }
        else if (ts != tspace) return 2;                    // "PT-point spacing conflict" (LAmerge.c:380-385)
        cur[(size_t)i] = Cur{bufs[i] + 12, bufs[i] + lens[i], k};
        novl += k; total += lens[i] - 12;
    }
    const int tbytes = (tspace <= 125) ? 1 : 2;
    uint8_t *img = (uint8_t *)img_alloc((size_t)total);
    if (img == nullptr) return 3;
    memcpy(img, &novl, 8); memcpy(img + 8, &tspace, 4);
    uint8_t *w = img + 12;
    auto key = [](const uint8_t *r, uint64_t &k0, uint64_t &k1) {
        int32_t f[9];
        memcpy(f, r, 36);                                   // tlen, diffs, abpos, bbpos, aepos, bepos, flags, aread, bread
        k0 = ((uint64_t)(uint32_t)f[7] << 32) | (uint32_t)f[8];
        k1 = ((uint64_t)((uint32_t)f[6] & 1u) << 32) | (uint32_t)f[2];
    };
    struct Item { uint64_t k0, k1; int src; };
    auto later = [](const Item &a, const Item &b) { return a.k0 != b.k0 ? a.k0 > b.k0 : a.k1 != b.k1 ? a.k1 > b.k1 : a.src > b.src; };
    std::vector<Item> heap;
    for (int i = 0; i < n; i++)
        if (cur[(size_t)i].left > 0) {
            if (cur[(size_t)i].p + 40 > cur[(size_t)i].end) { img_free(img); return 1; }
            Item it; it.src = i; key(cur[(size_t)i].p, it.k0, it.k1); heap.push_back(it);
        }
    std::make_heap(heap.begin(), heap.end(), later);
    while (!heap.empty()) {
        std::pop_heap(heap.begin(), heap.end(), later);
        const int i = heap.back().src;
        heap.pop_back();
        Cur &c = cur[(size_t)i];
        int32_t tlen; memcpy(&tlen, c.p, 4);
        const int64_t sz = 40 + (int64_t)tlen * tbytes;
        if (tlen < 0 || c.p + sz > c.end) { img_free(img); return 1; }
        memcpy(w, c.p, (size_t)sz); w += sz; c.p += sz; c.left--;
        if (c.left > 0) {
            if (c.p + 40 > c.end) { img_free(img); return 1; }
            Item it; it.src = i; key(c.p, it.k0, it.k1); heap.push_back(it);
            std::push_heap(heap.begin(), heap.end(), later);
        }
    }
    *out = img; *outlen = (int64_t)(w - img);
    return 0;
}

/* Byte comparison of two .las images that ignores bytes 36..39 of every record (struct padding the reference leaves
 * uninitialised, align.c:3345-3346).  Returns -1 when equal, -2 when the lengths differ or an image is malformed,
 * else the offset of the first differing byte.  Host only (test / bench helper). */
int64_t damar_las_compare(const uint8_t *x, int64_t xlen, const uint8_t *y, int64_t ylen) {
    if (xlen != ylen || xlen < 12) return -2;
    for (int i = 0; i < 12; i++) if (x[i] != y[i]) return i;
    int64_t novl; int32_t tspace;
    memcpy(&novl, x, 8); memcpy(&tspace, x + 8, 4);
    const int tbytes = (tspace <= 125) ? 1 : 2;
    int64_t o = 12;
    for (int64_t r = 0; r < novl; r++) {
        if (o + 40 > xlen) return -2;
        int32_t tl; memcpy(&tl, x + o, 4);
        if (memcmp(x + o, y + o, 36) != 0) { for (int i = 0; i < 36; i++) if (x[o + i] != y[o + i]) return o + i; }
        const int64_t nb = (int64_t)tl * tbytes;
        if (tl < 0 || o + 40 + nb > xlen) return -2;
        if (memcmp(x + o + 40, y + o + 40, (size_t)nb) != 0) { for (int64_t i = 0; i < nb; i++) if (x[o + 40 + i] != y[o + 40 + i]) return o + 40 + i; }
        o += 40 + nb;
    }
    return o == xlen ? -1 : -2;
}

void damar_free(void *p) { img_free(p); }

}  // extern "C"
