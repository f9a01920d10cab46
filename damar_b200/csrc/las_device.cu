// las_device.cu -- the output path on the device: Write_Overlap_Buffer (reference dalign/align.c:6166-6367) for records
// that never left HBM.
//
// damar_match_filter_sink leaves the records of a Match_Filter call on the device (DeviceResult); damar_dsink_serialise
// then, for the calls of one B block (its N and C comparison),
//   dsink_keys    flattens SORT_OVL (align.c:6104-6164: aread, bread, comp, abpos, then aepos, bbpos, bepos; ties by the
//                 order in which the report threads buffered the records = emission rank) into four key words per record,
//   radix_sort    orders the record numbers by them, least significant word first (the sort is stable),
//   dsink_sizes   + scan give every record its byte offset in the .las image (40-byte header + trace), and how many
//                 records / bytes belong to the first file when a block pair writes two (align.c:6330-6354),
//   dsink_build   one warp per record copies header and trace into the image in HBM,
// and ONE device-to-host copy per file moves the finished image into (pinned) host memory.  Round 1 did the sort and the
// image build on host threads, after shipping records, offsets, ranks and traces separately.
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace damar {

uint8_t *img_alloc_host(size_t n);       // las_sink.cu: pooled host images (pinned while the pool is small)

struct DeviceSink {
    std::vector<DeviceResult *> calls;
    std::mutex mu;
    ~DeviceSink() { for (auto *r : calls) delete r; }
};

namespace {

struct CallView {                        // one call's arrays inside the concatenated record numbering
    const damar_overlap *recs; const uint64_t *toff, *seq; const uint8_t *trace, *dead;
    uint32_t first, n; uint32_t ordinal; int tbytes;
};
struct SinkArgs {
    const CallView *call; int ncalls; uint32_t n;    // device table: a job's two Match_Filter calls, more when one was cut into chunks
    int only_identity, no_trace, pb, rb;             // pb: bits of a read coordinate, rb: bits of a DB-wide read id
    uint64_t *key; uint32_t *idx;
    unsigned int *counts;                            // [0] excluded records, [1] records of the first file
    int split_read;
};

__device__ __forceinline__ int call_of(const SinkArgs &a, uint32_t i) {
    int lo = 0, hi = a.ncalls - 1;                   // last call whose first record is <= i
    while (lo < hi) { const int m = (lo + hi + 1) >> 1; if (a.call[m].first <= i) lo = m; else hi = m - 1; }
    return lo;
}
// Key word `w` of record i (most significant first): 3 = aread | bread, 2 = comp | abpos | aepos, 1 = bbpos | bepos,
// 0 = call ordinal | emission rank.  Excluded records (dropped by Handle_Redundancies, or not aread == bread under -O,
// align.c:6185-6194) sort last: the bit above word 3's fields.
__device__ __forceinline__ uint64_t key_word(const SinkArgs &a, uint32_t i, int w, bool *excluded) {
    const int c = call_of(a, i);
    const CallView &v = a.call[c];
    const uint32_t j = i - v.first;
    const damar_overlap r = v.recs[j];
    const bool ex = (v.dead && v.dead[j]) || (a.only_identity && r.aread != r.bread);
    if (excluded) *excluded = ex;
    switch (w) {
    case 3: return ex ? (1ull << (2 * a.rb)) : (((uint64_t)(uint32_t)r.aread << a.rb) | (uint32_t)r.bread);
    case 2: return ((uint64_t)(r.flags & 1u) << (2 * a.pb)) | ((uint64_t)(uint32_t)r.abpos << a.pb) | (uint32_t)r.aepos;
    case 1: return ((uint64_t)(uint32_t)r.bbpos << a.pb) | (uint32_t)r.bepos;
    default: return ((uint64_t)v.ordinal << 48) | v.seq[j];
    }
}
__global__ void dsink_keys(SinkArgs a, int w, int first_pass) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const uint32_t i = first_pass ? t : a.idx[t];
    bool ex;
    a.key[t] = key_word(a, i, w, &ex);
    if (first_pass) { a.idx[t] = t; if (ex) atomicAdd(&a.counts[0], 1u); }
}
// bytes of the record at sorted position t; counts the records of the first file (aread <= split_read: a prefix)
__global__ void dsink_sizes(SinkArgs a, uint32_t nlive, uint32_t *__restrict__ bytes) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlive) return;
    const uint32_t i = a.idx[t];
    const CallView &v = a.call[call_of(a, i)];
    const damar_overlap r = v.recs[i - v.first];
    bytes[t] = 40u + (a.no_trace ? 0u : (uint32_t)r.tlen * (uint32_t)v.tbytes);
    if (a.split_read >= 0 && r.aread <= a.split_read) atomicAdd(&a.counts[1], 1u);
}
// one warp per record: header (padding zeroed; tlen = 0 under -T, align.c:6086-6090) and trace bytes into the image
__global__ void __launch_bounds__(256) dsink_build(SinkArgs a, uint32_t nlive, const uint64_t *__restrict__ off, uint32_t n1, uint64_t bytes1,
                                                  uint8_t *__restrict__ img1, uint8_t *__restrict__ img2) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= nlive) return;
    const uint32_t i = a.idx[t];
    const CallView &v = a.call[call_of(a, i)];
    const uint32_t j = i - v.first;
    damar_overlap r = v.recs[j];
    uint8_t *dst = (t < n1 || img2 == nullptr) ? img1 + 12 + off[t] : img2 + 12 + (off[t] - bytes1);
    const uint32_t tl = (uint32_t)r.tlen;
    if (a.no_trace) r.tlen = 0;
    r.pad = 0;
    if (lane < 10) {                                     // 40 bytes; the image is only byte aligned
        const uint32_t word = reinterpret_cast<const uint32_t *>(&r)[lane];
#pragma unroll
        for (int b = 0; b < 4; b++) dst[lane * 4 + b] = (uint8_t)(word >> (8 * b));
    }
    if (!a.no_trace) {
        const uint8_t *src = v.trace + v.toff[j] * (uint64_t)v.tbytes;
        const uint32_t nb = tl * (uint32_t)v.tbytes;
        for (uint32_t k = lane; k < nb; k += 32) dst[40 + k] = src[k];
    }
}

int bits_for_u(uint64_t maxval) { int b = 1; while (b < 63 && (1ull << b) <= maxval) b++; return b; }

}  // namespace

int dsink_serialise(DeviceSink *s, int trace_spacing, int no_trace, int only_identity, int split_read, int maxlen, int maxread,
                    uint8_t **buf1, int64_t *len1, uint8_t **buf2, int64_t *len2) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const int tbytes = (trace_spacing <= 125) ? 1 : 2;       // TRACE_XOVR, align.h:58
    SinkArgs a{};
    std::vector<DeviceResult *> calls;
    { std::lock_guard<std::mutex> lock(s->mu); calls = s->calls; }
    std::sort(calls.begin(), calls.end(), [](const DeviceResult *x, const DeviceResult *y) { return x->ordinal < y->ordinal; });
    uint64_t n = 0;
    std::vector<CallView> table;
    for (DeviceResult *r : calls) {
        if (r->nrec == 0) continue;
        if (!no_trace && r->tbytes != tbytes) { set_error("trace width of the buffered overlaps (%d) does not match -s%d", r->tbytes, trace_spacing); return 1; }
        table.emplace_back();
        CallView &v = table.back();
        v.recs = r->recs.p; v.toff = r->toff.p; v.seq = r->seq.p; v.trace = r->trace.p; v.dead = r->have_dead ? r->dead.p : nullptr;
        v.first = (uint32_t)n; v.n = (uint32_t)r->nrec; v.ordinal = (uint32_t)r->ordinal; v.tbytes = r->tbytes;
        n += r->nrec;
    }
    auto header_only = [&](uint8_t **buf, int64_t *len) {
        uint8_t *b = img_alloc_host(12);
        int64_t z = 0; int32_t ts = trace_spacing;
        memcpy(b, &z, 8); memcpy(b + 8, &ts, 4);
        *buf = b; *len = 12;
    };
    if (buf2) { *buf2 = nullptr; *len2 = 0; }
    if (n == 0) { header_only(buf1, len1); if (split_read >= 0 && buf2) header_only(buf2, len2); return 0; }
    if (n >= 0xfffffff0ull) { set_error("damar_dsink_serialise: %llu records in one file", (unsigned long long)n); return 1; }
    DevBuf<CallView> dtable; dtable.reserve(table.size());
    CUDA_CHECK(cudaMemcpyAsync(dtable.p, table.data(), sizeof(CallView) * table.size(), cudaMemcpyHostToDevice, st));
    a.call = dtable.p; a.ncalls = (int)table.size();
    a.n = (uint32_t)n; a.only_identity = only_identity; a.no_trace = no_trace; a.split_read = split_read;
    a.pb = bits_for_u((uint64_t)maxlen + 1); a.rb = bits_for_u((uint64_t)maxread + 1);

    DevBuf<uint64_t> keys, off; DevBuf<uint32_t> idx, bytes; DevBuf<unsigned int> counts; DevBuf<uint64_t> total;
    keys.reserve(2 * (size_t)n + 8); idx.reserve(2 * (size_t)n + 8); counts.reserve(4); total.reserve(1);
    CUDA_CHECK(cudaMemsetAsync(counts.p, 0, sizeof(unsigned int) * 4, st));
    a.counts = counts.p;
    uint64_t *k0 = keys.p, *k1 = keys.p + ((n + 3) & ~3ull);
    uint32_t *v0 = idx.p, *v1 = idx.p + ((n + 3) & ~3ull);
    const unsigned nb = (unsigned)((n + 255) / 256);
    const int wbits[4] = {48 + 4, 2 * a.pb, 2 * a.pb + 1, 2 * a.rb + 1};
    for (int w = 0; w < 4; w++) {                            // least significant word first; the sort is stable
        a.key = k0; a.idx = v0;
        dsink_keys<<<nb, 256, 0, st>>>(a, w, w == 0 ? 1 : 0);
        KERNEL_CHECK(); count_launch();
        if (radix_sort(k0, v0, k1, v1, (int64_t)n, 0, wbits[w], st) != 0) return 1;
    }
    a.idx = v0;
    unsigned int hcnt[2] = {0, 0};
    CUDA_CHECK(cudaMemcpyAsync(hcnt, counts.p, sizeof(hcnt), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    const uint32_t nlive = (uint32_t)n - hcnt[0];
    if (nlive == 0) { header_only(buf1, len1); if (split_read >= 0 && buf2) header_only(buf2, len2); return 0; }
    bytes.reserve((size_t)nlive + 1); off.reserve((size_t)nlive + 1);
    dsink_sizes<<<(nlive + 255) / 256, 256, 0, st>>>(a, nlive, bytes.p);
    KERNEL_CHECK(); count_launch();
    exclusive_scan_u32_to_u64(bytes.p, off.p, nlive, total.p, st);
    uint64_t tot = 0, b1 = 0; unsigned int n1 = nlive;
    CUDA_CHECK(cudaMemcpyAsync(&tot, total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (split_read >= 0) CUDA_CHECK(cudaMemcpyAsync(&n1, counts.p + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    const bool two = split_read >= 0 && buf2 != nullptr;
    if (two) {
        if (n1 < nlive) { CUDA_CHECK(cudaMemcpyAsync(&b1, off.p + n1, sizeof(uint64_t), cudaMemcpyDeviceToHost, st)); CUDA_CHECK(stream_wait(st)); }
        else b1 = tot;
    } else { n1 = nlive; b1 = tot; }
    DevBuf<uint8_t> img;
    img.reserve((size_t)tot + 64);
    uint8_t *d1 = img.p, *d2 = two ? img.p + 12 + b1 + 4 : nullptr;      // second image right behind the first
    {
        int64_t h1[2] = {(int64_t)n1, 0}, h2[2] = {(int64_t)(nlive - n1), 0};
        int32_t ts = trace_spacing;
        uint8_t hdr[12];
        memcpy(hdr, h1, 8); memcpy(hdr + 8, &ts, 4);
        CUDA_CHECK(cudaMemcpyAsync(d1, hdr, 12, cudaMemcpyHostToDevice, st));
        CUDA_CHECK(stream_wait(st));
        if (two) { memcpy(hdr, h2, 8); CUDA_CHECK(cudaMemcpyAsync(d2, hdr, 12, cudaMemcpyHostToDevice, st)); CUDA_CHECK(stream_wait(st)); }
    }
    dsink_build<<<(unsigned)(((uint64_t)nlive * 32 + 255) / 256), 256, 0, st>>>(a, nlive, off.p, n1, b1, d1, d2);
    KERNEL_CHECK(); count_launch();
    uint8_t *o1 = img_alloc_host(12 + b1);
    CUDA_CHECK(cudaMemcpyAsync(o1, d1, 12 + b1, cudaMemcpyDeviceToHost, st));
    uint8_t *o2 = nullptr;
    if (two) { o2 = img_alloc_host(12 + (tot - b1)); CUDA_CHECK(cudaMemcpyAsync(o2, d2, 12 + (tot - b1), cudaMemcpyDeviceToHost, st)); }
    CUDA_CHECK(stream_wait(st));
    *buf1 = o1; *len1 = (int64_t)(12 + b1);
    if (two) { *buf2 = o2; *len2 = (int64_t)(12 + (tot - b1)); }
    return 0;
}

void dsink_add(DeviceSink *s, DeviceResult *r) { std::lock_guard<std::mutex> lock(s->mu); s->calls.push_back(r); }

}  // namespace damar

using namespace damar;

extern "C" {

damar_device_sink *damar_dsink_new(void) { return reinterpret_cast<damar_device_sink *>(new DeviceSink()); }
void damar_dsink_free(damar_device_sink *s) { ensure_init(); delete reinterpret_cast<DeviceSink *>(s); }
void damar_dsink_reset(damar_device_sink *s_) {
    ensure_init();
    DeviceSink *s = reinterpret_cast<DeviceSink *>(s_);
    std::lock_guard<std::mutex> lock(s->mu);
    for (auto *r : s->calls) delete r;
    s->calls.clear();
}

int damar_dsink_serialise(damar_device_sink *s, int trace_spacing, int no_trace, int only_identity, int split_read, int maxlen, int maxread,
                          uint8_t **buf1, int64_t *len1, uint8_t **buf2, int64_t *len2) {
    ensure_init();
    HostProf hp("dsink_serialise");
    return run_with_memory_retry("damar_dsink_serialise", 1, [&] {
        return dsink_serialise(reinterpret_cast<DeviceSink *>(s), trace_spacing, no_trace, only_identity, split_read, maxlen, maxread,
                               buf1, len1, buf2, len2);
    });
}

}  // extern "C"
