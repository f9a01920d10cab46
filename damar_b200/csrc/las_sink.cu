// las_sink.cu -- host side of the output path: the per-job overlap buffer and the .las writer.
// Mirrors Overlap_IO_Buffer / AddOverlapToBuffer / Write_Overlap_Buffer / Reset_Overlap_Buffer
// (reference dalign/align.c:5969-6380), Write_Overlap (align.c:3365-3373) and Compress_TraceTo8
// (align.c:3375-3396).  Host-only code; nothing here touches the device.
#include <algorithm>
#include <sys/stat.h>

#include "common.cuh"

namespace damar {

struct LasSink {
    std::vector<damar_overlap> ovl;
    std::vector<uint64_t> toff;
    std::vector<uint16_t> trace;
};

// SORT_OVL (align.c:6104-6164).  The reference's last resort on a full key tie is the record
// address inside its gathered array, i.e. emission order; a stable sort reproduces that.
static bool ovl_less(const damar_overlap &l, const damar_overlap &r) {
    if (l.aread != r.aread) return l.aread < r.aread;
    if (l.bread != r.bread) return l.bread < r.bread;
    const int cl = l.flags & 1, cr = r.flags & 1;
    if (cl != cr) return cl < cr;
    if (l.abpos != r.abpos) return l.abpos < r.abpos;
    if (l.aepos != r.aepos) return l.aepos < r.aepos;
    if (l.bbpos != r.bbpos) return l.bbpos < r.bbpos;
    return l.bepos < r.bepos;
}

static int put_range(const LasSink &s, const std::vector<uint32_t> &idx, size_t from, size_t to, int tspace, int no_trace,
                     uint8_t **out, int64_t *len) {
    const int tbytes = (tspace <= 125) ? 1 : 2;               // TRACE_XOVR, align.h:58
    size_t sz = 12;
    for (size_t i = from; i < to; i++) sz += 40 + (no_trace ? 0 : (size_t)tbytes * s.ovl[idx[i]].tlen);
    uint8_t *buf = (uint8_t *)malloc(sz ? sz : 1);
    if (!buf) fatal("out of host memory serialising %zu bytes of .las", sz);
    int64_t n = (int64_t)(to - from); int32_t ts = tspace;
    memcpy(buf, &n, 8); memcpy(buf + 8, &ts, 4);
    size_t o = 12;
    for (size_t i = from; i < to; i++) {
        damar_overlap r = s.ovl[idx[i]];
        const uint16_t *t = s.trace.data() + s.toff[idx[i]];
        const int tl = r.tlen;
        if (no_trace) r.tlen = 0;                             // align.c:6086-6090
        r.pad = 0;
        memcpy(buf + o, &r, 40); o += 40;
        if (no_trace) continue;
        if (tbytes == 1) {
            for (int j = 0; j < tl; j++) {
                if (t[j] > 255) { free(buf); set_error("Compression of trace to bytes fails, value too big"); return 1; }
                buf[o++] = (uint8_t)t[j];
            }
        } else { memcpy(buf + o, t, 2 * (size_t)tl); o += 2 * (size_t)tl; }
    }
    *out = buf; *len = (int64_t)o;
    return 0;
}

}  // namespace damar

using namespace damar;

extern "C" {

damar_las_sink *damar_sink_new(void) { return reinterpret_cast<damar_las_sink *>(new LasSink()); }
void damar_sink_free(damar_las_sink *s) { delete reinterpret_cast<LasSink *>(s); }
void damar_sink_reset(damar_las_sink *s_) {
    LasSink *s = reinterpret_cast<LasSink *>(s_);
    s->ovl.clear(); s->toff.clear(); s->trace.clear();
}
int64_t damar_sink_count(const damar_las_sink *s) { return (int64_t) reinterpret_cast<const LasSink *>(s)->ovl.size(); }

void damar_sink_add(damar_las_sink *s_, const damar_overlaps *o) {
    LasSink *s = reinterpret_cast<LasSink *>(s_);
    const uint64_t base = s->trace.size();
    s->ovl.insert(s->ovl.end(), o->ovl, o->ovl + o->novl);
    for (int64_t i = 0; i < o->novl; i++) s->toff.push_back(base + o->trace_off[i]);
    s->trace.insert(s->trace.end(), o->trace, o->trace + o->ntrace);
}

int damar_sink_serialise(damar_las_sink *s_, int trace_spacing, int no_trace, int only_identity, int split_read,
                         uint8_t **buf1, int64_t *len1, uint8_t **buf2, int64_t *len2) {
    LasSink *s = reinterpret_cast<LasSink *>(s_);
    std::vector<uint32_t> idx;
    idx.reserve(s->ovl.size());
    for (uint32_t i = 0; i < s->ovl.size(); i++)
        if (!only_identity || s->ovl[i].aread == s->ovl[i].bread) idx.push_back(i);   // align.c:6185-6194
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return ovl_less(s->ovl[x], s->ovl[y]); });
    if (split_read < 0) {
        if (buf2) { *buf2 = nullptr; *len2 = 0; }
        return put_range(*s, idx, 0, idx.size(), trace_spacing, no_trace, buf1, len1);
    }
    size_t cut = 0;
    while (cut < idx.size() && s->ovl[idx[cut]].aread <= split_read) cut++;          // align.c:6330-6336
    if (put_range(*s, idx, 0, cut, trace_spacing, no_trace, buf1, len1)) return 1;
    return put_range(*s, idx, cut, idx.size(), trace_spacing, no_trace, buf2, len2);
}

void damar_free(void *p) { free(p); }

}  // extern "C"
