// path_fuse_host.cpp -- TEST INFRASTRUCTURE: applies the product's Handle_Redundancies (damar_b200/csrc/path_fuse_core.cuh,
// the very code the handle_redundancies kernel runs one pair per thread) on the CPU to the UNFUSED alignments the oracle buffered for one Match_Filter call, so that
// "oracle + path_fuse" can be compared with the reference's .las files on the CPU (tests/test_path_fuse.py).
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../damar_b200/csrc/path_fuse_core.cuh"
#include "../oracle/daligner_oracle.h"

using namespace damar;

namespace {
// the CPU stand-ins of what the kernel provides: reads as bytes, an arena that grows
struct ByteSeq {
    const uint8_t *p;
    int operator()(int i) const { return p[i]; }
};
struct HostArena {
    std::vector<std::unique_ptr<uint32_t[]>> chunks;
    uint16_t *take(int n) { chunks.emplace_back(new uint32_t[(size_t)n / 2 + 2]); return reinterpret_cast<uint16_t *>(chunks.back().get()); }
    int *take_ints(int n) { chunks.emplace_back(new uint32_t[(size_t)n + 2]); return reinterpret_cast<int *>(chunks.back().get()); }
};
struct FusePath {
    int abpos = 0, bbpos = 0, aepos = 0, bepos = 0, diffs = 0;
    std::vector<uint16_t> trace;
};
FusePath load(const ora_result *r, int64_t i) {
    FusePath p;
    const ora_ovl &o = r->ovl[i];
    p.abpos = o.abpos; p.bbpos = o.bbpos; p.aepos = o.aepos; p.bepos = o.bepos; p.diffs = o.diffs;
    p.trace.assign(r->trace + o.toff, r->trace + o.toff + o.tlen);
    return p;
}
// runs the product's post-pass (path_fuse_core.cuh) on vector-backed lists
void handle_redundancies(std::vector<FusePath> &am, std::vector<FusePath> *bm, int comp, int ts, const ora_block *ab, int ar,
                         const ora_block *bb, int br, bool do_bridge = true) {
    HostArena al;
    auto view = [](std::vector<FusePath> &v) {
        std::vector<pf::Path> o(v.size());
        for (size_t i = 0; i < v.size(); i++)
            o[i] = pf::Path{v[i].abpos, v[i].bbpos, v[i].aepos, v[i].bepos, v[i].diffs, (int)v[i].trace.size(), v[i].trace.data()};
        return o;
    };
    std::vector<pf::Path> pa = view(am), pb;
    if (bm) pb = view(*bm);
    const ByteSeq sa{reinterpret_cast<const uint8_t *>(ab->bases + ab->boff[ar])}, sb{reinterpret_cast<const uint8_t *>(bb->bases + bb->boff[br])};
    const int no = pf::handle_redundancies(pa.data(), bm ? pb.data() : nullptr, (int)am.size(), comp, ts, sa, ab->rlen[ar], sb, bb->rlen[br], al, do_bridge);
    auto back = [&](std::vector<pf::Path> &v, std::vector<FusePath> &dst) {
        std::vector<FusePath> o((size_t)no);
        for (int i = 0; i < no; i++) {
            o[i].abpos = v[i].abpos; o[i].bbpos = v[i].bbpos; o[i].aepos = v[i].aepos; o[i].bepos = v[i].bepos; o[i].diffs = v[i].diffs;
            o[i].trace.assign(v[i].trace, v[i].trace + v[i].tlen);
        }
        dst.swap(o);
    };
    back(pa, am);
    if (bm) back(pb, *bm);
}
}  // namespace

// Records [from, r->novl) were buffered by ONE ora_match_filter call (blocks ab / bb, bb complemented when comp): per pair
// all A records, then (if the pair reports both directions) all B records.  Rewrites them in place; returns the
// number of pairs whose alignment list changed.
extern "C" int64_t pf_fuse_result(ora_result *r, int64_t from, const ora_block *ab, const ora_block *bb, int self, int comp,
                                  int symmetric, int ts) {
    std::vector<ora_ovl> out(r->ovl, r->ovl + from);
    std::vector<uint16_t> tr(r->trace, r->trace + r->ttop);
    int64_t changed = 0;
    auto emit = [&](const ora_ovl &proto, const FusePath &p) {
        ora_ovl o = proto;
        o.abpos = p.abpos; o.bbpos = p.bbpos; o.aepos = p.aepos; o.bepos = p.bepos; o.diffs = p.diffs; o.tlen = (int32_t)p.trace.size();
        o.toff = (int64_t)tr.size();
        tr.insert(tr.end(), p.trace.begin(), p.trace.end());
        out.push_back(o);
    };
    for (int64_t i = from; i < r->novl;) {
        const int a = r->ovl[i].aread, b = r->ovl[i].bread;
        const int ar = a - ab->ufirst, br = b - bb->ufirst;
        const bool doB = symmetric && (ar != br || !self || !comp);
        int64_t e = i;
        while (e < r->novl && r->ovl[e].aread == a && r->ovl[e].bread == b) e++;
        int64_t na = e - i, nb = 0;
        if (doB) {
            if (a == b) { na /= 2; nb = na; }
            else { while (e < r->novl && r->ovl[e].aread == b && r->ovl[e].bread == a) e++; nb = e - i - na; }
        }
        std::vector<FusePath> am, bm;
        for (int64_t q = 0; q < na; q++) am.push_back(load(r, i + q));
        for (int64_t q = 0; q < nb; q++) bm.push_back(load(r, i + na + q));
        if (na > 1) handle_redundancies(am, nb > 1 ? &bm : nullptr, comp, ts, ab, ar, bb, br);
        else if (nb > 1) handle_redundancies(bm, nullptr, comp, ts, ab, ar, bb, br);
        if ((int64_t)am.size() != na || (int64_t)bm.size() != nb) changed++;
        for (size_t q = 0; q < am.size(); q++) emit(r->ovl[i], am[q]);
        for (size_t q = 0; q < bm.size(); q++) emit(r->ovl[i + na], bm[q]);
        i = e;
    }
    if ((int64_t)out.size() > r->omax) { r->omax = (int64_t)out.size() + 16; r->ovl = (ora_ovl *)realloc(r->ovl, sizeof(ora_ovl) * r->omax); }
    memcpy(r->ovl, out.data(), sizeof(ora_ovl) * out.size());
    r->novl = (int64_t)out.size();
    if ((int64_t)tr.size() > r->tmax) { r->tmax = (int64_t)tr.size() + 16; r->trace = (uint16_t *)realloc(r->trace, sizeof(uint16_t) * r->tmax); }
    memcpy(r->trace, tr.data(), sizeof(uint16_t) * tr.size());
    r->ttop = (int64_t)tr.size();
    return changed;
}

// datander (scrub/tandem.c:1140-1165): records [from, r->novl) were buffered by ora_match_self, per read (aread == bread) in
// alignment order; its Handle_Redundancies fuses and drops contained alignments but never bridges.
extern "C" int64_t pf_fuse_result_self(ora_result *r, int64_t from, const ora_block *blk, int ts) {
    std::vector<ora_ovl> out(r->ovl, r->ovl + from);
    std::vector<uint16_t> tr(r->trace, r->trace + r->ttop);
    int64_t changed = 0;
    for (int64_t i = from; i < r->novl;) {
        const int a = r->ovl[i].aread, ar = a - blk->ufirst;
        int64_t e = i;
        while (e < r->novl && r->ovl[e].aread == a) e++;
        std::vector<FusePath> am;
        for (int64_t q = i; q < e; q++) am.push_back(load(r, q));
        if (am.size() > 1) handle_redundancies(am, nullptr, 0, ts, blk, ar, blk, ar, false);
        if ((int64_t)am.size() != e - i) changed++;
        for (size_t q = 0; q < am.size(); q++) {
            ora_ovl o = r->ovl[i];
            const FusePath &p = am[q];
            o.abpos = p.abpos; o.bbpos = p.bbpos; o.aepos = p.aepos; o.bepos = p.bepos; o.diffs = p.diffs; o.tlen = (int32_t)p.trace.size();
            o.toff = (int64_t)tr.size();
            tr.insert(tr.end(), p.trace.begin(), p.trace.end());
            out.push_back(o);
        }
        i = e;
    }
    if ((int64_t)out.size() > r->omax) { r->omax = (int64_t)out.size() + 16; r->ovl = (ora_ovl *)realloc(r->ovl, sizeof(ora_ovl) * r->omax); }
    memcpy(r->ovl, out.data(), sizeof(ora_ovl) * out.size());
    r->novl = (int64_t)out.size();
    if ((int64_t)tr.size() > r->tmax) { r->tmax = (int64_t)tr.size() + 16; r->trace = (uint16_t *)realloc(r->trace, sizeof(uint16_t) * r->tmax); }
    memcpy(r->trace, tr.data(), sizeof(uint16_t) * tr.size());
    r->ttop = (int64_t)tr.size();
    return changed;
}
