// path_fuse_host.cpp -- TEST INFRASTRUCTURE: applies the product's host-side Handle_Redundancies
// (damar_b200/csrc/path_fuse.cu) to the UNFUSED alignments the oracle buffered for one Match_Filter call, so that
// "oracle + path_fuse" can be compared with the reference's .las files on the CPU (tests/test_path_fuse.py).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../damar_b200/csrc/path_fuse.cu"
#include "../oracle/daligner_oracle.h"

using namespace damar;

namespace {
struct OraSeqs : FuseSeqs {
    const ora_block *ab, *bb; int ar, br;
    void fetch() const override {}
    const uint8_t *aseq() const override { return reinterpret_cast<const uint8_t *>(ab->bases + ab->boff[ar]); }
    const uint8_t *bseq() const override { return reinterpret_cast<const uint8_t *>(bb->bases + bb->boff[br]); }
    int alen() const override { return ab->rlen[ar]; }
    int blen() const override { return bb->rlen[br]; }
};
FusePath load(const ora_result *r, int64_t i) {
    FusePath p;
    const ora_ovl &o = r->ovl[i];
    p.abpos = o.abpos; p.bbpos = o.bbpos; p.aepos = o.aepos; p.bepos = o.bepos; p.diffs = o.diffs;
    p.trace.assign(r->trace + o.toff, r->trace + o.toff + o.tlen);
    return p;
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
        OraSeqs sq; sq.ab = ab; sq.bb = bb; sq.ar = ar; sq.br = br;
        if (na > 1) handle_redundancies(am, nb > 1 ? &bm : nullptr, comp, ts, sq);
        else if (nb > 1) handle_redundancies(bm, nullptr, comp, ts, sq);
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
