// trace_script_host.cpp -- TEST INFRASTRUCTURE: runs the product's edit-script code (damar_b200/csrc/trace_script_core.cuh, the
// code the trace_point_scripts kernel runs one segment per thread) on the CPU over whole records, with Compute_Trace_PTS's
// segment loop (dalign/align.c:5650-5680), so that it can be compared with the reference's own output
// (tests/golden/trace_script.json, written by oracle/_ref/trace_pts_ref).
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../damar_b200/csrc/trace_script_core.cuh"

using namespace damar;

namespace {
struct Seq {                                   // base at an index relative to the segment start; the read's terminators are 4
    const int8_t *p;
    int operator()(int i) const { return p[i]; }
};
}  // namespace

// aseq / bseq point at base 0 of the two reads (a 4 before and after each read, as Load_Read leaves them); bseq already
// complemented for COMP records.  trace = the record's (d_i, b_i) values.  Returns the script length, or < 0.
extern "C" int ts_record_script(const int8_t *aseq, int alen, const int8_t *bseq, int blen, int abpos, int bbpos, int aepos, int bepos,
                                const uint16_t *trace, int tlen, int tspace, int *out, int cap, int *diffs) {
    int dmax = 0;
    for (int d = 1; d < tlen; d += 2) if (trace[d - 1] > dmax) dmax = trace[d - 1];
    std::vector<int> scratch;
    int n = 0, total = 0;
    int ab = abpos, ae = (abpos / tspace) * tspace, bb = bbpos;
    auto segment = [&](int ae_, int be_) -> int {
        if (ae_ > alen || be_ > blen) return -3;
        const int M = ae_ - ab, N = be_ - bb;
        scratch.resize((size_t)ts::scratch_ints(dmax, M - N));
        int cost = 0;
        const int k = ts::np_script(Seq{aseq + ab}, M, Seq{bseq + bb}, N, dmax, dmax, ab, bb, scratch.data(), out + n, cap - n, &cost);
        if (k < 0) return k;
        n += k; total += cost;
        return 0;
    };
    for (int i = 1; i < tlen - 2; i += 2) {
        ae += tspace;
        const int be = bb + trace[i];
        const int rc = segment(ae, be);
        if (rc) return rc;
        ab = ae; bb = be;
    }
    const int rc = segment(aepos, bepos);
    if (rc) return rc;
    if (diffs) *diffs = total;
    return n;
}
