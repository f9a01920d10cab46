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

// Compute_Trace_MID's loop (dalign/align.c:5763-5822) around np_middle / np_script: the middle of every trace-point segment,
// iter_np from middle to middle, and the quirk of the last two calls (`d += iter_np(...); diffs += d;` counts the piece
// that ends at the last middle twice).
extern "C" int ts_record_mid(const int8_t *aseq, int alen, const int8_t *bseq, int blen, int abpos, int bbpos, int aepos, int bepos,
                             const uint16_t *trace, int tlen, int tspace, int *out, int cap, int *diffs) {
    int dmax = 0;
    for (int d = 1; d < tlen; d += 2) if (trace[d - 1] > dmax) dmax = trace[d - 1];
    std::vector<int> scratch;
    int n = 0, total = 0, last = 0;
    int ab = abpos, ae = (abpos / tspace) * tspace, bb = bbpos, as = abpos, bs = bbpos, af = abpos, bf = bbpos;
    auto piece = [&](int a0, int b0, int a1, int b1) -> int {                   // iter_np over [a0, a1) x [b0, b1)
        const int M = a1 - a0, N = b1 - b0;
        scratch.resize((size_t)ts::scratch_ints(dmax, M - N));
        int cost = 0;
        const int k = ts::np_script(Seq{aseq + a0}, M, Seq{bseq + b0}, N, dmax, dmax, a0, b0, scratch.data(), out + n, cap - n, &cost);
        if (k < 0) return k;
        n += k; total += cost; last = cost;
        return 0;
    };
    auto middle = [&](int ae_, int be_) -> int {                                // middle_np of the segment [ab, ae_) x [bb, be_)
        if (ae_ > alen || be_ > blen) return -3;
        const int M = ae_ - ab, N = be_ - bb;
        scratch.resize((size_t)ts::scratch_ints(dmax, M - N));
        int ma = 0, mb = 0;
        if (ts::np_middle(Seq{aseq + ab}, M, Seq{bseq + bb}, N, dmax, dmax, scratch.data(), &ma, &mb)) return -1;
        af = ab + ma; bf = bb + mb;
        return 0;
    };
    for (int i = 1; i < tlen - 2; i += 2) {
        ae += tspace;
        const int be = bb + trace[i];
        int rc = middle(ae, be);
        if (rc == 0) rc = piece(as, bs, af, bf);
        if (rc) return rc;
        ab = ae; bb = be; as = af; bs = bf;
    }
    int rc = middle(aepos, bepos);
    if (rc == 0) rc = piece(as, bs, af, bf);
    if (rc) return rc;
    const int before = last;
    rc = piece(af, bf, aepos, bepos);
    if (rc) return rc;
    if (diffs) *diffs = total + before;
    return n;
}
