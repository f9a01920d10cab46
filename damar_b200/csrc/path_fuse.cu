// path_fuse.cu -- the multi-alignment post-pass of one read pair: Handle_Redundancies
// (reference dalign/filter.c:1804-2077) with Entwine (:1573), Fusion (:1691), Bridge (:1747),
// MapToTPAbove/Below (:1376, :1410), Check_Bridge (:1444), Compute_Bridge_Path (:1456) and the part of
// Compute_Alignment it needs (DIFF_TRACE: split_nd / trace_nd, dalign/align.c:4327-4640, 4734-4869).
//
// Runs on the HOST over the few pairs that produced more than one local alignment (none on simulator
// reads, a handful per thousand pairs on damaged / repetitive reads): it is sequential list surgery on
// at most a few paths per pair plus an exact O(ND) realignment of a ~5-trace-point box when two paths
// run narrowly parallel.  SURVEY.md section 8 lists a device version under (f)-1.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "path_fuse.h"

namespace damar {
namespace {

constexpr int TRACE_XOVR = 125;

// ---- trace-point walks ---------------------------------------------------------------------------
// first trace point of `p` at or above *x on the A (isA) or B axis; returns the other coordinate there
int tp_above(const FusePath &p, int *x, bool isA, int ts) {
    int a = (p.abpos / ts) * ts, b = p.bbpos;
    for (size_t i = 1; i < p.trace.size(); i += 2) {
        a += ts; b += p.trace[i];
        if (a > p.aepos) a = p.aepos;
        if (isA) { if (a >= *x) { *x = a; return b; } }
        else if (b >= *x) { *x = b; return a; }
    }
    if (isA) { *x = a; return b; }
    *x = b; return a;
}
// last trace point at or below *x
int tp_below(const FusePath &p, int *x, bool isA, int ts) {
    int a = ((p.aepos + (ts - 1)) / ts) * ts, b = p.bepos;
    for (int i = (int)p.trace.size() - 1; i >= 0; i -= 2) {
        a -= ts; b -= p.trace[i];
        if (a < p.abpos) a = p.abpos;
        if (isA) { if (a <= *x) { *x = a; return b; } }
        else if (b <= *x) { *x = b; return a; }
    }
    if (isA) { *x = a; return b; }
    *x = b; return a;
}

// Minimum B distance of two paths over their common trace points; *where = an A coordinate where they meet
// (distance 0).  -1 when they share no interior trace point.  (filter.c:1573-1689)
int entwine(const FusePath &jp, const FusePath &kp, int ts, int *where) {
    int mn = 10000, den = 0;
    int y2 = jp.bbpos, j = jp.abpos / ts;
    int b2 = kp.bbpos, k = kp.abpos / ts;
    int ac;
    if (jp.abpos == kp.abpos) {
        mn = abs(y2 - b2);
        if (mn == 0) *where = kp.abpos;
    }
    if (j < k) {
        ac = k * ts;
        j = 1 + 2 * (k - j); k = 1;
        for (int i = 1; i < j; i += 2) y2 += jp.trace[i];
    } else {
        ac = j * ts;
        k = 1 + 2 * (j - k); j = 1;
        for (int i = 1; i < k; i += 2) b2 += kp.trace[i];
    }
    const int ae = std::min(jp.aepos, kp.aepos);
    for (;;) {
        ac += ts;
        if (ac >= ae) break;
        y2 += jp.trace[j]; b2 += kp.trace[k];
        j += 2; k += 2;
        const int d = abs(y2 - b2);
        if (d <= mn) { mn = d; if (d == 0) *where = ac; }
        den += 1;
    }
    if (jp.aepos == kp.aepos) {
        const int d = abs(jp.bepos - kp.bepos);
        if (d <= mn) { mn = d; if (d == 0) *where = kp.aepos; }
    }
    return den == 0 ? -1 : mn;
}

// p1 := p1 up to the trace point at A coordinate ap, then p2 from there (filter.c:1691-1741)
void fusion(FusePath &p1, int ap, const FusePath &p2, int ts) {
    const int k1 = 2 * ((ap / ts) - (p1.abpos / ts));
    const int k2 = 2 * ((ap / ts) - (p2.abpos / ts));
    std::vector<uint16_t> t;
    int diff = 0;
    for (int k = 0; k < k1; k += 2) { t.push_back(p1.trace[k]); t.push_back(p1.trace[k + 1]); diff += p1.trace[k]; }
    for (int k = k2; k < (int)p2.trace.size(); k += 2) { t.push_back(p2.trace[k]); t.push_back(p2.trace[k + 1]); diff += p2.trace[k]; }
    p1.aepos = p2.aepos; p1.bepos = p2.bepos; p1.diffs = diff;
    p1.trace.swap(t);
}
// p1 := p1 up to mid's start, mid, p3 from mid's end (filter.c:1747-1802)
void bridge(FusePath &p1, const FusePath &mid, const FusePath &p3, int ts) {
    const int k1 = 2 * ((mid.abpos / ts) - (p1.abpos / ts));
    const int k2 = (mid.aepos == p3.aepos) ? (int)p3.trace.size() : 2 * ((mid.aepos / ts) - (p3.abpos / ts));
    std::vector<uint16_t> t;
    int diff = 0;
    for (int k = 0; k < k1; k += 2) { t.push_back(p1.trace[k]); t.push_back(p1.trace[k + 1]); diff += p1.trace[k]; }
    for (int k = 0; k < (int)mid.trace.size(); k += 2) { t.push_back(mid.trace[k]); t.push_back(mid.trace[k + 1]); diff += mid.trace[k]; }
    for (int k = k2; k < (int)p3.trace.size(); k += 2) { t.push_back(p3.trace[k]); t.push_back(p3.trace[k + 1]); diff += p3.trace[k]; }
    p1.aepos = p3.aepos; p1.bepos = p3.bepos; p1.diffs = diff;
    p1.trace.swap(t);
}
bool bridge_too_coarse(const FusePath &p, int ts) {           // Check_Bridge, filter.c:1444-1454
    if (ts <= TRACE_XOVR)
        for (uint16_t v : p.trace) if (v > 250) return true;
    return false;
}

// ---- exact O(ND) alignment of a box with trace points (Compute_Alignment, task DIFF_TRACE) --------------
// A and B are views whose index 0 is the box corner; `aorg` is the A coordinate of A[0] in the frame whose
// multiples of ts are the trace points (the reference's A - wave->Aabs).
struct NdWaves {
    std::vector<int> store;
    int *VF = nullptr, *VB = nullptr;
    uint16_t *trace = nullptr;       // indexed by 2 * (A coordinate / ts)
};

// the middle snake of the box (align.c:4327-4496): returns D and the split point (*px, *py)
int split_nd(const uint8_t *A, int M, const uint8_t *B, int N, NdWaves &w, int *px, int *py) {
    int *VF = w.VF, *VB = w.VB;
    int x, y, flow, blow, bhgh;
    const uint8_t *a;

    y = 0;
    if (N < M) { while (y < N && B[y] == A[y]) y += 1; }
    else {
        while (y < M && B[y] == A[y]) y += 1;
        if (y >= M && N == M) { *px = *py = M; return 0; }
    }
    flow = 0;
    VF[0] = y; VF[-1] = -2;

    x = N - M;
    a = A - x;
    y = N - 1;
    if (N > M) { while (y >= x && B[y] == a[y]) y -= 1; }
    else { while (y >= 0 && B[y] == a[y]) y -= 1; }
    blow = bhgh = -x;
    VB += x;
    VB[blow] = y; VB[blow - 1] = N + 1;

    for (int D = 1;; D += 1) {
        int k, r, am, ac, ap;
        // forward wave
        flow -= 1;
        am = ac = VF[flow - 1] = -2;
        a = A + D;
        x = M - D;
        for (k = D; k >= flow; k--) {
            ap = ac; ac = am + 1; am = VF[k - 1];
            if (ac < am) y = (ap < am) ? am : ap;
            else y = (ap < ac) ? ac : ap;
            if (blow <= k && k <= bhgh) {
                r = VB[k];
                if (y > r) {
                    D = (D << 1) - 1;
                    if (ap > r) y = ap; else if (ac > r) y = ac; else y = r + 1;
                    *px = k + y; *py = y;
                    return D;
                }
            }
            if (N < x) { while (y < N && B[y] == a[y]) y += 1; }
            else { while (y < x && B[y] == a[y]) y += 1; }
            VF[k] = y;
            a -= 1; x += 1;
        }
        // reverse wave
        bhgh += 1; blow -= 1;
        am = ac = VB[blow - 1] = N + 1;
        a = A + bhgh;
        x = -bhgh;
        for (k = bhgh; k >= blow; k--) {
            ap = ac + 1; ac = am; am = VB[k - 1];
            if (ac > am) y = (ap > am) ? am : ap;
            else y = (ap > ac) ? ac : ap;
            if (flow <= k && k <= D) {
                r = VF[k];
                if (y <= r) {
                    D = (D << 1);
                    if (ap <= r) y = ap; else if (ac <= r) y = ac; else y = r;
                    *px = k + y; *py = y;
                    return D;
                }
            }
            y -= 1;
            if (x > 0) { while (y >= x && B[y] == a[y]) y -= 1; }
            else { while (y >= 0 && B[y] == a[y]) y -= 1; }
            VB[k] = y;
            a -= 1; x += 1;
        }
    }
}

// divide and conquer over the middle snakes, accumulating (diffs, b-advance) per trace point (align.c:4498-4640)
int trace_nd(const uint8_t *A, int aorg, int M, const uint8_t *B, int N, NdWaves &w, int ts) {
    int x, y, s;
    if (M <= 0) {
        y = (aorg / ts) << 1;
        w.trace[y] += N; w.trace[y + 1] += N;
        return N;
    }
    if (N <= 0) {
        x = aorg; y = x / ts; x = (y + 1) * ts - x; y <<= 1;
        for (s = M; s > 0; s -= x, x = ts) { if (x > s) x = s; w.trace[y] += x; y += 2; }
        return M;
    }
    const int D = split_nd(A, M, B, N, w, &x, &y);
    if (D > 1) {
        s = aorg;
        if ((s / ts + 1) * ts - s >= x) { s = (s / ts) << 1; w.trace[s] += (D + 1) / 2; w.trace[s + 1] += y; }
        else trace_nd(A, aorg, x, B, y, w, ts);
        s = aorg + x;
        if ((s / ts + 1) * ts - s >= M - x) { s = (s / ts) << 1; w.trace[s] += D / 2; w.trace[s + 1] += N - y; }
        else trace_nd(A + x, aorg + x, M - x, B + y, N - y, w, ts);
    } else {
        int u, v;
        s = (D == 0 || M < N) ? x : x - 1;
        if (s > 0) {
            u = aorg; v = u / ts; u = (v + 1) * ts - u;
            for (v <<= 1; s > 0; s -= u, u = ts) { if (u > s) u = s; w.trace[v + 1] += u; v += 2; }
        }
        if (D == 0) return D;
        if (M < N) y = ((aorg + x) / ts) << 1;
        else y = ((aorg + (x - 1)) / ts) << 1;
        w.trace[y] += 1;
        if (M <= N) w.trace[y + 1] += 1;
        s = M - x;
        if (s > 0) {
            u = aorg + x; v = u / ts; u = (v + 1) * ts - u;
            for (v <<= 1; s > 0; s -= u, u = ts) { if (u > s) u = s; w.trace[v + 1] += u; v += 2; }
        }
    }
    return D;
}

// Compute_Alignment(align, work, DIFF_TRACE, ts) (align.c:4734-4869) for the box of `p`; `ashift` is the amount the
// reference has moved align->aseq down by (A coordinates of p are in that shifted frame).
void compute_alignment_trace(FusePath &p, const uint8_t *aseq, int ashift, const uint8_t *bseq, int ts) {
    const int asub = p.aepos - p.abpos, bsub = p.bepos - p.bbpos;
    const int n = 2 * (((p.aepos + (ts - 1)) / ts - p.abpos / ts) + 1);
    std::vector<uint16_t> st((size_t)n, 0);
    NdWaves w;
    const int big = std::max(asub, bsub);
    w.store.assign((size_t)(4 * big + 6) + 8, 0);
    w.VF = w.store.data() + (big + 1) + 2;
    w.VB = w.VF + (2 * big + 3);
    w.trace = st.data() - 2 * (p.abpos / ts);
    p.diffs = trace_nd(aseq + (p.abpos - ashift), p.abpos, asub, bseq + p.bbpos, bsub, w, ts);
    if (st[n - 1] != 0) { st[n - 3] += st[n - 1]; st[n - 4] += st[n - 2]; }   // inserts on the last TP boundary
    st.resize((size_t)(n - 2));
    p.trace.swap(st);
}

// The exact realignment that bridges path1 into path2 (filter.c:1456-1571): `out` = the box two trace points
// beyond the crossing on either side.  `comp` only for the B-side call of a complement comparison, whose paths are
// in the un-complemented frames while the sequences are the complemented ones.
void compute_bridge_path(const FusePath &path1, const FusePath &path2, FusePath &out, const uint8_t *aseq, int alen,
                         const uint8_t *bseq, int blen, int comp, int aovl, int bovl, int ts) {
    int ain, aout, bin, bout, boff = 0, p;
    if (bovl > aovl) {
        bin = path2.bbpos; bout = path1.bepos;
        ain = tp_below(path1, &bin, false, ts);
        aout = tp_above(path2, &bout, false, ts);
    } else {
        ain = path2.abpos; aout = path1.aepos;
        bin = tp_below(path1, &ain, true, ts);
        bout = tp_above(path2, &aout, true, ts);
    }
    out.abpos = ain - 2 * ts;
    out.aepos = aout + 2 * ts;
    out.bbpos = tp_below(path1, &out.abpos, true, ts);
    out.bepos = tp_above(path2, &out.aepos, true, ts);
    if (comp) {
        boff = ts - out.aepos % ts;
        p = alen - out.abpos; out.abpos = alen - out.aepos; out.aepos = p;
        p = blen - out.bbpos; out.bbpos = blen - out.bepos; out.bepos = p;
        boff = boff - out.abpos % ts;      // align->aseq -= boff: the trace grid of the un-complemented frame
        out.abpos += boff; out.aepos += boff; alen += boff;
    }
    compute_alignment_trace(out, aseq, boff, bseq, ts);
    if (comp) {
        std::vector<uint16_t> &t = out.trace;
        for (int i = 0, j = (int)t.size() - 2; i < j; i += 2, j -= 2) { std::swap(t[i], t[j]); std::swap(t[i + 1], t[j + 1]); }
        out.abpos -= boff; out.aepos -= boff; alen -= boff;
        p = alen - out.abpos; out.abpos = alen - out.aepos; out.aepos = p;
        p = blen - out.bbpos; out.bbpos = blen - out.bepos; out.bepos = p;
    }
    (void)bin; (void)bout;
}

}  // namespace

// Handle_Redundancies (filter.c:1804-2077).  `am` are the pair's A paths in emission order, `bm` the B paths (or null).
// Dead paths are marked with abpos < 0; returns the number of survivors (compacted to the front of both lists).
int handle_redundancies(std::vector<FusePath> &am, std::vector<FusePath> *bm, int comp, int ts, const FuseSeqs &seqs) {
    const int novls = (int)am.size();
    const bool hasB = (bm != nullptr);
    int awhen = 0, bwhen = 0, dist;

    // alignments that share a trace point: fuse, or drop the contained one
    for (int j = 1; j < novls; j++) {
        FusePath &jp = am[j];
        for (int k = j - 1; k >= 0; k--) {
            FusePath &kp = am[k];
            if (kp.abpos < 0) continue;
            if (jp.abpos < kp.abpos) {
                if (kp.abpos <= jp.aepos && kp.bbpos <= jp.bepos) {
                    dist = entwine(jp, kp, ts, &awhen);
                    if (dist == 0) {
                        if (kp.aepos > jp.aepos) {
                            if (hasB) {
                                if (comp) {
                                    dist = entwine((*bm)[k], (*bm)[j], ts, &bwhen);
                                    if (dist != 0) continue;
                                    fusion(jp, awhen, kp, ts);
                                    fusion((*bm)[k], bwhen, (*bm)[j], ts);
                                    (*bm)[j] = (*bm)[k];
                                } else {
                                    dist = entwine((*bm)[j], (*bm)[k], ts, &bwhen);
                                    if (dist != 0) continue;
                                    fusion(jp, awhen, kp, ts);
                                    fusion((*bm)[j], bwhen, (*bm)[k], ts);
                                }
                            } else fusion(jp, awhen, kp, ts);
                            k = j;
                        }
                        kp.abpos = -1;
                    }
                }
            } else {                                   // kp.abpos <= jp.abpos
                if (jp.abpos <= kp.aepos && jp.bbpos <= kp.bepos) {
                    dist = entwine(kp, jp, ts, &awhen);
                    if (dist == 0) {
                        if (kp.abpos == jp.abpos) {
                            if (kp.aepos > jp.aepos) { jp = kp; if (hasB) (*bm)[j] = (*bm)[k]; }
                        } else if (jp.aepos > kp.aepos) {
                            if (hasB) {
                                if (comp) {
                                    dist = entwine((*bm)[j], (*bm)[k], ts, &bwhen);
                                    if (dist != 0) continue;
                                    fusion(kp, awhen, jp, ts);
                                    jp = kp;
                                    fusion((*bm)[j], bwhen, (*bm)[k], ts);
                                } else {
                                    dist = entwine((*bm)[k], (*bm)[j], ts, &bwhen);
                                    if (dist != 0) continue;
                                    fusion(kp, awhen, jp, ts);
                                    jp = kp;
                                    fusion((*bm)[k], bwhen, (*bm)[j], ts);
                                    (*bm)[j] = (*bm)[k];
                                }
                            } else { fusion(kp, awhen, jp, ts); jp = kp; }
                            k = j;
                        } else { jp = kp; if (hasB) (*bm)[j] = (*bm)[k]; }
                        kp.abpos = -1;
                    }
                }
            }
        }
    }

    // alignments with a narrow parallel overlap: bridge them with an exact realignment of the crossing
    for (int j = 1; j < novls; j++) {
        FusePath &jp = am[j];
        if (jp.abpos < 0) continue;
        for (int k = j - 1; k >= 0; k--) {
            FusePath &kp = am[k];
            if (kp.abpos < 0) continue;
            const bool jfirst = jp.abpos < kp.abpos;
            FusePath &path1 = jfirst ? jp : kp, &path2 = jfirst ? kp : jp;
            if (path2.abpos >= path1.aepos || path1.aepos >= path2.aepos || path1.bbpos >= path2.bbpos ||
                path2.bbpos >= path1.bepos || path1.bepos >= path2.bepos)
                continue;
            const int aovl = path1.aepos - path2.abpos, bovl = path1.bepos - path2.bbpos;
            if (abs(aovl - bovl) > .2 * (aovl + bovl)) continue;
            int b1 = 0, b2 = 0;
            if (hasB) {
                if ((comp != 0) == jfirst) { b1 = k; b2 = j; } else { b1 = j; b2 = k; }
                if ((*bm)[b1].abpos > (*bm)[b2].abpos) { printf("  SYMFAIL %d %d\n", j, k); continue; }
            }
            seqs.fetch();
            FusePath mid;
            compute_bridge_path(path1, path2, mid, seqs.aseq(), seqs.alen(), seqs.bseq(), seqs.blen(), 0, aovl, bovl, ts);
            if (bridge_too_coarse(mid, ts)) continue;
            const FusePath jback = jp, kback = kp;
            bridge(path1, mid, path2, ts);
            if (&path1 != &jp) jp = path1;
            kp.abpos = -1;
            if (hasB) {
                FusePath bmid;
                compute_bridge_path((*bm)[b1], (*bm)[b2], bmid, seqs.bseq(), seqs.blen(), seqs.aseq(), seqs.alen(), comp, bovl, aovl, ts);
                if (bridge_too_coarse(bmid, ts)) { jp = jback; kp = kback; continue; }
                bridge((*bm)[b1], bmid, (*bm)[b2], ts);
                if (b1 != j) (*bm)[j] = (*bm)[b1];
            }
        }
    }

    int no = 0;
    for (int j = 0; j < novls; j++)
        if (am[j].abpos >= 0) {
            if (hasB && no != j) (*bm)[no] = (*bm)[j];
            if (no != j) am[no] = am[j];
            no++;
        }
    am.resize((size_t)no);
    if (hasB) bm->resize((size_t)no);
    return no;
}

}  // namespace damar
