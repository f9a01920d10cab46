// path_fuse_core.cuh -- the multi-alignment post-pass of one read pair: Handle_Redundancies
// (reference dalign/filter.c:1804-2077) with Entwine (:1573), Fusion (:1691), Bridge (:1747),
// MapToTPAbove/Below (:1376, :1410), Check_Bridge (:1444), Compute_Bridge_Path (:1456) and the part of
// Compute_Alignment it needs (DIFF_TRACE: split_nd / trace_nd, dalign/align.c:4327-4640, 4734-4869).
//
// One THREAD handles one read pair that produced more than one local alignment: it is sequential list surgery on
// a few paths plus an exact O(ND) realignment of a ~5-trace-point box when two paths run narrowly parallel.  The
// code is __host__ __device__ and free of containers: paths are small structs whose traces live in a bump
// allocator the caller provides (ALLOC::take), the reads are seen through an accessor (SEQ: base at index i), and
// the divide-and-conquer of trace_nd runs on an explicit stack.  The product only calls it from the
// handle_redundancies kernel (pair_align.cu); tests/path_fuse_host.cpp compiles the same header for the CPU and checks
// "oracle + this post-pass == reference .las" (test infrastructure).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PF_HD __host__ __device__
#else
#define PF_HD
#endif

namespace damar {
namespace pf {

constexpr int TRACE_XOVR = 125;
constexpr int ND_STACK = 96;          // explicit stack of trace_nd: boxes are a few hundred bases, depth ~ 2 log2

struct Path {                         // Path (align.h:125-132); trace = uncompressed (d_i, b_i) pairs in the arena
    int abpos, bbpos, aepos, bepos, diffs, tlen;
    uint16_t *trace;
};

PF_HD inline int iabs(int x) { return x < 0 ? -x : x; }
PF_HD inline int imin(int a, int b) { return a < b ? a : b; }
PF_HD inline int imax(int a, int b) { return a > b ? a : b; }

// ---- trace-point walks ---------------------------------------------------------------------------
// first trace point of `p` at or above *x on the A (isA) or B axis; returns the other coordinate there
PF_HD inline int tp_above(const Path &p, int *x, bool isA, int ts) {
    int a = (p.abpos / ts) * ts, b = p.bbpos;
    for (int i = 1; i < p.tlen; i += 2) {
        a += ts; b += p.trace[i];
        if (a > p.aepos) a = p.aepos;
        if (isA) { if (a >= *x) { *x = a; return b; } }
        else if (b >= *x) { *x = b; return a; }
    }
    if (isA) { *x = a; return b; }
    *x = b; return a;
}
// last trace point at or below *x
PF_HD inline int tp_below(const Path &p, int *x, bool isA, int ts) {
    int a = ((p.aepos + (ts - 1)) / ts) * ts, b = p.bepos;
    for (int i = p.tlen - 1; i >= 0; i -= 2) {
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
PF_HD inline int entwine(const Path &jp, const Path &kp, int ts, int *where) {
    int mn = 10000, den = 0;
    int y2 = jp.bbpos, j = jp.abpos / ts;
    int b2 = kp.bbpos, k = kp.abpos / ts;
    int ac;
    if (jp.abpos == kp.abpos) {
        mn = iabs(y2 - b2);
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
    const int ae = imin(jp.aepos, kp.aepos);
    for (;;) {
        ac += ts;
        if (ac >= ae) break;
        y2 += jp.trace[j]; b2 += kp.trace[k];
        j += 2; k += 2;
        const int d = iabs(y2 - b2);
        if (d <= mn) { mn = d; if (d == 0) *where = ac; }
        den += 1;
    }
    if (jp.aepos == kp.aepos) {
        const int d = iabs(jp.bepos - kp.bepos);
        if (d <= mn) { mn = d; if (d == 0) *where = kp.aepos; }
    }
    return den == 0 ? -1 : mn;
}

// p1 := p1 up to the trace point at A coordinate ap, then p2 from there (filter.c:1691-1741).  false: arena dry.
template <class ALLOC>
PF_HD inline bool fusion(Path &p1, int ap, const Path &p2, int ts, ALLOC &al) {
    const int k1 = 2 * ((ap / ts) - (p1.abpos / ts));
    const int k2 = 2 * ((ap / ts) - (p2.abpos / ts));
    const int n = k1 + (p2.tlen - k2);
    uint16_t *t = al.take(n);
    if (t == nullptr) return false;
    int diff = 0, o = 0;
    for (int k = 0; k < k1; k += 2) { t[o++] = p1.trace[k]; t[o++] = p1.trace[k + 1]; diff += p1.trace[k]; }
    for (int k = k2; k < p2.tlen; k += 2) { t[o++] = p2.trace[k]; t[o++] = p2.trace[k + 1]; diff += p2.trace[k]; }
    p1.aepos = p2.aepos; p1.bepos = p2.bepos; p1.diffs = diff;
    p1.trace = t; p1.tlen = n;
    return true;
}
// p1 := p1 up to mid's start, mid, p3 from mid's end (filter.c:1747-1802)
template <class ALLOC>
PF_HD inline bool bridge(Path &p1, const Path &mid, const Path &p3, int ts, ALLOC &al) {
    const int k1 = 2 * ((mid.abpos / ts) - (p1.abpos / ts));
    const int k2 = (mid.aepos == p3.aepos) ? p3.tlen : 2 * ((mid.aepos / ts) - (p3.abpos / ts));
    const int n = k1 + mid.tlen + (p3.tlen - k2);
    uint16_t *t = al.take(n);
    if (t == nullptr) return false;
    int diff = 0, o = 0;
    for (int k = 0; k < k1; k += 2) { t[o++] = p1.trace[k]; t[o++] = p1.trace[k + 1]; diff += p1.trace[k]; }
    for (int k = 0; k < mid.tlen; k += 2) { t[o++] = mid.trace[k]; t[o++] = mid.trace[k + 1]; diff += mid.trace[k]; }
    for (int k = k2; k < p3.tlen; k += 2) { t[o++] = p3.trace[k]; t[o++] = p3.trace[k + 1]; diff += p3.trace[k]; }
    p1.aepos = p3.aepos; p1.bepos = p3.bepos; p1.diffs = diff;
    p1.trace = t; p1.tlen = n;
    return true;
}
PF_HD inline bool bridge_too_coarse(const Path &p, int ts) {           // Check_Bridge, filter.c:1444-1454
    if (ts <= TRACE_XOVR)
        for (int i = 0; i < p.tlen; i++) if (p.trace[i] > 250) return true;
    return false;
}

// ---- exact O(ND) alignment of a box with trace points (Compute_Alignment, task DIFF_TRACE) --------------
// A(ao + i), B(bo + i) are the bases of the box (i = 0 at its corner); `aorg` is the A coordinate of A(ao) in the
// frame whose multiples of ts are the trace points (the reference's A - wave->Aabs).
struct NdWaves {
    int *VF, *VB;
    uint16_t *trace;                 // indexed by 2 * (A coordinate / ts)
    int *stk;                        // ND_STACK x 5 ints: the explicit stack of trace_nd
};

// the middle snake of the box (align.c:4327-4496): returns D and the split point (*px, *py)
template <class SEQ>
PF_HD inline int split_nd(const SEQ &A, int ao, int M, const SEQ &B, int bo, int N, const NdWaves &w, int *px, int *py) {
    int *VF = w.VF, *VB = w.VB;
    int x, y, flow, blow, bhgh, a;

    y = 0;
    if (N < M) { while (y < N && B(bo + y) == A(ao + y)) y += 1; }
    else {
        while (y < M && B(bo + y) == A(ao + y)) y += 1;
        if (y >= M && N == M) { *px = *py = M; return 0; }
    }
    flow = 0;
    VF[0] = y; VF[-1] = -2;

    x = N - M;
    a = ao - x;
    y = N - 1;
    if (N > M) { while (y >= x && B(bo + y) == A(a + y)) y -= 1; }
    else { while (y >= 0 && B(bo + y) == A(a + y)) y -= 1; }
    blow = bhgh = -x;
    VB += x;
    VB[blow] = y; VB[blow - 1] = N + 1;

    for (int D = 1;; D += 1) {
        int k, r, am, ac, ap;
        // forward wave
        flow -= 1;
        am = ac = VF[flow - 1] = -2;
        a = ao + D;
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
            if (N < x) { while (y < N && B(bo + y) == A(a + y)) y += 1; }
            else { while (y < x && B(bo + y) == A(a + y)) y += 1; }
            VF[k] = y;
            a -= 1; x += 1;
        }
        // reverse wave
        bhgh += 1; blow -= 1;
        am = ac = VB[blow - 1] = N + 1;
        a = ao + bhgh;
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
            if (x > 0) { while (y >= x && B(bo + y) == A(a + y)) y -= 1; }
            else { while (y >= 0 && B(bo + y) == A(a + y)) y -= 1; }
            VB[k] = y;
            a -= 1; x += 1;
        }
    }
}

// divide and conquer over the middle snakes, accumulating (diffs, b-advance) per trace point (align.c:4498-4640).
// The two halves only add into disjoint or shared trace slots, so their order is free: explicit stack instead of recursion.
// Returns D of the whole box, or -1 when the stack overflowed.
template <class SEQ>
PF_HD inline int trace_nd(const SEQ &A, int ao0, int aorg0, int M0, const SEQ &B, int bo0, int N0, const NdWaves &w, int ts) {
    int (*stk)[5] = reinterpret_cast<int (*)[5]>(w.stk);
    int sp = 0, Dtop = 0;
    bool first = true;
    stk[sp][0] = ao0; stk[sp][1] = aorg0; stk[sp][2] = M0; stk[sp][3] = bo0; stk[sp][4] = N0; sp++;
    while (sp > 0) {
        sp--;
        const int ao = stk[sp][0], aorg = stk[sp][1], M = stk[sp][2], bo = stk[sp][3], N = stk[sp][4];
        int x, y, s, D;
        if (M <= 0) {
            y = (aorg / ts) << 1;
            w.trace[y] += N; w.trace[y + 1] += N;
            D = N;
        } else if (N <= 0) {
            x = aorg; y = x / ts; x = (y + 1) * ts - x; y <<= 1;
            for (s = M; s > 0; s -= x, x = ts) { if (x > s) x = s; w.trace[y] += x; y += 2; }
            D = M;
        } else {
            D = split_nd(A, ao, M, B, bo, N, w, &x, &y);
            if (D > 1) {
                s = aorg;
                if ((s / ts + 1) * ts - s >= x) { s = (s / ts) << 1; w.trace[s] += (D + 1) / 2; w.trace[s + 1] += y; }
                else {
                    if (sp >= ND_STACK) return -1;
                    stk[sp][0] = ao; stk[sp][1] = aorg; stk[sp][2] = x; stk[sp][3] = bo; stk[sp][4] = y; sp++;
                }
                s = aorg + x;
                if ((s / ts + 1) * ts - s >= M - x) { s = (s / ts) << 1; w.trace[s] += D / 2; w.trace[s + 1] += N - y; }
                else {
                    if (sp >= ND_STACK) return -1;
                    stk[sp][0] = ao + x; stk[sp][1] = aorg + x; stk[sp][2] = M - x; stk[sp][3] = bo + y; stk[sp][4] = N - y; sp++;
                }
            } else {
                int u, v;
                s = (D == 0 || M < N) ? x : x - 1;
                if (s > 0) {
                    u = aorg; v = u / ts; u = (v + 1) * ts - u;
                    for (v <<= 1; s > 0; s -= u, u = ts) { if (u > s) u = s; w.trace[v + 1] += u; v += 2; }
                }
                if (D != 0) {
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
            }
        }
        if (first) { Dtop = D; first = false; }
    }
    return Dtop;
}

// Compute_Alignment(align, work, DIFF_TRACE, ts) (align.c:4734-4869) for the box of `p`; `ashift` is the amount the
// reference has moved align->aseq down by (A coordinates of p are in that shifted frame).  false: arena dry / stack.
template <class SEQ, class ALLOC>
PF_HD inline bool compute_alignment_trace(Path &p, const SEQ &aseq, int ashift, const SEQ &bseq, int ts, ALLOC &al) {
    const int asub = p.aepos - p.abpos, bsub = p.bepos - p.bbpos;
    const int n = 2 * (((p.aepos + (ts - 1)) / ts - p.abpos / ts) + 1);
    uint16_t *st = al.take(n);
    if (st == nullptr) return false;
    for (int i = 0; i < n; i++) st[i] = 0;
    const int big = imax(asub, bsub);
    int *store = al.take_ints(4 * big + 14);
    if (store == nullptr) return false;
    NdWaves w;
    w.stk = al.take_ints(ND_STACK * 5);
    if (w.stk == nullptr) return false;
    w.VF = store + (big + 1) + 2;
    w.VB = w.VF + (2 * big + 3);
    w.trace = st - 2 * (p.abpos / ts);
    const int d = trace_nd(aseq, p.abpos - ashift, p.abpos, asub, bseq, p.bbpos, bsub, w, ts);
    if (d < 0) return false;
    p.diffs = d;
    if (st[n - 1] != 0) { st[n - 3] += st[n - 1]; st[n - 4] += st[n - 2]; }   // inserts on the last TP boundary
    p.trace = st; p.tlen = n - 2;
    return true;
}

// The exact realignment that bridges path1 into path2 (filter.c:1456-1571): `out` = the box two trace points
// beyond the crossing on either side.  `comp` only for the B-side call of a complement comparison, whose paths are
// in the un-complemented frames while the sequences are the complemented ones.
template <class SEQ, class ALLOC>
PF_HD inline bool compute_bridge_path(const Path &path1, const Path &path2, Path &out, const SEQ &aseq, int alen,
                                      const SEQ &bseq, int blen, int comp, int aovl, int bovl, int ts, ALLOC &al) {
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
    (void)bin; (void)bout;
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
    if (!compute_alignment_trace(out, aseq, boff, bseq, ts, al)) return false;
    if (comp) {
        uint16_t *t = out.trace;
        for (int i = 0, j = out.tlen - 2; i < j; i += 2, j -= 2) {
            uint16_t v = t[i]; t[i] = t[j]; t[j] = v;
            v = t[i + 1]; t[i + 1] = t[j + 1]; t[j + 1] = v;
        }
        out.abpos -= boff; out.aepos -= boff; alen -= boff;
        p = alen - out.abpos; out.abpos = alen - out.aepos; out.aepos = p;
        p = blen - out.bbpos; out.bbpos = blen - out.bepos; out.bepos = p;
    }
    return true;
}

// Handle_Redundancies (filter.c:1804-2077).  am[0..novls) are the pair's A paths in emission order, bm the B paths (or
// null).  Dead paths are marked with abpos < 0; returns the number of survivors, compacted to the front of both lists,
// or -1 when the arena ran dry (the lists are then unusable).
template <class SEQ, class ALLOC>
PF_HD inline int handle_redundancies(Path *am, Path *bm, int novls, int comp, int ts, const SEQ &aseq, int alen, const SEQ &bseq, int blen,
                                     ALLOC &al, bool do_bridge = true) {
    const bool hasB = (bm != nullptr);
    int awhen = 0, bwhen = 0, dist;

    // alignments that share a trace point: fuse, or drop the contained one
    for (int j = 1; j < novls; j++) {
        Path &jp = am[j];
        for (int k = j - 1; k >= 0; k--) {
            Path &kp = am[k];
            if (kp.abpos < 0) continue;
            if (jp.abpos < kp.abpos) {
                if (kp.abpos <= jp.aepos && kp.bbpos <= jp.bepos) {
                    dist = entwine(jp, kp, ts, &awhen);
                    if (dist == 0) {
                        if (kp.aepos > jp.aepos) {
                            if (hasB) {
                                if (comp) {
                                    dist = entwine(bm[k], bm[j], ts, &bwhen);
                                    if (dist != 0) continue;
                                    if (!fusion(jp, awhen, kp, ts, al) || !fusion(bm[k], bwhen, bm[j], ts, al)) return -1;
                                    bm[j] = bm[k];
                                } else {
                                    dist = entwine(bm[j], bm[k], ts, &bwhen);
                                    if (dist != 0) continue;
                                    if (!fusion(jp, awhen, kp, ts, al) || !fusion(bm[j], bwhen, bm[k], ts, al)) return -1;
                                }
                            } else if (!fusion(jp, awhen, kp, ts, al)) return -1;
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
                            if (kp.aepos > jp.aepos) { jp = kp; if (hasB) bm[j] = bm[k]; }
                        } else if (jp.aepos > kp.aepos) {
                            if (hasB) {
                                if (comp) {
                                    dist = entwine(bm[j], bm[k], ts, &bwhen);
                                    if (dist != 0) continue;
                                    if (!fusion(kp, awhen, jp, ts, al)) return -1;
                                    jp = kp;
                                    if (!fusion(bm[j], bwhen, bm[k], ts, al)) return -1;
                                } else {
                                    dist = entwine(bm[k], bm[j], ts, &bwhen);
                                    if (dist != 0) continue;
                                    if (!fusion(kp, awhen, jp, ts, al)) return -1;
                                    jp = kp;
                                    if (!fusion(bm[k], bwhen, bm[j], ts, al)) return -1;
                                    bm[j] = bm[k];
                                }
                            } else { if (!fusion(kp, awhen, jp, ts, al)) return -1; jp = kp; }
                            k = j;
                        } else { jp = kp; if (hasB) bm[j] = bm[k]; }
                        kp.abpos = -1;
                    }
                }
            }
        }
    }

    // alignments with a narrow parallel overlap: bridge them with an exact realignment of the crossing
    // (datander's Handle_Redundancies, scrub/tandem.c:767-850, is the loop above alone: do_bridge == false)
    for (int j = 1; do_bridge && j < novls; j++) {
        Path &jp = am[j];
        if (jp.abpos < 0) continue;
        for (int k = j - 1; k >= 0; k--) {
            Path &kp = am[k];
            if (kp.abpos < 0) continue;
            const bool jfirst = jp.abpos < kp.abpos;
            Path &path1 = jfirst ? jp : kp, &path2 = jfirst ? kp : jp;
            if (path2.abpos >= path1.aepos || path1.aepos >= path2.aepos || path1.bbpos >= path2.bbpos ||
                path2.bbpos >= path1.bepos || path1.bepos >= path2.bepos)
                continue;
            const int aovl = path1.aepos - path2.abpos, bovl = path1.bepos - path2.bbpos;
            if (iabs(aovl - bovl) > .2 * (aovl + bovl)) continue;          // the one floating-point compare, as written (:1980)
            int b1 = 0, b2 = 0;
            if (hasB) {
                if ((comp != 0) == jfirst) { b1 = k; b2 = j; } else { b1 = j; b2 = k; }
                if (bm[b1].abpos > bm[b2].abpos) continue;                 // "SYMFAIL" (:1992)
            }
            Path mid;
            if (!compute_bridge_path(path1, path2, mid, aseq, alen, bseq, blen, 0, aovl, bovl, ts, al)) return -1;
            if (bridge_too_coarse(mid, ts)) continue;
            const Path jback = jp, kback = kp;
            if (!bridge(path1, mid, path2, ts, al)) return -1;
            if (&path1 != &jp) jp = path1;
            kp.abpos = -1;
            if (hasB) {
                Path bmid;
                if (!compute_bridge_path(bm[b1], bm[b2], bmid, bseq, blen, aseq, alen, comp, bovl, aovl, ts, al)) return -1;
                if (bridge_too_coarse(bmid, ts)) { jp = jback; kp = kback; continue; }
                if (!bridge(bm[b1], bmid, bm[b2], ts, al)) return -1;
                if (b1 != j) bm[j] = bm[b1];
            }
        }
    }

    int no = 0;
    for (int j = 0; j < novls; j++)
        if (am[j].abpos >= 0) {
            if (hasB && no != j) bm[no] = bm[j];
            if (no != j) am[no] = am[j];
            no++;
        }
    return no;
}

}  // namespace pf
}  // namespace damar
