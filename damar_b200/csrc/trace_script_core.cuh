// trace_script_core.cuh -- Compute_Trace_PTS's edit script for ONE trace-point segment, as scalar __host__ __device__ code.
//
// Replaces iter_np (reference dalign/align.c:4892-5261) in its GREEDIEST mode, the mode every caller of
// Compute_Trace_PTS uses (utils/LAshow.c:261, 346; dalign/filter.c:2018, 2048).  The cost of a segment does not depend on
// tie-breaking (trace_pts.cu computes it with a plain furthest-reaching recurrence); the SCRIPT does, so this file keeps
// the reference's wave order and choices:
//   * O(NP) waves: wave D holds, per diagonal k (A index = k + B index), the furthest B index F reachable with D
//     "non-del" differences; diagonals above the end diagonal del = M - N are swept downwards taking their upper
//     neighbour from the SAME wave, diagonals below it upwards, del itself last (align.c:5030-5052);
//   * the three-way choice of FS_MOVE (align.c:5002-5028) with its tie rules, and the direction code it records
//     (-1 / +1: from k -/+ 1 two waves back, 0: same diagonal one wave back, 2 / 4: from k -/+ 1 of this wave);
//   * the band grows by one diagonal on either side every second wave, up to +-dmax (align.c:4990-4996), dmax being the
//     largest d_i of the record's trace (Compute_Trace_PTS, align.c:5618-5626): more than dmax waves is the reference's
//     "Bad alignment between trace points" error;
//   * the path is recovered by reversing the direction codes from (D, del) to (0, 0) and replayed forwards, writing one
//     value per insertion / deletion (substitutions write nothing): bb + 1 + j for a step to a higher diagonal, -(ab + 1 +
//     k + j) for a step to a lower one, j = the B index the step leaves from, absolute read coordinates (align.c:5238-5254).
// Reads are seen through an accessor returning the base at an index RELATIVE to the segment start; like the reference the
// first waves probe index -1 (the base before the segment, or the read's terminator), so the accessor must answer there.
// One deviation: the reference's self-comparison clause (align.c:4956-4973, taken when both reads live in the same
// buffer) is never taken by its callers, which load A and B into separate buffers, and is not restated.
//
// np_middle restates middle_np (align.c:5263-5566) in the same mode: the same waves, then half of the path walked back;
// Compute_Trace_MID (align.c:5694-5830, the flavour corrector/LAcorrect.c:545 calls) runs iter_np between the middles of
// successive trace-point segments instead of between the trace points.
//
// Everything here is __host__ __device__: tests/trace_script_host.cpp runs the very same code on the CPU against the
// reference's own output (tests/golden/trace_script.json); the product only calls it from trace_pts.cu's kernel.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TS_HD __host__ __device__
#else
#define TS_HD
#endif

namespace damar {
namespace ts {

// scratch ints a segment with D waves and end diagonal del needs (two planes: furthest points, direction codes)
TS_HD inline int64_t scratch_ints(int D, int del) {
    const int ad = del < 0 ? -del : del;
    return 2ll * (int64_t)(D + 3) * (int64_t)(ad + 2 * D + 6);
}

// The two planes of one segment's waves: V(d, k) = furthest B index on diagonal k after d waves, H(d, k) = direction code.
struct Planes {
    int *V, *H; int W, kmin;
    TS_HD int &v(int d, int k) const { return V[(int64_t)(d + 2) * W + (k - kmin)]; }
    TS_HD int &h(int d, int k) const { return H[(int64_t)(d + 2) * W + (k - kmin)]; }
};

// The forward waves shared by iter_np and middle_np (align.c:4975-5066 and :5322-5393 are the same loop).  Returns the
// number of waves D at which diagonal del reached N, or -1: more than dmax (or Dcap) waves ("Bad alignment between trace
// points").  Dcap = number of waves the scratch was sized for.
template <class SEQA, class SEQB>
TS_HD inline int np_forward(const SEQA &A, int M, const SEQB &B, int N, int dmax, int Dcap, int *scratch, Planes &P) {
    const int del = M - N;
    const int lo0 = del < 0 ? del : 0, hi0 = del > 0 ? del : 0;
    P.kmin = lo0 - Dcap - 3; P.W = (hi0 - lo0) + 2 * Dcap + 6;            // diagonals [kmin, kmin + W)
    P.V = scratch; P.H = scratch + (int64_t)(Dcap + 3) * P.W;
    int low = lo0, hgh = hi0;
    const int posl = -dmax, posh = dmax;
    for (int k = low - 1; k <= hgh + 1; k++) { P.v(-2, k) = -2; P.v(-1, k) = -2; }
    P.v(-1, 0) = -1;
    low += 1; hgh -= 1;
    for (int D = 0;; D++) {
        if (D > dmax || D > Dcap) return -1;
        if ((D & 1) == 0) { if (low > posl) low -= 1; if (hgh < posh) hgh += 1; }
        P.v(D, hgh + 1) = -2; P.v(D, low - 1) = -2;
        auto move = [&](int k, int am, int ap, int mdir, int pdir) -> int {
            const int ac = P.v(D - 1, k) + 1;
            int j, e;
            if (ac < am) { if (ap < am) { e = mdir; j = am; } else { e = pdir; j = ap; } }
            else { if (ap < ac) { e = 0; j = ac; } else { e = pdir; j = ap; } }
            const int i = M - k, lim = N < i ? N : i;
            while (j < lim && B(j) == A(k + j)) j++;
            P.v(D, k) = j; P.h(D, k) = e;
            return j;
        };
        int j = -2;
        for (int k = hgh; k > del; k--) j = move(k, P.v(D - 2, k - 1), j + 1, -1, 4);
        j = -2;
        for (int k = low; k < del; k++) j = move(k, j, P.v(D - 2, k + 1) + 1, 2, 1);
        move(del, j, P.v(D, del + 1) + 1, 2, 4);
        if (P.v(D, del) >= N) return D;
    }
}

// iter_np (GREEDIEST).  Returns the number of values written to `out` (<= cap), or -1: more than dmax waves ("Bad
// alignment"), -2: `out` full.  *cost receives iter_np's return value D + |M - N|.
template <class SEQA, class SEQB>
TS_HD inline int np_script(const SEQA &A, int M, const SEQB &B, int N, int dmax, int Dcap, int ab, int bb, int *scratch,
                           int *out, int cap, int *cost) {
    const int del = M - N;
    Planes P;
    const int D = np_forward(A, M, B, N, dmax, Dcap, scratch, P);
    if (D < 0) return -1;
    if (cost) *cost = D + (del < 0 ? -del : del);
    // reverse the direction codes along the path (D, del) -> (0, 0) (GREEDIEST, align.c:5218-5232)
    P.h(0, 0) = 3;
    {
        int k = del, d = D, e = P.h(d, k);
        P.h(d, k) = 3;
        while (e != 3) {
            int h = k + e;
            if (e > 1) h -= 3; else if (e == 0) d -= 1; else d -= 2;
            const int m = P.h(d, h);
            P.h(d, h) = e; e = m; k = h;
        }
    }
    // replay forwards, one value per indel (align.c:5234-5258)
    int n = 0;
    {
        int k = 0, d = 0, e = P.h(0, 0);
        while (e != 3) {
            int h = k - e;
            const int c = P.v(d, k);
            if (e > 1) h += 3; else if (e == 0) d += 1; else d += 2;
            if (h != k) {
                if (n >= cap) return -2;
                out[n++] = (h > k) ? (bb + 1 + c) : -(ab + 1 + c + k);
            }
            k = h; e = P.h(d, h);
        }
    }
    return n;
}

// middle_np (GREEDIEST, align.c:5263-5566): the point of the segment's path that lies ceil(cost / 2) edges before its end,
// as offsets (*ma, *mb) from the segment start.  Returns 0, or -1 ("Bad alignment").
template <class SEQA, class SEQB>
TS_HD inline int np_middle(const SEQA &A, int M, const SEQB &B, int N, int dmax, int Dcap, int *scratch, int *ma, int *mb) {
    const int del = M - N;
    Planes P;
    int D = np_forward(A, M, B, N, dmax, Dcap, scratch, P);
    if (D < 0) return -1;
    int d = D + (del < 0 ? -del : del), k = del;
    for (const int f = d / 2; d > f; d--) {                // align.c:5548-5559
        const int e = P.h(D, k);
        int h = k + e;
        if (e > 1) h -= 3; else if (e == 0) D -= 1; else D -= 2;
        k = h;
    }
    *mb = P.v(D, k); *ma = k + P.v(D, k);
    return 0;
}

}  // namespace ts
}  // namespace damar
