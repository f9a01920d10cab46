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

// Returns the number of values written to `out` (<= cap), or -1: more than dmax waves ("Bad alignment"), -2: `out` full.
// *cost receives iter_np's return value D + |M - N|.  Dcap = number of waves the scratch was sized for (>= the D the
// segment needs, which the cost kernel has already computed).
template <class SEQA, class SEQB>
TS_HD inline int np_script(const SEQA &A, int M, const SEQB &B, int N, int dmax, int Dcap, int ab, int bb, int *scratch,
                           int *out, int cap, int *cost) {
    const int del = M - N;
    const int lo0 = del < 0 ? del : 0, hi0 = del > 0 ? del : 0;
    const int kmin = lo0 - Dcap - 3, W = (hi0 - lo0) + 2 * Dcap + 6;      // diagonals [kmin, kmin + W)
    int *V = scratch, *H = scratch + (int64_t)(Dcap + 3) * W;
#define TSV(d, k) V[(int64_t)((d) + 2) * W + ((k) - kmin)]
#define TSH(d, k) H[(int64_t)((d) + 2) * W + ((k) - kmin)]
    int low = lo0, hgh = hi0;
    const int posl = -dmax, posh = dmax;
    for (int k = low - 1; k <= hgh + 1; k++) { TSV(-2, k) = -2; TSV(-1, k) = -2; }
    TSV(-1, 0) = -1;
    low += 1; hgh -= 1;
    int D;
    for (D = 0;; D++) {
        if (D > dmax || D > Dcap) return -1;
        if ((D & 1) == 0) { if (low > posl) low -= 1; if (hgh < posh) hgh += 1; }
        TSV(D, hgh + 1) = -2; TSV(D, low - 1) = -2;
        auto move = [&](int k, int am, int ap, int mdir, int pdir) -> int {
            const int ac = TSV(D - 1, k) + 1;
            int j, e;
            if (ac < am) { if (ap < am) { e = mdir; j = am; } else { e = pdir; j = ap; } }
            else { if (ap < ac) { e = 0; j = ac; } else { e = pdir; j = ap; } }
            const int i = M - k, lim = N < i ? N : i;
            while (j < lim && B(j) == A(k + j)) j++;
            TSV(D, k) = j; TSH(D, k) = e;
            return j;
        };
        int j = -2;
        for (int k = hgh; k > del; k--) j = move(k, TSV(D - 2, k - 1), j + 1, -1, 4);
        j = -2;
        for (int k = low; k < del; k++) j = move(k, j, TSV(D - 2, k + 1) + 1, 2, 1);
        move(del, j, TSV(D, del + 1) + 1, 2, 4);
        if (TSV(D, del) >= N) break;
    }
    if (cost) *cost = D + (del < 0 ? -del : del);
    // reverse the direction codes along the path (D, del) -> (0, 0) (GREEDIEST, align.c:5218-5232)
    TSH(0, 0) = 3;
    {
        int k = del, d = D, e = TSH(d, k);
        TSH(d, k) = 3;
        while (e != 3) {
            int h = k + e;
            if (e > 1) h -= 3; else if (e == 0) d -= 1; else d -= 2;
            const int m = TSH(d, h);
            TSH(d, h) = e; e = m; k = h;
        }
    }
    // replay forwards, one value per indel (align.c:5234-5258)
    int n = 0;
    {
        int k = 0, d = 0, e = TSH(0, 0);
        while (e != 3) {
            int h = k - e;
            const int c = TSV(d, k);
            if (e > 1) h += 3; else if (e == 0) d += 1; else d += 2;
            if (h != k) {
                if (n >= cap) return -2;
                out[n++] = (h > k) ? (bb + 1 + c) : -(ab + 1 + c + k);
            }
            k = h; e = TSH(d, h);
        }
    }
#undef TSV
#undef TSH
    return n;
}

}  // namespace ts
}  // namespace damar
