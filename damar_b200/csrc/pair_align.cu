// pair_align.cu -- stages 3 and 4: per read pair, diagonal-band hit scoring with the -h
// threshold, then O(ND) forward/reverse wave extension with trace-point pebbles.
//
// Replaces report_thread (reference dalign/filter.c:2128-2511, incl. Diagonal_Span :2079-2110)
// and Local_Alignment / forward_wave / reverse_wave (dalign/align.c:1904-2097, 409-1122,
// 1126-1898).  One warp owns one (bread, aread) pair at a time:
//   * the pair's hits are walked in A-panels of 50 kbp (10 kbp overlap);
//   * pass 1 accumulates score[diag>>w] (covered bases per 2^w-diagonal bucket) with the
//     lanes taking 32 hits at a time and resolving same-bucket conflicts by lane order;
//   * pass 2 evaluates the threshold for 32 hits at once and then consumes candidates in hit
//     order, each one running a warp-wide local alignment unless an earlier alignment of the
//     pair already covers it (the `lasta` rule, kept here as a short interval list);
//   * the wave front lives one diagonal per lane (strided when wider than a warp); the three-way
//     furthest-reaching recurrence reads the previous wave, so all diagonals update in parallel;
//     the reference's loop-order tie-breaks (best / trim / clip bookkeeping) are reproduced
//     with ordered ballots; bases are compared 32 at a time on the 2-bit packed reads.
#include <algorithm>

#include "common.cuh"

namespace damar {
namespace {

constexpr int PANEL_SIZE = 50000;      // filter.c:73
constexpr int PANEL_OVERLAP = 10000;   // filter.c:74
constexpr int TRIM_LEN = 15;           // align.c:188
constexpr int PATH_LEN = 60;           // align.c:192
constexpr uint64_t PATH_INT = 0x0fffffffffffffffull;
constexpr uint64_t PATH_MASK61 = 0x1fffffffffffffffull;   // M = popcount of bits 0..60 (align.c:827-855)
constexpr int TRIM_MASK = 0x7fff;
constexpr int TRIM_MLAG = 200;
constexpr int WAVE_LAG = 30;
constexpr int FRACTION = 1000;
constexpr int MAX_LA_PER_PAIR = 64;
constexpr int IMAX = 0x7fffffff;

constexpr int AL_WARPS = 4;            // warps per CTA of the alignment kernel
constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// pair segmentation
// ---------------------------------------------------------------------------------------------

// slice ends of the reference's NTHREADS report threads (filter.c:2804-2816); ends[T-1] = n
__global__ void slice_ends(const uint64_t *__restrict__ keys, int64_t n, int nthreads, int nshift, int rshift,
                           int64_t *__restrict__ ends) {
    int t = threadIdx.x + 1;
    if (t > nthreads) return;
    if (t == nthreads) { ends[t - 1] = n; return; }
    int64_t p = (n * t) >> nshift;
    if (p > 0) {
        uint64_t d = keys[p - 1] >> rshift;          // bread of the previous hit
        int64_t l = p, r = n;                        // first index with bread > d
        while (l < r) { int64_t m = (l + r) >> 1; if ((keys[m] >> rshift) <= d) l = m + 1; else r = m; }
        p = l;
    }
    ends[t - 1] = p;
}

// flag[i] = 1 iff hit i starts a pair that report_thread would process: at least minhit hits
// (filter.c:2210-2221) and not the "exactly minhit hits at the very end of a slice" case.
__global__ void flag_work_pairs(const uint64_t *__restrict__ keys, int64_t n, int pbits, int minhit,
                                const int64_t *__restrict__ ends, int nends, uint32_t *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t pr = keys[i] >> pbits;
    uint32_t f = 0;
    if ((i == 0 || (keys[i - 1] >> pbits) != pr) && i + minhit - 1 < n && (keys[i + minhit - 1] >> pbits) == pr) {
        f = 1;
        const int64_t e = i + minhit;
        if (e >= n || (keys[e] >> pbits) != pr)      // exactly minhit hits: dropped if it ends a slice
            for (int t = 0; t < nends; t++) if (ends[t] == e) f = 0;
    }
    flag[i] = f;
}
__global__ void collect_work_pairs(const uint32_t *__restrict__ flag, const uint64_t *__restrict__ off, int64_t n,
                                   uint32_t *__restrict__ work) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    work[off[i]] = (uint32_t)i;
}

// ---------------------------------------------------------------------------------------------
// alignment kernel
// ---------------------------------------------------------------------------------------------

struct AlignArgs {
    // hits
    const uint64_t *keys; const uint32_t *bpos; int64_t nhits; int pbits, abits;
    const uint32_t *work; uint32_t nwork; unsigned int *next_work;
    const uint32_t *work_wi;           // optional: original work index of work[i] (second, strided pass)
    uint32_t *redo_sidx, *redo_wi; unsigned int *redo_count; uint8_t *pair_redo;   // fast pass: pairs handed to the strided pass
    // blocks
    const uint32_t *awords, *bwords; const uint32_t *aboff, *bboff; const int32_t *arlen, *brlen;
    int afirst, bfirst;
    // parameters
    int K, W, H, minhit, TS, minover, comp, self, symmetric;
    int ave_path; const int16_t *score_tab, *table_tab;
    // per-warp scratch
    int ring;                          // power of two
    int32_t *V, *HA, *HB, *HM, *HN, *NA, *NB; uint64_t *T;    // [nwarps][2][ring] except NA/NB [nwarps][ring]
    int4 *cells; int cells_cap;        // [nwarps][cells_cap]
    uint16_t *tbuf; int maxtp;         // [nwarps][8*maxtp]
    int32_t *bscore; int nbuck, dmin;  // [nwarps][2][nbuck]  (score, lastp) indexed by d - dmin
    // output
    damar_overlap *recs; uint64_t *rec_toff, *rec_seq; unsigned long long rec_cap; unsigned long long *rec_count;
    void *trace; int tbytes; unsigned long long trace_cap; unsigned long long *trace_count;
    unsigned long long *stats;         // [0]=LA calls [1]=multi-LA pairs [2]=wave cells [3]=overflow flags [4]=pairs
};

struct Path {
    int abpos, bbpos, aepos, bepos, diffs, tlen;
    uint16_t *trace;
};

// 32 bases (64 bits, first base most significant) starting at global base index s
__device__ __forceinline__ uint64_t window_at(const uint32_t *__restrict__ words, int64_t s) {
    const int64_t w = s >> 4;                       // arithmetic shift: s may be slightly negative (front pad)
    const int o = (int)(s & 15) * 2;
    const uint32_t x0 = __ldg(words + w), x1 = __ldg(words + w + 1), x2 = __ldg(words + w + 2);
    return ((uint64_t)__funnelshift_l(x1, x0, o) << 32) | __funnelshift_l(x2, x1, o);
}

// number of equal bases A[ax+i] == B[by+i], i = 0.., capped at maxn
__device__ __forceinline__ int slide_fwd(const uint32_t *__restrict__ aw, int64_t ax, const uint32_t *__restrict__ bw, int64_t by, int maxn) {
    int n = 0;
    while (n < maxn) {
        const uint64_t d = window_at(aw, ax + n) ^ window_at(bw, by + n);
        if (d != 0) { n += __clzll((long long)d) >> 1; break; }
        n += 32;
    }
    return n < maxn ? n : maxn;
}
// number of equal bases A[ax-1-i] == B[by-1-i], capped at maxn
__device__ __forceinline__ int slide_rev(const uint32_t *__restrict__ aw, int64_t ax, const uint32_t *__restrict__ bw, int64_t by, int maxn) {
    int n = 0;
    while (n < maxn) {
        const uint64_t d = window_at(aw, ax - n - 32) ^ window_at(bw, by - n - 32);
        if (d != 0) { n += (__ffsll((long long)d) - 1) >> 1; break; }
        n += 32;
    }
    return n < maxn ? n : maxn;
}
__device__ __forceinline__ uint64_t push_matches(uint64_t b, int n) {
    return (n >= 64) ? ~0ull : ((b << n) | ((1ull << n) - 1ull));
}

struct WarpLA {
    // scratch views for this warp
    int32_t *V[2], *HA[2], *HB[2], *HM[2], *HN[2], *NA, *NB; uint64_t *T[2];
    int4 *cells; int cells_cap; int ncell;
    int rmask;
    const int16_t *score_tab, *table_tab;
    int TS, PATH_AVE;
    int overflow;
    unsigned long long ncells_done;
};

// One wave-extension pass.  DIR=+1: forward_wave (align.c:409-1122); DIR=-1: reverse_wave (align.c:1126-1898).
// On return: end point (ex,ey), diffs, chain heads (ha,hb).  All scalars are warp-uniform.
template <int DIR>
__device__ void wave_pass(WarpLA &S, const uint32_t *__restrict__ aw, int64_t abase, int alen,
                          const uint32_t *__restrict__ bw, int64_t bbase, int blen,
                          int diag, int anti, int minp, int maxp, int aoff, int boff,
                          int &ex, int &ey, int &ediff, int &eha, int &ehb) {
    const int lane = threadIdx.x & 31;
    const int TS = S.TS, PATH_AVE = S.PATH_AVE, RM = S.rmask;
    int low = diag, hgh = diag, dif = 0, cur = 0;
    int more = 1;
    int aclip = (DIR > 0) ? IMAX : -IMAX, bclip = (DIR > 0) ? -IMAX : IMAX;
    int besta, besty, trima, trimy, trimd = 0, trimha = 0, trimhb = 1;
    int morea, morey, mored = 0, moreha = 0, morehb = 1, morem = -1, lasta;

    S.ncell = 0;
    besta = trima = morea = lasta = anti;
    besty = trimy = morey = (anti - diag) >> 1;

    // ---- wave 0 on the seed diagonal (align.c:491-626 / 1206-1340); uniform, lane 0 stores ----
    {
        const int k = diag;
        int y = (anti - k) >> 1, x = y + k, na, nb, ha, hb, ham, hbm, n;
        if (DIR > 0) {
            na = ((x + (TS - aoff)) / TS - 1) * TS + aoff;
            nb = ((y + (TS - boff)) / TS - 1) * TS + boff;
            if (lane == 0) { S.cells[0] = make_int4(-1, k, 0, na); S.cells[1] = make_int4(-1, k, 0, nb); }
            ha = 0; hb = 1; ham = na; hbm = nb; S.ncell = 2;
            na += TS; nb += TS;
            const int ra = alen - x, rb = blen - y;
            n = slide_fwd(aw, abase + x, bw, bbase + y, ra < rb ? ra : rb);
            y += n; x += n;
            if (y == blen) { more = 0; if (bclip < k) bclip = k; }
            else if (x == alen) { more = 0; aclip = k; }
            while (x >= na) { if (lane == 0) S.cells[S.ncell] = make_int4(ha, k, 0, na); ha = S.ncell++; ham = na; na += TS; }
            while (y >= nb) { if (lane == 0) S.cells[S.ncell] = make_int4(hb, k, 0, nb); hb = S.ncell++; hbm = nb; nb += TS; }
        } else {
            na = ((x + (TS - aoff) - 1) / TS - 1) * TS + aoff;
            nb = ((y + (TS - boff) - 1) / TS - 1) * TS + boff;
            if (lane == 0) { S.cells[0] = make_int4(-1, k, 0, x); S.cells[1] = make_int4(-1, k, 0, y); }
            ha = 0; hb = 1; ham = x; hbm = y; S.ncell = 2;
            n = slide_rev(aw, abase + x, bw, bbase + y, x < y ? x : y);
            y -= n; x -= n;
            if (y == 0) { more = 0; if (bclip > k) bclip = k; }
            else if (x == 0) { more = 0; aclip = k; }
            while (x <= na) { if (lane == 0) S.cells[S.ncell] = make_int4(ha, k, 0, na); ha = S.ncell++; ham = na; na -= TS; }
            while (y <= nb) { if (lane == 0) S.cells[S.ncell] = make_int4(hb, k, 0, nb); hb = S.ncell++; hbm = nb; nb -= TS; }
        }
        const int c = (y << 1) + k;
        if (DIR > 0 ? (c > besta) : (c < besta)) { besta = trima = lasta = c; besty = trimy = y; trimha = ha; trimhb = hb; }
        if (lane == 0) {
            const int s = k & RM;
            S.V[cur][s] = c; S.T[cur][s] = PATH_INT; S.HA[cur][s] = ha; S.HB[cur][s] = hb; S.HM[cur][s] = ham; S.HN[cur][s] = hbm;
            S.NA[s] = na; S.NB[s] = nb;
        }
        __syncwarp();
    }

    for (;;) {
        // ---- boundary clipping (align.c:628-658, 943-975 / 1342-1372, 1656-1684) ----
        if (more == 0) {
            bool inside;
            if (DIR > 0) inside = (besty < blen) && (besta - besty < alen);
            else inside = (besty > 0) && (besta - besty > 0);
            if (inside) more = 1;
            // first clip examined: A for forward / A for reverse, both on their own side
            for (int side = 0; side < 2; side++) {
                int kc; bool hit;
                if (DIR > 0) { kc = side == 0 ? aclip : bclip; hit = side == 0 ? (hgh >= aclip) : (low <= bclip); }
                else         { kc = side == 0 ? aclip : bclip; hit = side == 0 ? (low <= aclip) : (hgh >= bclip); }
                if (!hit) continue;
                if (DIR > 0) { if (side == 0) hgh = kc - 1; else low = kc + 1; }
                else         { if (side == 0) low = kc + 1; else hgh = kc - 1; }
                const int s = kc & RM;
                const int mc = __popcll(S.T[cur][s] & PATH_MASK61);
                if (morem <= mc) {
                    morem = mc; morea = S.V[cur][s]; morey = (morea - kc) / 2; mored = dif;
                    moreha = S.HA[cur][s]; morehb = S.HB[cur][s];
                }
            }
            aclip = (DIR > 0) ? IMAX : -IMAX; bclip = (DIR > 0) ? -IMAX : IMAX;
        }
        // ---- prune lagging ends (align.c:977-986 / 1686-1695) ----
        if (dif > 0) {
            if (DIR > 0) {
                const int n = besta - WAVE_LAG;
                while (hgh >= low)
                    if (S.V[cur][hgh & RM] < n) hgh -= 1;
                    else { while (S.V[cur][low & RM] < n) low += 1; break; }
            } else {
                const int n = besta + WAVE_LAG;
                while (hgh >= low)
                    if (S.V[cur][hgh & RM] > n) hgh -= 1;
                    else { while (S.V[cur][low & RM] > n) low += 1; break; }
            }
            if (hgh >= low) S.ncells_done += (unsigned)(hgh - low + 1);
        }
        if (!(more && (DIR > 0 ? (lasta >= besta - TRIM_MLAG) : (lasta <= besta + TRIM_MLAG)))) break;
        if (hgh < low) break;            // every diagonal ran off a read end (the reference reads garbage here)

        // ---- next wave: widen (align.c:756-772 / 1455-1475) ----
        const int NEG = (DIR > 0) ? -1 : IMAX;              // "no point" value
        low -= 1; hgh += 1;
        if (hgh - low + 4 > RM) { S.overflow = 1; break; }   // wave wider than the ring
        if (lane == 0) {
            if (low >= minp) { S.NA[low & RM] = S.NA[(low + 1) & RM]; S.NB[low & RM] = S.NB[(low + 1) & RM]; S.V[cur][low & RM] = NEG; }
            if (hgh <= maxp) { S.NA[hgh & RM] = S.NA[(hgh - 1) & RM]; S.NB[hgh & RM] = S.NB[(hgh - 1) & RM]; S.V[cur][hgh & RM] = NEG; }
        }
        if (low < minp) low += 1;
        if (hgh > maxp) hgh -= 1;
        if (lane == 0) { S.V[cur][(hgh + 1) & RM] = NEG; S.V[cur][(low - 1) & RM] = NEG; }
        __syncwarp();
        dif += 1;
        const int prv = cur; cur ^= 1;

        int run_best = besta;                                // running besta in visiting order
        for (int base = 0; base <= hgh - low; base += 32) {
            const int k = (DIR > 0) ? (hgh - base - lane) : (low + base + lane);
            const bool act = (DIR > 0) ? (k >= low) : (k <= hgh);
            int c = (DIR > 0) ? -IMAX : IMAX, y = 0, ha = 0, hb = 0, m = 0;
            uint64_t b = 0;
            bool clipA = false, clipB = false;
            int na = 0, nb = 0, ham = 0, hbm = 0;
            if (act) {
                const int vm = S.V[prv][(k - 1) & RM], vc = S.V[prv][k & RM], vp = S.V[prv][(k + 1) & RM];
                int src;
                if (DIR > 0) {                               // align.c:793-825
                    if (vc < vm) { if (vm < vp) { c = vp + 1; src = k + 1; } else { c = vm + 1; src = k - 1; } }
                    else if (vc < vp) { c = vp + 1; src = k + 1; }
                    else { c = vc + 2; src = k; }
                } else {                                     // align.c:1503-1535
                    if (vc > vp) { if (vp > vm) { c = vm - 1; src = k - 1; } else { c = vp - 1; src = k + 1; } }
                    else if (vc > vm) { c = vm - 1; src = k - 1; }
                    else { c = vc - 2; src = k; }
                }
                const int s = src & RM;
                b = S.T[prv][s] << 1; ha = S.HA[prv][s]; hb = S.HB[prv][s]; ham = S.HM[prv][s]; hbm = S.HN[prv][s];
                y = (c - k) >> 1;
                int x = y + k, n;
                if (DIR > 0) {
                    const int ra = alen - x, rb = blen - y;
                    n = slide_fwd(aw, abase + x, bw, bbase + y, ra < rb ? ra : rb);
                    y += n; x += n;
                    if (y == blen) clipB = true; else if (x == alen) clipA = true;
                } else {
                    n = slide_rev(aw, abase + x, bw, bbase + y, x < y ? x : y);
                    y -= n; x -= n;
                    if (y == 0) clipB = true; else if (x == 0) clipA = true;
                }
                b = push_matches(b, n);
                c = (y << 1) + k;
                m = __popcll(b & PATH_MASK61);
                na = S.NA[k & RM]; nb = S.NB[k & RM];
            }
            // ---- pebbles (align.c:859-909 / 1568-1618): allocate cells with ordered ballots ----
            {
                const int x = y + k;
                for (;;) {
                    const bool need = act && (DIR > 0 ? (x >= na) : (x <= na));
                    if (!__any_sync(FULL, need)) break;
                    const bool push = need && (DIR > 0 ? (ham < na) : (ham > na));
                    const unsigned pm = __ballot_sync(FULL, push);
                    if (push) {
                        const int id = S.ncell + __popc(pm & ((1u << lane) - 1u));
                        if (id < S.cells_cap) S.cells[id] = make_int4(ha, k, dif, na);
                        ha = id; ham = na;
                    }
                    S.ncell += __popc(pm);
                    if (need) na += DIR * TS;
                }
                for (;;) {
                    const bool need = act && (DIR > 0 ? (y >= nb) : (y <= nb));
                    if (!__any_sync(FULL, need)) break;
                    const bool push = need && (DIR > 0 ? (hbm < nb) : (hbm > nb));
                    const unsigned pm = __ballot_sync(FULL, push);
                    if (push) {
                        const int id = S.ncell + __popc(pm & ((1u << lane) - 1u));
                        if (id < S.cells_cap) S.cells[id] = make_int4(hb, k, dif, nb);
                        hb = id; hbm = nb;
                    }
                    S.ncell += __popc(pm);
                    if (need) nb += DIR * TS;
                }
                if (S.ncell >= S.cells_cap) S.overflow = 1;
            }
            // ---- best / lasta / trim in visiting order (align.c:911-928 / 1620-1637) ----
            {
                int pre = c;                                  // inclusive prefix extremum over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL, pre, o);
                    if (lane >= o) pre = (DIR > 0) ? max(pre, t) : min(pre, t);
                }
                int excl = __shfl_up_sync(FULL, pre, 1);
                if (lane == 0) excl = run_best; else excl = (DIR > 0) ? max(excl, run_best) : min(excl, run_best);
                const bool setter = act && (DIR > 0 ? (c > excl) : (c < excl));
                const bool good = setter && (m >= PATH_AVE);
                bool trim = false;
                if (good) {
                    const int lo15 = (int)(b & TRIM_MASK), hi15 = (int)((b >> TRIM_LEN) & TRIM_MASK);
                    if (S.table_tab[lo15] >= 0 && S.table_tab[hi15] + S.score_tab[lo15] >= 0) trim = true;
                }
                const unsigned sm = __ballot_sync(FULL, setter), gm = __ballot_sync(FULL, good), tm = __ballot_sync(FULL, trim);
                if (sm) { const int l = 31 - __clz(sm); besta = __shfl_sync(FULL, c, l); besty = __shfl_sync(FULL, y, l); }
                if (gm) { const int l = 31 - __clz(gm); lasta = __shfl_sync(FULL, c, l); }
                if (tm) {
                    const int l = 31 - __clz(tm);
                    trima = __shfl_sync(FULL, c, l); trimy = __shfl_sync(FULL, y, l); trimd = dif;
                    trimha = __shfl_sync(FULL, ha, l); trimhb = __shfl_sync(FULL, hb, l);
                }
                run_best = besta;
            }
            // ---- clip bookkeeping in visiting order (align.c:836-851 / 1545-1560) ----
            {
                const unsigned am = __ballot_sync(FULL, clipA), bm = __ballot_sync(FULL, clipB);
                if (am | bm) more = 0;
                if (am) { const int l = 31 - __clz(am); aclip = (DIR > 0) ? (hgh - base - l) : (low + base + l); }   // last visited
                if (bm) {
                    const int l = __ffs(bm) - 1;                                                                    // first visited is extreme
                    const int kb = (DIR > 0) ? (hgh - base - l) : (low + base + l);
                    if (DIR > 0) { if (bclip < kb) bclip = kb; } else { if (bclip > kb) bclip = kb; }
                }
            }
            if (act) {
                const int s = k & RM;
                S.V[cur][s] = c; S.T[cur][s] = b; S.HA[cur][s] = ha; S.HB[cur][s] = hb; S.HM[cur][s] = ham; S.HN[cur][s] = hbm;
                S.NA[s] = na; S.NB[s] = nb;
            }
        }
        __syncwarp();
        if (S.overflow) break;
    }

    if (morem >= 0) { ex = morea - morey; ey = morey; ediff = mored; eha = moreha; ehb = morehb; }   // REACH = 1 (daligner.c:951)
    else { ex = trima - trimy; ey = trimy; ediff = trimd; eha = trimha; ehb = trimhb; }
}

// Walk a pebble chain from its head back to the root and write the trace pairs
// (align.c:1001-1111 forward, 1711-1880 reverse).  Lane 0 only.  `isA` selects the A chain
// (marks are A coordinates) or the B chain.  Forward: pairs are appended at tr[0..); reverse:
// pairs are prepended before tr[0] and the first partial segment is merged into tr[0], tr[1].
// Returns the number of values written (forward) or prepended (reverse).
__device__ int emit_chain_fwd(const int4 *cells, int head, bool isA, int anti, int ex, int ey, int ediff, uint16_t *tr) {
    // count chain length first so that pairs can be written front to back while walking backwards
    int n = 0;
    for (int h = head; h >= 0; h = cells[h].x) n++;
    // cells in forward order are c_0 (root) .. c_{n-1} (head); pair i (i>=1) = (diff_i - diff_{i-1}, p_i - p_{i-1})
    int len = 2 * (n - 1);
    int h = head;
    int4 cur = cells[h];
    const int lastk = cur.y;
    int lastp, laste;
    {
        // coordinate "other" of head cell: A chain -> b = mark - diag ; B chain -> a = mark + diag
        for (int i = n - 1; i >= 1; i--) {
            const int4 prv = cells[cur.x];
            const int pc = isA ? (cur.w - cur.y) : (cur.w + cur.y);
            int pp, pe;
            if (i == 1) { pp = isA ? ((anti - prv.y) / 2) : ((anti + prv.y) / 2); pe = 0; }
            else { pp = isA ? (prv.w - prv.y) : (prv.w + prv.y); pe = prv.z; }
            tr[2 * (i - 1)] = (uint16_t)(cur.z - pe);
            tr[2 * (i - 1) + 1] = (uint16_t)(pc - pp);
            cur = prv;
        }
        // recompute last logged point from the head cell
        const int4 hc = cells[head];
        if (n == 1) { lastp = isA ? ((anti - hc.y) / 2) : ((anti + hc.y) / 2); laste = 0; }
        else { lastp = isA ? (hc.w - hc.y) : (hc.w + hc.y); laste = hc.z; }
    }
    if (isA) {
        if (lastp + lastk != ex) { tr[len++] = (uint16_t)(ediff - laste); tr[len++] = (uint16_t)(ey - lastp); }
        else if (lastp != ey) { tr[len - 1] = (uint16_t)(tr[len - 1] + (ey - lastp)); tr[len - 2] = (uint16_t)(tr[len - 2] + (ediff - laste)); }
    } else {
        if (lastp - lastk != ey) { tr[len++] = (uint16_t)(ediff - laste); tr[len++] = (uint16_t)(ex - lastp); }
        else if (lastp != ex) { tr[len - 1] = (uint16_t)(tr[len - 1] + (ex - lastp)); tr[len - 2] = (uint16_t)(tr[len - 2] + (ediff - laste)); }
    }
    return len;
}

// reverse pass: cells in chain order root=c_0 (mark = seed coordinate) .. head.  `flen` = number of
// values the forward pass left at tr[0..); returns number of values prepended (result starts at tr - ret).
__device__ int emit_chain_rev(int4 *cells, int head, bool isA, int ex, int ey, int ediff, int TS, int off,
                              int flen, uint16_t *tr) {
    int n = 0;
    for (int h = head; h >= 0; h = cells[h].x) n++;
    // gather the chain in root->head order is needed because merging happens at the root end;
    // walk with an index stack emulated by repeated traversal is O(n^2); instead reverse links in place.
    int a = -1, b;
    for (int h = head; h >= 0; h = b) { int4 c = cells[h]; b = c.x; c.x = a; cells[h] = c; a = h; }
    int h = a;                                       // root
    int4 c = cells[h];
    int k = c.y;
    int pb = isA ? (c.w - k) : (c.w + k);            // "other" coordinate of the running point
    int e = 0, at = 0;
    const int own = isA ? (pb + k) : (pb - k);       // coordinate on this chain's own axis
    if (own % TS != off) {                           // seed not on a trace boundary: merge first segment
        h = c.x;
        int pa, d;
        if (h < 0) { pa = isA ? ey : ex; d = ediff; }
        else { c = cells[h]; k = c.y; pa = isA ? (c.w - k) : (c.w + k); d = c.z; }
        if (flen == 0) {
            tr[--at] = (uint16_t)(pb - pa);
            tr[--at] = isA ? (uint16_t)(d - e) : (uint16_t)(pb - pa);   // sic (align.c:1843-1844)
        } else { tr[1] = (uint16_t)(tr[1] + (pb - pa)); tr[0] = (uint16_t)(tr[0] + (d - e)); }
        pb = pa; e = d;
    }
    if (h >= 0) {
        for (h = c.x; h >= 0; h = c.x) {
            c = cells[h]; k = c.y;
            const int pa = isA ? (c.w - k) : (c.w + k);
            tr[--at] = (uint16_t)(pb - pa);
            tr[--at] = (uint16_t)(c.z - e);
            pb = pa; e = c.z;
        }
        const int endo = isA ? ey : ex;              // end point on the "other" axis
        const int endown = isA ? ex : ey;
        const int lastown = isA ? (pb + k) : (pb - k);
        if (lastown != endown) { tr[--at] = (uint16_t)(pb - endo); tr[--at] = (uint16_t)(ediff - e); }
        else if (pb != endo) { tr[at + 1] = (uint16_t)(tr[at + 1] + (pb - endo)); tr[at] = (uint16_t)(tr[at] + (ediff - e)); }
    }
    (void)n;
    return -at;
}

// ---------------------------------------------------------------------------------------------
// Register-resident wave pass: the whole wave front (<= 32 live diagonals) lives in registers, one
// diagonal per lane (lane = k mod 32), neighbours are fetched with shuffles, and nothing but the
// pebble cells touches memory.  Same arithmetic and visiting-order semantics as wave_pass; returns
// false (S.overflow = 2) when the front grows beyond 32 diagonals so that the pair is redone by
// the strided scratch-memory kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned rotr32(unsigned x, int s) { return __funnelshift_r(x, x, s & 31); }

template <int DIR>
__device__ void wave_pass_fast(WarpLA &S, const uint32_t *__restrict__ aw, int64_t abase, int alen,
                               const uint32_t *__restrict__ bw, int64_t bbase, int blen,
                               int diag, int anti, int minp, int maxp, int aoff, int boff,
                               int &ex, int &ey, int &ediff, int &eha, int &ehb) {
    const int lane = threadIdx.x & 31;
    const int TS = S.TS, PATH_AVE = S.PATH_AVE;
    const int NEG = (DIR > 0) ? -1 : IMAX;
    int low = diag, hgh = diag, dif = 0, more = 1;
    int aclip = (DIR > 0) ? IMAX : -IMAX, bclip = (DIR > 0) ? -IMAX : IMAX;
    int besta, besty, trima, trimy, trimd = 0, trimha = 0, trimhb = 1;
    int morea, morey, mored = 0, moreha = 0, morehb = 1, morem = -1, lasta;
    // per-lane diagonal state
    int V, HA, HB, HM, HN, NA, NB; uint64_t T;
    unsigned long long cells_done = 0;

    S.ncell = 0;
    besta = trima = morea = lasta = anti;
    besty = trimy = morey = (anti - diag) >> 1;

    {   // ---- wave 0 (uniform) ----
        const int k = diag;
        int y = (anti - k) >> 1, x = y + k, na, nb, ha, hb, ham, hbm, n;
        if (DIR > 0) {
            na = ((x + (TS - aoff)) / TS - 1) * TS + aoff;
            nb = ((y + (TS - boff)) / TS - 1) * TS + boff;
            if (lane == 0) { S.cells[0] = make_int4(-1, k, 0, na); S.cells[1] = make_int4(-1, k, 0, nb); }
            ha = 0; hb = 1; ham = na; hbm = nb; S.ncell = 2;
            na += TS; nb += TS;
            const int ra = alen - x, rb = blen - y;
            n = slide_fwd(aw, abase + x, bw, bbase + y, ra < rb ? ra : rb);
            y += n; x += n;
            if (y == blen) { more = 0; if (bclip < k) bclip = k; }
            else if (x == alen) { more = 0; aclip = k; }
            while (x >= na) { if (lane == 0) S.cells[S.ncell] = make_int4(ha, k, 0, na); ha = S.ncell++; ham = na; na += TS; }
            while (y >= nb) { if (lane == 0) S.cells[S.ncell] = make_int4(hb, k, 0, nb); hb = S.ncell++; hbm = nb; nb += TS; }
        } else {
            na = ((x + (TS - aoff) - 1) / TS - 1) * TS + aoff;
            nb = ((y + (TS - boff) - 1) / TS - 1) * TS + boff;
            if (lane == 0) { S.cells[0] = make_int4(-1, k, 0, x); S.cells[1] = make_int4(-1, k, 0, y); }
            ha = 0; hb = 1; ham = x; hbm = y; S.ncell = 2;
            n = slide_rev(aw, abase + x, bw, bbase + y, x < y ? x : y);
            y -= n; x -= n;
            if (y == 0) { more = 0; if (bclip > k) bclip = k; }
            else if (x == 0) { more = 0; aclip = k; }
            while (x <= na) { if (lane == 0) S.cells[S.ncell] = make_int4(ha, k, 0, na); ha = S.ncell++; ham = na; na -= TS; }
            while (y <= nb) { if (lane == 0) S.cells[S.ncell] = make_int4(hb, k, 0, nb); hb = S.ncell++; hbm = nb; nb -= TS; }
        }
        const int c = (y << 1) + k;
        if (DIR > 0 ? (c > besta) : (c < besta)) { besta = trima = lasta = c; besty = trimy = y; trimha = ha; trimhb = hb; }
        V = c; T = PATH_INT; HA = ha; HB = hb; HM = ham; HN = hbm; NA = na; NB = nb;     // every lane holds a copy; only lane k&31 is live
    }

    for (;;) {
        // ---- boundary clipping ----
        if (more == 0) {
            bool inside;
            if (DIR > 0) inside = (besty < blen) && (besta - besty < alen);
            else inside = (besty > 0) && (besta - besty > 0);
            if (inside) more = 1;
#pragma unroll
            for (int side = 0; side < 2; side++) {
                const int kc = side == 0 ? aclip : bclip;
                bool hit;
                if (DIR > 0) hit = side == 0 ? (hgh >= aclip) : (low <= bclip);
                else         hit = side == 0 ? (low <= aclip) : (hgh >= bclip);
                if (!hit) continue;
                if (DIR > 0) { if (side == 0) hgh = kc - 1; else low = kc + 1; }
                else         { if (side == 0) low = kc + 1; else hgh = kc - 1; }
                const int sl = kc & 31;
                const uint64_t tc = __shfl_sync(FULL, T, sl);
                const int vc = __shfl_sync(FULL, V, sl), hac = __shfl_sync(FULL, HA, sl), hbc = __shfl_sync(FULL, HB, sl);
                const int mc = __popcll(tc & PATH_MASK61);
                if (morem <= mc) { morem = mc; morea = vc; morey = (morea - kc) / 2; mored = dif; moreha = hac; morehb = hbc; }
            }
            aclip = (DIR > 0) ? IMAX : -IMAX; bclip = (DIR > 0) ? -IMAX : IMAX;
        }
        // ---- prune lagging ends ----
        if (dif > 0 && hgh >= low) {
            const int k = low + ((lane - low) & 31);
            const bool keep = (k <= hgh) && (DIR > 0 ? (V >= besta - WAVE_LAG) : (V <= besta + WAVE_LAG));
            const unsigned km = __ballot_sync(FULL, keep);
            if (km == 0) hgh = low - 1;
            else {
                // bit r of `byrank` <-> diagonal hgh - r
                const unsigned byrank = rotr32(__brev(km), 31 - (hgh & 31));
                const int top = __ffs(byrank) - 1, bot = 31 - __clz(byrank);
                const int nh = hgh - top, nl = hgh - bot;
                hgh = nh; low = nl;
            }
            if (hgh >= low) cells_done += (unsigned)(hgh - low + 1);
        }
        if (!(more && (DIR > 0 ? (lasta >= besta - TRIM_MLAG) : (lasta <= besta + TRIM_MLAG)))) break;
        if (hgh < low) break;

        // ---- widen ----
        const int plow = low, phgh = hgh;
        if (low - 1 >= minp) low -= 1;
        if (hgh + 1 <= maxp) hgh += 1;
        if (hgh - low + 1 > 32) { S.overflow = 2; break; }
        dif += 1;
        const int k = low + ((lane - low) & 31);
        const bool act = k <= hgh;
        const int lm = (lane - 1) & 31, lp = (lane + 1) & 31;
        {   // new end diagonals inherit the trace cursors of their inner neighbour
            const int naR = __shfl_sync(FULL, NA, lp), nbR = __shfl_sync(FULL, NB, lp);
            const int naL = __shfl_sync(FULL, NA, lm), nbL = __shfl_sync(FULL, NB, lm);
            if (act && k < plow) { NA = naR; NB = nbR; }
            else if (act && k > phgh) { NA = naL; NB = nbL; }
        }
        const int vl = __shfl_sync(FULL, V, lm), vr = __shfl_sync(FULL, V, lp);
        const int vm = (k - 1 >= plow && k - 1 <= phgh) ? vl : NEG;
        const int vc = (k >= plow && k <= phgh) ? V : NEG;
        const int vp = (k + 1 >= plow && k + 1 <= phgh) ? vr : NEG;
        int c, dir;
        if (DIR > 0) {
            if (vc < vm) { if (vm < vp) { c = vp + 1; dir = 1; } else { c = vm + 1; dir = -1; } }
            else if (vc < vp) { c = vp + 1; dir = 1; }
            else { c = vc + 2; dir = 0; }
        } else {
            if (vc > vp) { if (vp > vm) { c = vm - 1; dir = -1; } else { c = vp - 1; dir = 1; } }
            else if (vc > vm) { c = vm - 1; dir = -1; }
            else { c = vc - 2; dir = 0; }
        }
        const int sl = (lane + dir) & 31;
        uint64_t b = __shfl_sync(FULL, T, sl) << 1;
        int ha = __shfl_sync(FULL, HA, sl), hb = __shfl_sync(FULL, HB, sl);
        int ham = __shfl_sync(FULL, HM, sl), hbm = __shfl_sync(FULL, HN, sl);
        int y = 0, m = 0;
        bool clipA = false, clipB = false;
        if (act) {
            y = (c - k) >> 1;
            int x = y + k, n;
            if (DIR > 0) {
                const int ra = alen - x, rb = blen - y;
                n = slide_fwd(aw, abase + x, bw, bbase + y, ra < rb ? ra : rb);
                y += n; x += n;
                if (y == blen) clipB = true; else if (x == alen) clipA = true;
            } else {
                n = slide_rev(aw, abase + x, bw, bbase + y, x < y ? x : y);
                y -= n; x -= n;
                if (y == 0) clipB = true; else if (x == 0) clipA = true;
            }
            b = push_matches(b, n);
            c = (y << 1) + k;
            m = __popcll(b & PATH_MASK61);
        } else c = (DIR > 0) ? -IMAX : IMAX;
        // ---- pebbles ----
        {
            const int x = y + k;
            for (;;) {
                const bool need = act && (DIR > 0 ? (x >= NA) : (x <= NA));
                if (!__any_sync(FULL, need)) break;
                const bool push = need && (DIR > 0 ? (ham < NA) : (ham > NA));
                const unsigned pm = __ballot_sync(FULL, push);
                if (push) {
                    const int id = S.ncell + __popc(pm & ((1u << lane) - 1u));
                    if (id < S.cells_cap) S.cells[id] = make_int4(ha, k, dif, NA);
                    ha = id; ham = NA;
                }
                S.ncell += __popc(pm);
                if (need) NA += DIR * TS;
            }
            for (;;) {
                const bool need = act && (DIR > 0 ? (y >= NB) : (y <= NB));
                if (!__any_sync(FULL, need)) break;
                const bool push = need && (DIR > 0 ? (hbm < NB) : (hbm > NB));
                const unsigned pm = __ballot_sync(FULL, push);
                if (push) {
                    const int id = S.ncell + __popc(pm & ((1u << lane) - 1u));
                    if (id < S.cells_cap) S.cells[id] = make_int4(hb, k, dif, NB);
                    hb = id; hbm = NB;
                }
                S.ncell += __popc(pm);
                if (need) NB += DIR * TS;
            }
            if (S.ncell >= S.cells_cap) { S.overflow = 1; break; }
        }
        // ---- best / lasta / trim (align.c:911-928 / 1620-1637).  In the reference's visiting order (forward
        //      hgh -> low, reverse low -> hgh) a diagonal is a record setter iff its point beats besta and every
        //      point visited before it.  The last setter is the wave's extreme point; lasta / trim come from the
        //      last setter that also passes the m / table tests, found by walking the few improving candidates
        //      from the best one down (setters are strictly monotone in visiting order). ----
#define BYRANK(mask) ((DIR > 0) ? rotr32(__brev(mask), 31 - (hgh & 31)) : rotr32((mask), low & 31))
#define RANK2LANE(r) ((((DIR > 0) ? (hgh - (r)) : (low + (r)))) & 31)
        {
            const bool beat = act && (DIR > 0 ? (c > besta) : (c < besta));
            if (__any_sync(FULL, beat)) {
                const int ext = (DIR > 0) ? __reduce_max_sync(FULL, beat ? c : -IMAX) : __reduce_min_sync(FULL, beat ? c : IMAX);
                {
                    const unsigned em = __ballot_sync(FULL, beat && c == ext);
                    const int l = RANK2LANE(__ffs(BYRANK(em)) - 1);          // first visited lane holding the extreme
                    besta = ext; besty = __shfl_sync(FULL, y, l);
                }
                unsigned gm = __ballot_sync(FULL, beat && m >= PATH_AVE);
                bool have_lasta = false;
                while (gm) {
                    const bool ing = (gm >> lane) & 1u;
                    const int gext = (DIR > 0) ? __reduce_max_sync(FULL, ing ? c : -IMAX) : __reduce_min_sync(FULL, ing ? c : IMAX);
                    const unsigned gem = __ballot_sync(FULL, ing && c == gext);
                    const int r = __ffs(BYRANK(gem)) - 1;                    // visiting rank of the candidate
                    const int l = RANK2LANE(r);
                    const unsigned geq = __ballot_sync(FULL, act && (DIR > 0 ? (c >= gext) : (c <= gext)));
                    if ((BYRANK(geq) & ((1u << r) - 1u)) == 0) {             // nothing visited earlier is as good: a setter
                        if (!have_lasta) { lasta = gext; have_lasta = true; }
                        int tok = 0;
                        if (lane == l) {
                            const int lo15 = (int)(b & TRIM_MASK), hi15 = (int)((b >> TRIM_LEN) & TRIM_MASK);
                            tok = (S.table_tab[lo15] >= 0) && (S.table_tab[hi15] + S.score_tab[lo15] >= 0);
                        }
                        if (__shfl_sync(FULL, tok, l)) {
                            trima = gext; trimy = __shfl_sync(FULL, y, l); trimd = dif;
                            trimha = __shfl_sync(FULL, ha, l); trimhb = __shfl_sync(FULL, hb, l);
                            break;
                        }
                    }
                    gm &= ~gem;
                }
            }
        }
        // ---- clip bookkeeping in visiting order (align.c:836-851 / 1545-1560) ----
        if (__any_sync(FULL, clipA || clipB)) {
            const unsigned am = __ballot_sync(FULL, clipA), bm = __ballot_sync(FULL, clipB);
            more = 0;
            if (am) {       // aclip = last visited clipped diagonal
                const int r = 31 - __clz(BYRANK(am));
                aclip = (DIR > 0) ? (hgh - r) : (low + r);
            }
            if (bm) {       // bclip = extreme = first visited
                const int r = __ffs(BYRANK(bm)) - 1;
                const int kb = (DIR > 0) ? (hgh - r) : (low + r);
                if (DIR > 0) { if (bclip < kb) bclip = kb; } else { if (bclip > kb) bclip = kb; }
            }
        }
#undef BYRANK
#undef RANK2LANE
        if (act) { V = c; T = b; HA = ha; HB = hb; HM = ham; HN = hbm; }
    }
    S.ncells_done += cells_done;
    if (morem >= 0) { ex = morea - morey; ey = morey; ediff = mored; eha = moreha; ehb = morehb; }
    else { ex = trima - trimy; ey = trimy; ediff = trimd; eha = trimha; ehb = trimhb; }
}

// Local_Alignment (align.c:1904-2097) for daligner's call shape.  Warp-collective.
template <bool FAST>
__device__ void local_alignment(WarpLA &S, const AlignArgs &P, uint16_t *tbuf, int ar, int br, int alen, int blen,
                                int diag, int anti, bool selfie, Path &ap, Path &bp) {
    const int lane = threadIdx.x & 31;
    const int64_t abase = P.aboff[ar], bbase = P.bboff[br];
    const int minp = (selfie && diag >= 0) ? 1 : -IMAX;      // align.c:1951-1968
    const int maxp = (selfie && diag <= 0) ? -1 : IMAX;
    const int aoff = 0, boff = P.comp ? (blen % P.TS) : 0;   // align.c:1970-1984
    uint16_t *atr = tbuf + 2 * P.maxtp, *btr = tbuf + 6 * P.maxtp;
    int ex, ey, ed, ha, hb;
    int flenA = 0, flenB = 0;

    if (FAST) wave_pass_fast<+1>(S, P.awords, abase, alen, P.bwords, bbase, blen, diag, anti, minp, maxp, aoff, boff, ex, ey, ed, ha, hb);
    else wave_pass<+1>(S, P.awords, abase, alen, P.bwords, bbase, blen, diag, anti, minp, maxp, aoff, boff, ex, ey, ed, ha, hb);
    if (S.overflow) return;
    __syncwarp();
    if (lane == 0) {
        flenA = emit_chain_fwd(S.cells, ha, true, anti, ex, ey, ed, atr);
        flenB = emit_chain_fwd(S.cells, hb, false, anti, ex, ey, ed, btr);
    }
    flenA = __shfl_sync(FULL, flenA, 0); flenB = __shfl_sync(FULL, flenB, 0);
    ap.aepos = ex; ap.bepos = ey; ap.diffs = ed;
    __syncwarp();

    if (FAST) wave_pass_fast<-1>(S, P.awords, abase, alen, P.bwords, bbase, blen, diag, anti, minp, maxp, aoff, boff, ex, ey, ed, ha, hb);
    else wave_pass<-1>(S, P.awords, abase, alen, P.bwords, bbase, blen, diag, anti, minp, maxp, aoff, boff, ex, ey, ed, ha, hb);
    if (S.overflow) return;
    __syncwarp();
    int preA = 0, preB = 0;
    if (lane == 0) {
        preA = emit_chain_rev(S.cells, ha, true, ex, ey, ed, P.TS, aoff, flenA, atr);
        preB = emit_chain_rev(S.cells, hb, false, ex, ey, ed, P.TS, boff, flenB, btr);
    }
    preA = __shfl_sync(FULL, preA, 0); preB = __shfl_sync(FULL, preB, 0);
    __syncwarp();
    ap.abpos = ex; ap.bbpos = ey; ap.diffs += ed;
    ap.tlen = flenA + preA; ap.trace = atr - preA;
    bp.tlen = flenB + preB; bp.trace = btr - preB;
    bp.diffs = ap.diffs;
    if (P.comp) {                                            // align.c:2033-2055
        bp.abpos = blen - ap.bepos; bp.bbpos = alen - ap.aepos;
        bp.aepos = blen - ap.bbpos; bp.bepos = alen - ap.abpos;
        uint16_t *tr = bp.trace;
        const int npair = bp.tlen >> 1;
        for (int i = lane; i < npair / 2; i += 32) {         // reverse the (d, delta) pairs as pairs
            const int j = npair - 1 - i;
            const uint16_t d0 = tr[2 * i], d1 = tr[2 * i + 1];
            tr[2 * i] = tr[2 * j]; tr[2 * i + 1] = tr[2 * j + 1];
            tr[2 * j] = d0; tr[2 * j + 1] = d1;
        }
        __syncwarp();
    } else {                                                 // align.c:2056-2062
        bp.aepos = ap.bepos; bp.bepos = ap.aepos; bp.abpos = ap.bbpos; bp.bbpos = ap.abpos;
    }
}

__device__ void emit_record(const AlignArgs &P, const Path &p, int aread, int bread, uint64_t seq) {
    const int lane = threadIdx.x & 31;
    unsigned long long ri = 0, to = 0;
    if (lane == 0) {
        ri = atomicAdd(P.rec_count, 1ull);
        to = atomicAdd(P.trace_count, (unsigned long long)p.tlen);
    }
    ri = __shfl_sync(FULL, ri, 0); to = __shfl_sync(FULL, to, 0);
    if (ri >= P.rec_cap || to + p.tlen > P.trace_cap) { if (lane == 0) atomicOr(&P.stats[3], 1ull); return; }
    if (lane == 0) {
        damar_overlap r;
        r.tlen = p.tlen; r.diffs = p.diffs; r.abpos = p.abpos; r.bbpos = p.bbpos; r.aepos = p.aepos; r.bepos = p.bepos;
        r.flags = (uint32_t)P.comp; r.aread = aread; r.bread = bread; r.pad = 0;
        P.recs[ri] = r; P.rec_toff[ri] = to; P.rec_seq[ri] = seq;
    }
    if (P.tbytes == 1) {                                     // Compress_TraceTo8 (align.c:3375-3396)
        uint8_t *t8 = reinterpret_cast<uint8_t *>(P.trace);
        bool big = false;
        for (int i = lane; i < p.tlen; i += 32) { const uint16_t v = p.trace[i]; big |= v > 255; t8[to + i] = (uint8_t)v; }
        if (__any_sync(FULL, big) && lane == 0) atomicOr(&P.stats[3], 8ull);
    } else {
        uint16_t *t16 = reinterpret_cast<uint16_t *>(P.trace);
        for (int i = lane; i < p.tlen; i += 32) t16[to + i] = p.trace[i];
    }
}

template <bool FAST>
__global__ void __launch_bounds__(AL_WARPS * 32, FAST ? 4 : 3) align_pairs(AlignArgs P) {
    __shared__ int la_low[AL_WARPS][MAX_LA_PER_PAIR], la_hgh[AL_WARPS][MAX_LA_PER_PAIR], la_ae[AL_WARPS][MAX_LA_PER_PAIR];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * AL_WARPS + wib;
    const uint64_t pmask = (1ull << P.pbits) - 1ull;

    WarpLA S;
    {
        const size_t r = (size_t)P.ring;
        for (int i = 0; i < 2; i++) {
            S.V[i] = P.V + ((size_t)gw * 2 + i) * r; S.HA[i] = P.HA + ((size_t)gw * 2 + i) * r;
            S.HB[i] = P.HB + ((size_t)gw * 2 + i) * r; S.HM[i] = P.HM + ((size_t)gw * 2 + i) * r;
            S.HN[i] = P.HN + ((size_t)gw * 2 + i) * r; S.T[i] = P.T + ((size_t)gw * 2 + i) * r;
        }
        S.NA = P.NA + (size_t)gw * r; S.NB = P.NB + (size_t)gw * r;
        S.cells = P.cells + (size_t)gw * P.cells_cap; S.cells_cap = P.cells_cap; S.ncell = 0;
        S.rmask = P.ring - 1; S.score_tab = P.score_tab; S.table_tab = P.table_tab;
        S.TS = P.TS; S.PATH_AVE = P.ave_path; S.overflow = 0; S.ncells_done = 0;
    }
    uint16_t *tbuf = P.tbuf + (size_t)gw * 8 * P.maxtp;
    int32_t *score = P.bscore + (size_t)gw * 2 * P.nbuck - P.dmin + 2;   // +2 margin for d-1
    int32_t *lastp = score + P.nbuck;
    unsigned long long n_la = 0, n_multi = 0, n_pairs = 0;

    for (;;) {
        unsigned wslot = 0;
        if (lane == 0) wslot = atomicAdd(P.next_work, 1u);
        wslot = __shfl_sync(FULL, wslot, 0);
        if (wslot >= P.nwork) break;
        const int64_t sidx = P.work[wslot];
        const unsigned wi = P.work_wi ? P.work_wi[wslot] : wslot;
        const unsigned long long la0 = n_la, multi0 = n_multi, cells0 = S.ncells_done;
        const uint64_t pair = P.keys[sidx] >> P.pbits;
        const int ar = (int)(pair & ((1ull << P.abits) - 1ull)), br = (int)(pair >> P.abits);
        const int alen = P.arlen[ar], blen = P.brlen[br];
        const bool selfie = P.self && !P.comp && ar == br;    // aseq == bseq pointer test, align.c:1949
        const bool doB = P.symmetric && (ar != br || !P.self || !P.comp);   // filter.c:2301
        n_pairs++;

        // extent of the pair
        int64_t pend = sidx;
        for (;;) {
            const int64_t j = pend + lane;
            const bool same = (j < P.nhits) && ((P.keys[j] >> P.pbits) == pair);
            const unsigned m = __ballot_sync(FULL, same);
            if (m == FULL) { pend += 32; continue; }
            pend += __ffs(~m) - 1;
            break;
        }
        int nla = 0, novl = 0;
        int amark2 = 0;
        int64_t nidx = sidx;
        while (nidx < pend) {                                  // A-panels, filter.c:2251-2266
            const int amark = amark2 + PANEL_SIZE; amark2 = amark - PANEL_OVERLAP;
            const int64_t lidx = nidx;
            // panel = hits up to and including the first with apos > amark; h2 = first with apos > amark2
            int64_t pe = pend, h2 = pend;
            {
                bool gotpe = false, goth2 = false;
                for (int64_t j0 = lidx; j0 < pend && !(gotpe && goth2); j0 += 32) {
                    const int64_t j = j0 + lane;
                    const int apos = (j < pend) ? (int)(P.keys[j] & pmask) : IMAX;
                    const unsigned over = __ballot_sync(FULL, (j < pend) && apos > amark);
                    const unsigned over2 = __ballot_sync(FULL, (j < pend) && apos > amark2);
                    if (!goth2 && over2) { h2 = j0 + __ffs(over2) - 1; goth2 = true; }
                    if (!gotpe && over) { pe = j0 + __ffs(over) - 1 + 1; gotpe = true; }
                }
                if (h2 > pe) h2 = pe;
            }
            nidx = h2;
            if (pe - lidx < P.minhit) continue;

            // pass 1: covered bases per bucket (filter.c:2268-2277)
            for (int64_t j0 = lidx; j0 < pe; j0 += 32) {
                const int64_t j = j0 + lane;
                const bool v = j < pe;
                int apos = 0, d = IMAX - lane;                 // distinct dummies so inactive lanes never match
                if (v) { apos = (int)(P.keys[j] & pmask); d = (apos - (int)P.bpos[j]) >> P.W; }
                const unsigned peers = __match_any_sync(FULL, d);
                const unsigned lower = peers & ((1u << lane) - 1u);
                int prev = 0;
                const int src = lower ? (31 - __clz(lower)) : lane;
                const int prev_peer = __shfl_sync(FULL, apos, src);
                if (v) prev = lower ? prev_peer : lastp[d];
                __syncwarp();
                if (v) {
                    const int add = (apos - prev >= P.K) ? P.K : (apos - prev);
                    atomicAdd(&score[d], add);
                    if ((peers >> lane) == 1u) lastp[d] = apos;   // highest lane of the group
                }
                __syncwarp();
            }
            // pass 2: threshold + alignment (filter.c:2283-2405)
            for (int64_t j0 = lidx; j0 < pe; j0 += 32) {
                const int64_t j = j0 + lane;
                const bool v = j < pe;
                int apos = 0, bpos = 0, d = 0;
                bool cand = false;
                if (v) {
                    apos = (int)(P.keys[j] & pmask); bpos = (int)P.bpos[j]; d = (apos - bpos) >> P.W;
                    const int s = score[d];
                    cand = (s + score[d + 1] >= P.H) || (s + score[d - 1] >= P.H);
                }
                unsigned cm = __ballot_sync(FULL, cand);
                while (cm) {
                    const int l = __ffs(cm) - 1; cm &= cm - 1;
                    const int capos = __shfl_sync(FULL, apos, l), cbpos = __shfl_sync(FULL, bpos, l), cd = __shfl_sync(FULL, d, l);
                    // apos > lasta[d] (filter.c:2289) against the pair's alignment list
                    bool blocked = false;
                    for (int q = lane; q < nla; q += 32)
                        if (la_low[wib][q] <= cd && cd <= la_hgh[wib][q] && capos <= la_ae[wib][q]) blocked = true;
                    if (__any_sync(FULL, blocked)) continue;
                    Path ap, bp;
                    n_la++;
                    local_alignment<FAST>(S, P, tbuf, ar, br, alen, blen, capos - cbpos, capos + cbpos, selfie, ap, bp);
                    if (S.overflow) break;
                    // Diagonal_Span (filter.c:2079-2110) on the uncompressed A trace
                    int lo, hi;
                    {
                        int dd = ap.abpos - ap.bbpos; lo = hi = dd;
                        dd = ap.aepos - ap.bepos; if (dd < lo) lo = dd; else if (dd > hi) hi = dd;
                        // running diagonal at interior trace points: prefix sums of (TS - delta_b)
                        const int nint = (ap.tlen - 2) / 2;    // pairs 0 .. nint-1
                        int base = (ap.abpos / P.TS) * P.TS - ap.bbpos;
                        for (int q0 = 0; q0 < nint; q0 += 32) {
                            const int q = q0 + lane;
                            int val = (q < nint) ? (P.TS - (int)ap.trace[2 * q + 1]) : 0;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, val, o); if (lane >= o) val += t; }
                            const int dd2 = base + val;
                            int mn = (q < nint) ? dd2 : IMAX, mx = (q < nint) ? dd2 : -IMAX;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(FULL, mn, o)); mx = max(mx, __shfl_xor_sync(FULL, mx, o)); }
                            if (mn < lo) lo = mn;
                            if (mx > hi) hi = mx;
                            base += __shfl_sync(FULL, val, 31);
                        }
                        lo = (lo >> P.W) - 1; hi = (hi >> P.W) + 1;
                        if (cd < lo) lo = cd; else if (cd > hi) hi = cd;
                    }
                    if (nla < MAX_LA_PER_PAIR) {
                        if (lane == 0) { la_low[wib][nla] = lo; la_hgh[wib][nla] = hi; la_ae[wib][nla] = ap.aepos; }
                        nla++;
                    } else if (lane == 0) atomicOr(&P.stats[3], 2ull);
                    __syncwarp();
                    if ((ap.aepos - ap.abpos) + (ap.bepos - ap.bbpos) >= P.minover) {   // filter.c:2335
                        const uint64_t seq = ((uint64_t)wi << 8) | (uint64_t)(novl & 127);
                        emit_record(P, ap, ar + P.afirst, br + P.bfirst, seq);
                        if (doB) emit_record(P, bp, br + P.bfirst, ar + P.afirst, seq | 128ull);
                        novl++;
                    }
                    __syncwarp();
                }
                if (S.overflow) break;
            }
            // pass 3: clear the touched buckets (filter.c:2407-2411)
            for (int64_t j0 = lidx; j0 < pe; j0 += 32) {
                const int64_t j = j0 + lane;
                if (j < pe) {
                    const int apos = (int)(P.keys[j] & pmask);
                    const int d = (apos - (int)P.bpos[j]) >> P.W;
                    score[d] = 0; lastp[d] = 0;
                }
            }
            __syncwarp();
            if (S.overflow) break;
        }
        if (novl > 1) n_multi++;
        if (FAST && S.overflow == 2) {
            // wave front outgrew the warp: hand the whole pair to the strided pass; whatever this pass
            // emitted for the pair is dropped on the host (pair_redo)
            if (lane == 0) {
                const unsigned q = atomicAdd(P.redo_count, 1u);
                P.redo_sidx[q] = (uint32_t)sidx; P.redo_wi[q] = wi; P.pair_redo[wi] = 1;
            }
            n_la = la0; n_multi = multi0; S.ncells_done = cells0; n_pairs--;
            S.overflow = 0;
        }
        if (S.overflow) { if (lane == 0) atomicOr(&P.stats[3], 4ull); S.overflow = 0; }
    }
    if (lane == 0) {
        atomicAdd(&P.stats[0], n_la); atomicAdd(&P.stats[1], n_multi);
        atomicAdd(&P.stats[2], S.ncells_done); atomicAdd(&P.stats[4], n_pairs);
    }
}

// New_Align_Spec + set_table (align.c:226-318): floating point only here, evaluated exactly as the
// reference does (float sum of freq[0]+freq[3], double afterwards).
double Bias_Factor[10] = {.690, .690, .690, .690, .780, .850, .900, .933, .966, 1.000};
void build_trim_tables(const damar_options &opt, int &ave_path, std::vector<int16_t> &score, std::vector<int16_t> &table) {
    double match = opt.freq[0] + opt.freq[3];
    if (match > .5) match = 1. - match;
    int bias = (int)((match + .025) * 20. - 1.);
    if (match < .2) {
        fprintf(stderr, "Warning: Base bias worse than 80/20%% ! (New_Align_Spec)\n         Capping bias at this ratio.\n");
        bias = 3;
    }
    ave_path = (int)(PATH_LEN * (1. - Bias_Factor[bias] * (1. - opt.ave_corr)));
    const int mscore = (int)(FRACTION * Bias_Factor[bias] * (1. - opt.ave_corr));
    const int dscore = FRACTION - mscore;
    score.resize(TRIM_MASK + 1); table.resize(TRIM_MASK + 1);
    for (int x = 0; x <= TRIM_MASK; x++) {
        int s = 0, mx = 0;
        for (int bit = TRIM_LEN - 1; bit >= 0; bit--) { if (s > mx) mx = s; s += ((x >> bit) & 1) ? mscore : -dscore; }
        table[x] = (int16_t)(s - mx); score[x] = (int16_t)s;
    }
}

}  // namespace

void report_and_align(const DeviceBlock *ab, const DeviceBlock *bb, int comp, int self, const SeedHits &hits, HostOverlaps &out) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const FilterParams &fp = c.fp;
    const int64_t n = hits.n;
    const int minhit = (fp.hitmin - 1) / fp.kmer + 1;

    // ---- work list ----
    HostProf *hp1 = new HostProf("ra.worklist");
    Timer t_pairs(st);
    DevBuf<int64_t> ends; ends.reserve(fp.nthreads_pow2);
    DevBuf<uint32_t> flag; flag.reserve((size_t)n);
    DevBuf<uint64_t> off; off.reserve((size_t)n);
    DevBuf<uint64_t> total; total.reserve(1);
    slice_ends<<<1, 1024, 0, st>>>(hits.keys, n, fp.nthreads_pow2 > 1024 ? 1024 : fp.nthreads_pow2, fp.nshift, hits.pbits + hits.abits, ends.p);
    KERNEL_CHECK();
    const unsigned hb = (unsigned)((n + 255) / 256);
    flag_work_pairs<<<hb, 256, 0, st>>>(hits.keys, n, hits.pbits, minhit, ends.p, fp.nthreads_pow2 > 1024 ? 1024 : fp.nthreads_pow2, flag.p);
    KERNEL_CHECK();
    exclusive_scan_u32_to_u64(flag.p, off.p, n, total.p, st);
    uint64_t nwork = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nwork, total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    DevBuf<uint32_t> work; work.reserve((size_t)nwork + 1);
    collect_work_pairs<<<hb, 256, 0, st>>>(flag.p, off.p, n, work.p);
    KERNEL_CHECK();
    count_launch(3);
    c.tm.pairs_ms = t_pairs.stop();
    if (nwork == 0) { delete hp1; return; }

    delete hp1;
    // ---- alignment ----
    HostProf *hp2 = new HostProf("ra.align_setup");
    Timer t_align(st);
    int ave_path; std::vector<int16_t> hs, ht;
    build_trim_tables(c.opt, ave_path, hs, ht);
    DevBuf<int16_t> dscore, dtable; dscore.reserve(hs.size()); dtable.reserve(ht.size());
    CUDA_CHECK(cudaMemcpyAsync(dscore.p, hs.data(), hs.size() * 2, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(dtable.p, ht.data(), ht.size() * 2, cudaMemcpyHostToDevice, st));

    const int TS = c.opt.trace_spacing;
    const int blocks = NUM_SMS_B200 * 4;
    const int nwarps = blocks * AL_WARPS;
    const int ring = 2048;
    const int maxlen = ab->maxlen > bb->maxlen ? ab->maxlen : bb->maxlen;
    const int maxtp = 2 * (maxlen / TS + 2) + 2;
    int cells_cap = 4 * (2 * (maxlen / TS + 2)) * 24 + 4096;
    const int dmin = (-bb->maxlen) >> fp.binshift, dmax = ab->maxlen >> fp.binshift;
    const int nbuck = dmax - dmin + 1 + 4;

    DevBuf<int32_t> V, HA, HB, HM, HN, NA, NB, bscore; DevBuf<uint64_t> T; DevBuf<int4> cells; DevBuf<uint16_t> tbuf;
    V.reserve((size_t)nwarps * 2 * ring); HA.reserve((size_t)nwarps * 2 * ring); HB.reserve((size_t)nwarps * 2 * ring);
    HM.reserve((size_t)nwarps * 2 * ring); HN.reserve((size_t)nwarps * 2 * ring); T.reserve((size_t)nwarps * 2 * ring);
    NA.reserve((size_t)nwarps * ring); NB.reserve((size_t)nwarps * ring);
    cells.reserve((size_t)nwarps * cells_cap); tbuf.reserve((size_t)nwarps * 8 * maxtp);
    bscore.reserve((size_t)nwarps * 2 * nbuck);
    CUDA_CHECK(cudaMemsetAsync(bscore.p, 0, sizeof(int32_t) * (size_t)nwarps * 2 * nbuck, st));

    DevBuf<unsigned long long> counters; counters.reserve(8);
    DevBuf<unsigned int> next; next.reserve(4);
    DevBuf<uint32_t> redo_sidx, redo_wi; DevBuf<uint8_t> pair_redo;
    redo_sidx.reserve((size_t)nwork + 1); redo_wi.reserve((size_t)nwork + 1); pair_redo.reserve((size_t)nwork + 1);
    unsigned int nredo = 0; unsigned long long nrec_fast = 0;
    const int tbytes = (TS <= 125) ? 1 : 2;                  // TRACE_XOVR, align.h:58
    unsigned long long rec_cap = 4 * nwork + 1024;
    unsigned long long trace_cap = (unsigned long long)rec_cap * (unsigned long long)(2 * (maxlen / TS + 2)) / 8 + (1ull << 20);
    DevBuf<damar_overlap> recs; DevBuf<uint64_t> rtoff, rseq; DevBuf<uint8_t> trace;
    unsigned long long hc[8];
    delete hp2;
    HostProf *hp3 = new HostProf("ra.align_run");
    for (int attempt = 0;; attempt++) {
        recs.reserve((size_t)rec_cap); rtoff.reserve((size_t)rec_cap); rseq.reserve((size_t)rec_cap);
        trace.reserve((size_t)trace_cap * tbytes);
        CUDA_CHECK(cudaMemsetAsync(counters.p, 0, sizeof(unsigned long long) * 8, st));
        CUDA_CHECK(cudaMemsetAsync(next.p, 0, sizeof(unsigned int) * 4, st));
        CUDA_CHECK(cudaMemsetAsync(pair_redo.p, 0, (size_t)nwork + 1, st));
        AlignArgs P{};
        P.keys = hits.keys; P.bpos = hits.bpos; P.nhits = n; P.pbits = hits.pbits; P.abits = hits.abits;
        P.work = work.p; P.nwork = (uint32_t)nwork; P.next_work = next.p; P.work_wi = nullptr;
        P.redo_sidx = redo_sidx.p; P.redo_wi = redo_wi.p; P.redo_count = next.p + 1; P.pair_redo = pair_redo.p;
        P.awords = ab->packed[0]; P.bwords = bb->packed[(comp && !bb->as_is) ? 1 : 0];
        P.aboff = ab->boff; P.bboff = bb->boff; P.arlen = ab->rlen; P.brlen = bb->rlen;
        P.afirst = ab->ufirst; P.bfirst = bb->ufirst;
        P.K = fp.kmer; P.W = fp.binshift; P.H = fp.hitmin; P.minhit = minhit; P.TS = TS; P.minover = c.opt.minover;
        P.comp = comp ? 1 : 0; P.self = self; P.symmetric = c.opt.symmetric;
        P.ave_path = ave_path; P.score_tab = dscore.p; P.table_tab = dtable.p;
        P.ring = ring; P.V = V.p; P.HA = HA.p; P.HB = HB.p; P.HM = HM.p; P.HN = HN.p; P.NA = NA.p; P.NB = NB.p; P.T = T.p;
        P.cells = cells.p; P.cells_cap = cells_cap; P.tbuf = tbuf.p; P.maxtp = maxtp;
        P.bscore = bscore.p; P.nbuck = nbuck; P.dmin = dmin;
        P.recs = recs.p; P.rec_toff = rtoff.p; P.rec_seq = rseq.p; P.rec_cap = rec_cap; P.rec_count = counters.p + 5;
        P.trace = trace.p; P.tbytes = tbytes; P.trace_cap = trace_cap; P.trace_count = counters.p + 6;
        P.stats = counters.p;
        Timer t_k(st);
        align_pairs<true><<<blocks, AL_WARPS * 32, 0, st>>>(P);
        KERNEL_CHECK(); count_launch();
        c.tm.align_kernel_ms += t_k.stop(); c.tm.align_launches += 1;
        CUDA_CHECK(cudaMemcpyAsync(&nredo, next.p + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(&nrec_fast, counters.p + 5, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (nredo > 0) {                               // strided scratch-memory pass for the wide pairs
            CUDA_CHECK(cudaMemsetAsync(next.p, 0, sizeof(unsigned int), st));
            P.work = redo_sidx.p; P.work_wi = redo_wi.p; P.nwork = nredo;
            Timer t_r(st);
            align_pairs<false><<<blocks, AL_WARPS * 32, 0, st>>>(P);
            KERNEL_CHECK(); count_launch();
            c.tm.redo_kernel_ms += t_r.stop();
        }
        CUDA_CHECK(cudaMemcpyAsync(hc, counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if ((hc[3] & 1ull) && attempt < 4) {            // output buffers too small: grow and rerun
            if (hc[5] > rec_cap) rec_cap = hc[5] + 1024;
            if (hc[6] > trace_cap) trace_cap = hc[6] + 1024;
            CUDA_CHECK(cudaMemsetAsync(bscore.p, 0, sizeof(int32_t) * (size_t)nwarps * 2 * nbuck, st));
            continue;
        }
        break;
    }
    c.tm.align_ms = t_align.stop();
    delete hp3;
    if (hc[3] & 6ull)
        fatal("alignment scratch overflow (flags %llu): wave wider than %d diagonals, pebble arena or per-pair alignment list exhausted",
              hc[3], ring);
    if (hc[3] & 1ull) fatal("alignment output buffers overflowed repeatedly");
    if (hc[3] & 8ull) fatal("Compression of trace to bytes fails, value too big");   // align.c:3385-3388

    // ---- download into pinned staging; records stay in device emission order, `seq` carries the
    //      order the reference's thread buffers would have had ----
    HostProf hp4("ra.download");
    Timer t_dl(st);
    static PinnedBuf<damar_overlap> h_recs; static PinnedBuf<uint64_t> h_toff, h_seq; static PinnedBuf<uint8_t> h_trace;
    const uint64_t nrec = hc[5], ntr = hc[6];
    h_recs.reserve(nrec + 1); h_toff.reserve(nrec + 1); h_seq.reserve(nrec + 1); h_trace.reserve(ntr * tbytes + 16);
    if (nrec) {
        CUDA_CHECK(cudaMemcpyAsync(h_recs.p, recs.p, sizeof(damar_overlap) * nrec, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(h_toff.p, rtoff.p, sizeof(uint64_t) * nrec, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(h_seq.p, rseq.p, sizeof(uint64_t) * nrec, cudaMemcpyDeviceToHost, st));
    }
    if (ntr) CUDA_CHECK(cudaMemcpyAsync(h_trace.p, trace.p, (size_t)ntr * tbytes, cudaMemcpyDeviceToHost, st));
    std::vector<uint8_t> h_redo;
    if (nredo > 0) { h_redo.resize(nwork); CUDA_CHECK(cudaMemcpyAsync(h_redo.data(), pair_redo.p, nwork, cudaMemcpyDeviceToHost, st)); }
    CUDA_CHECK(cudaStreamSynchronize(st));
    uint64_t nkeep = nrec;
    if (nredo > 0) {                                   // drop what the fast pass emitted for pairs it gave up on
        nkeep = 0;
        for (uint64_t i = 0; i < nrec; i++) {
            if (i < nrec_fast && h_redo[h_seq.p[i] >> 8]) continue;
            if (nkeep != i) { h_recs.p[nkeep] = h_recs.p[i]; h_toff.p[nkeep] = h_toff.p[i]; h_seq.p[nkeep] = h_seq.p[i]; }
            nkeep++;
        }
    }
    out.ovl = h_recs.p; out.toff = h_toff.p; out.seq = h_seq.p; out.trace = h_trace.p;
    out.novl = (int64_t)nkeep; out.ntrace = (int64_t)ntr; out.tbytes = tbytes; out.redo = (int64_t)nredo;
    out.nfilt = (int64_t)hc[0]; out.multi = (int64_t)hc[1]; out.cells = (int64_t)hc[2]; out.pairs = (int64_t)hc[4];
    c.tm.download_ms = t_dl.stop();
}

}  // namespace damar
