// pair_align.cu -- stages 3 and 4: per read pair, diagonal-band hit scoring with the -h
// threshold, then O(ND) forward/reverse wave extension with trace-point pebbles.
//
// Replaces report_thread (reference dalign/filter.c:2128-2511, incl. Diagonal_Span :2079-2110)
// and Local_Alignment / forward_wave / reverse_wave (dalign/align.c:1904-2097, 409-1122,
// 1126-1898).  Two implementations live here:
//   * the task pipeline (second half of the file, the default): stage 3 and stage 4 are separate
//     kernels; every wave pass of every local alignment is one task run by ONE THREAD
//     (wave_core.cuh), so the reference's sequential tie-breaks need no reformulation and no lane
//     idles on a narrow front;
//   * the strided warp kernel (first half): one warp owns one (bread, aread) pair, the wave front is
//     strided over the lanes in global scratch.  It only sees the pairs whose front outgrows the
//     task kernel's 64-diagonal ring.
// Common to both: A-panels of 50 kbp (10 kbp overlap); pass 1 accumulates score[diag>>w] with the
// lanes taking 32 hits at a time (same-bucket conflicts resolved by lane order); pass 2 evaluates the
// threshold for 32 hits at once and consumes candidates in hit order unless an earlier alignment of
// the pair already covers them (the `lasta` rule, kept as a short interval list per pair).
#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "path_fuse_core.cuh"
#include "wave_core.cuh"

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
// Emission rank of a record = pair << (SEQ_LA_BITS + 1) | B side << SEQ_LA_BITS | alignment number of the pair.  The
// reference keeps a pair's alignments in realloc'ed arrays (filter.c:2318-2380); here the per-pair interval list grows
// in a device arena and only the rank field bounds the count (an error return, never a truncation).
constexpr int SEQ_LA_BITS = 12;
constexpr int MAX_LA_PER_PAIR = (1 << SEQ_LA_BITS) - 1;
constexpr uint64_t SEQ_B = 1ull << SEQ_LA_BITS;
constexpr int SEQ_PAIR_SHIFT = SEQ_LA_BITS + 1;
constexpr int LA_CAP0 = 4;             // interval slots a pair starts with
constexpr int IMAX = 0x7fffffff;

constexpr int AL_WARPS = 4;            // warps per CTA of the alignment kernel
constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// pair segmentation
// ---------------------------------------------------------------------------------------------

// slice ends of the reference's NTHREADS report threads (filter.c:2804-2816); ends[T-1] = n
// A comparison cut into chunks of B reads (api.cu) presents hits [base, base + n) of `total`: a slice end outside the chunk
// is -1 (no pair ends there); one at the chunk's first hit stays there, the previous chunk ended a B read.
__global__ void slice_ends(const uint64_t *__restrict__ keys, int64_t n, int nthreads, int nshift, int rshift,
                           int64_t *__restrict__ ends, int64_t base, int64_t total) {
    int t = threadIdx.x + 1;
    if (t > nthreads) return;
    if (t == nthreads) { ends[t - 1] = total - base; return; }
    int64_t p = ((total * t) >> nshift) - base;
    if (p < 0 || p > n) { ends[t - 1] = -1; return; }
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
    // one hit = key (bread | aread | apos) + bpos, or -- bpos == nullptr -- one word (bread | aread | apos | bpos) with
    // kshift = number of bpos bits (seed_merge.cu)
    const uint64_t *keys; const uint32_t *bpos; int64_t nhits; int pbits, abits, kshift; uint32_t qmask;
    const uint32_t *work; uint32_t nwork; unsigned int *next_work;
    const uint32_t *work_wi;           // optional: original work index of work[i] (second, strided pass)
    uint32_t *redo_sidx, *redo_wi; unsigned int *redo_count; uint8_t *pair_redo;   // fast pass: pairs handed to the strided pass
    // blocks
    const uint32_t *awords, *bwords; const uint32_t *aboff, *bboff; const int32_t *arlen, *brlen;
    int afirst, bfirst;
    // parameters
    int K, W, H, minhit, TS, minover, comp, self, symmetric;
    int reach;                         // Align_Spec.reach: 1 for daligner (daligner.c:951), 0 for datander (datander.c:246)
    int tandem;                        // datander's report_thread (scrub/tandem.c:895): panels over apos, every read, A records only
    int ave_path, dscore; const int16_t *score_tab, *table_tab;
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
    int32_t *la_scratch;               // strided kernel: [nwarps][3][MAX_LA_PER_PAIR + 1] interval lists
};

__device__ __forceinline__ uint64_t hit_pair(const AlignArgs &P, int64_t j) { return P.keys[j] >> (P.pbits + P.kshift); }
__device__ __forceinline__ int hit_apos(const AlignArgs &P, int64_t j) { return (int)((P.keys[j] >> P.kshift) & ((1ull << P.pbits) - 1ull)); }
__device__ __forceinline__ int hit_bpos(const AlignArgs &P, int64_t j) { return P.bpos ? (int)P.bpos[j] : (int)((uint32_t)P.keys[j] & P.qmask); }

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
    int TS, PATH_AVE, reach;
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

    if (morem >= 0 && S.reach) { ex = morea - morey; ey = morey; ediff = mored; eha = moreha; ehb = morehb; }   // REACH (align.c:1009)
    else { ex = trima - trimy; ey = trimy; ediff = trimd; eha = trimha; ehb = trimhb; }
}

// Local_Alignment (align.c:1904-2097) for daligner's call shape.  Warp-collective.
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

    wave_pass<+1>(S, P.awords, abase, alen, P.bwords, bbase, blen, diag, anti, minp, maxp, aoff, boff, ex, ey, ed, ha, hb);
    if (S.overflow) return;
    __syncwarp();
    if (lane == 0) {
        flenA = wc::emit_chain_fwd(reinterpret_cast<const wc::Cell *>(S.cells), ha, true, anti, ex, ey, ed, atr);
        flenB = wc::emit_chain_fwd(reinterpret_cast<const wc::Cell *>(S.cells), hb, false, anti, ex, ey, ed, btr);
    }
    flenA = __shfl_sync(FULL, flenA, 0); flenB = __shfl_sync(FULL, flenB, 0);
    ap.aepos = ex; ap.bepos = ey; ap.diffs = ed;
    __syncwarp();

    wave_pass<-1>(S, P.awords, abase, alen, P.bwords, bbase, blen, diag, anti, minp, maxp, aoff, boff, ex, ey, ed, ha, hb);
    if (S.overflow) return;
    __syncwarp();
    int preA = 0, preB = 0;
    if (lane == 0) {
        preA = wc::emit_chain_rev(reinterpret_cast<wc::Cell *>(S.cells), wc::reverse_chain(reinterpret_cast<wc::Cell *>(S.cells), ha), true, ex, ey, ed, P.TS, aoff, flenA, atr);
        preB = wc::emit_chain_rev(reinterpret_cast<wc::Cell *>(S.cells), wc::reverse_chain(reinterpret_cast<wc::Cell *>(S.cells), hb), false, ex, ey, ed, P.TS, boff, flenB, btr);
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

// Opens the next panel of a pair's hits (warp-collective); false when the pair has none left.
//   daligner (filter.c:2251-2266): a panel = the hits from nidx up to AND INCLUDING the first with apos > amark
//     (amark = amark2 + 50000, amark2 <- amark - 10000, 0 at the pair's start); the next panel starts at the first hit with
//     apos > amark2.
//   datander (scrub/tandem.c:966-971, 1112-1116): panel n = the read's positions [amarkb, amarke) with amarkb = k (n = 0) or
//     40000 n, amarke = 40000 n + 50000, replaced by alen + 1 once it reaches the read's end (>= for the first panel, > after
//     it, as written); amark2 carries n here (-1: finished) and nidx the first hit of the previous panel.
__device__ bool open_panel(const AlignArgs &P, int64_t pend, int alen, int64_t &nidx, int &amark2, int64_t &lidx, int64_t &pe) {
    const int lane = threadIdx.x & 31;
    if (!P.tandem) {
        if (nidx >= pend) return false;
        const int amark = amark2 + PANEL_SIZE; amark2 = amark - PANEL_OVERLAP;
        lidx = nidx;
        int64_t e = pend, h2 = pend;
        bool gotpe = false, goth2 = false;
        for (int64_t j0 = lidx; j0 < pend && !(gotpe && goth2); j0 += 32) {
            const int64_t j = j0 + lane;
            const int apos = (j < pend) ? hit_apos(P, j) : IMAX;
            const unsigned over = __ballot_sync(FULL, (j < pend) && apos > amark);
            const unsigned over2 = __ballot_sync(FULL, (j < pend) && apos > amark2);
            if (!goth2 && over2) { h2 = j0 + __ffs(over2) - 1; goth2 = true; }
            if (!gotpe && over) { e = j0 + __ffs(over) - 1 + 1; gotpe = true; }
        }
        if (h2 > e) h2 = e;
        pe = e; nidx = h2;
        return true;
    }
    const int n = amark2;
    if (n < 0) return false;
    const int amarkb = (n == 0) ? P.K : n * (PANEL_SIZE - PANEL_OVERLAP);
    int amarke = n * (PANEL_SIZE - PANEL_OVERLAP) + PANEL_SIZE;
    if (n == 0 ? (amarke >= alen) : (amarke > alen)) amarke = alen + 1;
    int64_t b = pend, e = pend;
    bool gotb = false, gote = false;
    for (int64_t j0 = nidx; j0 < pend && !(gotb && gote); j0 += 32) {
        const int64_t j = j0 + lane;
        const int apos = (j < pend) ? hit_apos(P, j) : IMAX;
        const unsigned inb = __ballot_sync(FULL, (j < pend) && apos >= amarkb);
        const unsigned ine = __ballot_sync(FULL, (j < pend) && apos >= amarke);
        if (!gotb && inb) { b = j0 + __ffs(inb) - 1; gotb = true; }
        if (!gote && ine) { e = j0 + __ffs(ine) - 1; gote = true; }
    }
    if (b > e) b = e;
    lidx = b; pe = e; nidx = b;
    amark2 = (amarke > alen) ? -1 : n + 1;
    return true;
}

__global__ void __launch_bounds__(AL_WARPS * 32, 3) align_pairs_strided(AlignArgs P) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * AL_WARPS + wib;
    int32_t *la_low = P.la_scratch + (size_t)gw * 3 * (MAX_LA_PER_PAIR + 1), *la_hgh = la_low + (MAX_LA_PER_PAIR + 1), *la_ae = la_hgh + (MAX_LA_PER_PAIR + 1);

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
        S.TS = P.TS; S.PATH_AVE = P.ave_path; S.reach = P.reach; S.overflow = 0; S.ncells_done = 0;
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
        const uint64_t pair = hit_pair(P, sidx);
        const int ar = (int)(pair & ((1ull << P.abits) - 1ull)), br = (int)(pair >> P.abits);
        const int alen = P.arlen[ar], blen = P.brlen[br];
        const bool selfie = P.self && !P.comp && ar == br;    // aseq == bseq pointer test, align.c:1949
        const bool doB = !P.tandem && P.symmetric && (ar != br || !P.self || !P.comp);   // filter.c:2301; datander reports one record (tandem.c:1155)
        n_pairs++;

        // extent of the pair
        int64_t pend = sidx;
        for (;;) {
            const int64_t j = pend + lane;
            const bool same = (j < P.nhits) && (hit_pair(P, j) == pair);
            const unsigned m = __ballot_sync(FULL, same);
            if (m == FULL) { pend += 32; continue; }
            pend += __ffs(~m) - 1;
            break;
        }
        int nla = 0, novl = 0;
        int amark2 = 0;
        int64_t nidx = sidx;
        for (;;) {                                             // panels (filter.c:2251-2266 / tandem.c:966-1116)
            int64_t lidx, pe;
            if (!open_panel(P, pend, alen, nidx, amark2, lidx, pe)) break;
            if (pe - lidx < P.minhit) continue;

            // pass 1: covered bases per bucket (filter.c:2268-2277)
            for (int64_t j0 = lidx; j0 < pe; j0 += 32) {
                const int64_t j = j0 + lane;
                const bool v = j < pe;
                int apos = 0, d = IMAX - lane;                 // distinct dummies so inactive lanes never match
                if (v) { apos = hit_apos(P, j); d = (apos - hit_bpos(P, j)) >> P.W; }
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
                    apos = hit_apos(P, j); bpos = hit_bpos(P, j); d = (apos - bpos) >> P.W;
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
                        if (la_low[q] <= cd && cd <= la_hgh[q] && capos <= la_ae[q]) blocked = true;
                    if (__any_sync(FULL, blocked)) continue;
                    Path ap, bp;
                    n_la++;
                    local_alignment(S, P, tbuf, ar, br, alen, blen, capos - cbpos, capos + cbpos, selfie, ap, bp);
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
                        if (lane == 0) { la_low[nla] = lo; la_hgh[nla] = hi; la_ae[nla] = ap.aepos; }
                        nla++;
                    } else if (lane == 0) atomicOr(&P.stats[3], 2ull);
                    __syncwarp();
                    if ((ap.aepos - ap.abpos) + (ap.bepos - ap.bbpos) >= P.minover) {   // filter.c:2335
                        const uint64_t seq = ((uint64_t)wi << SEQ_PAIR_SHIFT) | (uint64_t)(novl & MAX_LA_PER_PAIR);
                        emit_record(P, ap, ar + P.afirst, br + P.bfirst, seq);
                        if (doB) emit_record(P, bp, br + P.bfirst, ar + P.afirst, seq | SEQ_B);
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
                    const int apos = hit_apos(P, j);
                    const int d = (apos - hit_bpos(P, j)) >> P.W;
                    score[d] = 0; lastp[d] = 0;
                }
            }
            __syncwarp();
            if (S.overflow) break;
        }
        if (novl > 1) n_multi++;
        if (S.overflow) { if (lane == 0) atomicOr(&P.stats[3], 4ull); S.overflow = 0; }
    }
    if (lane == 0) {
        atomicAdd(&P.stats[0], n_la); atomicAdd(&P.stats[1], n_multi);
        atomicAdd(&P.stats[2], S.ncells_done); atomicAdd(&P.stats[4], n_pairs);
    }
}


// =============================================================================================
// Task pipeline (the default path).
//
// Stage 3 and stage 4 are separate kernels that meet in small tables:
//   filter_pairs   one warp per (bread, aread) pair: panel scoring + -h threshold, flags the pairs that
//                  have at least one seed candidate (filter.c:2251-2290)                     [HBM bound]
//   scan_round     one warp per active pair: the next candidate that no earlier alignment of the pair
//                  covers (the `lasta` rule, filter.c:2289, kept as a per-pair interval list)
//   wave_tasks     ONE THREAD PER TASK, a task being the forward or the reverse wave pass of one local
//                  alignment (wave_core.cuh); threads refill from a global queue, the wave fronts sit in
//                  shared memory, pebbles go to a global pool                                [integer bound]
//   finish_las     one warp per alignment: pebble chains -> trace, B path, Diagonal_Span, lasta update,
//                  MINOVER, records (align.c:1001-1111, 1711-1895, 2003-2063; filter.c:2318-2395)
// The alignments of one pair depend on each other through lasta, so a pair contributes one alignment
// per round; rounds repeat until no pair has a candidate left (almost all pairs need one round).
// A task whose front outgrows the 32-diagonal ring is rerun with a 64-diagonal ring; beyond that the
// whole pair goes to the strided warp kernel above.
// =============================================================================================
enum : uint32_t { PF_DONE = 1u, PF_REDO = 2u };

struct PairTable {          // one slot per active pair
    uint32_t *wslot;        // index into the work list = emission rank of the pair
    int64_t *sidx, *pend, *nidx, *lidx, *pe, *j;
    int32_t *amark2, *nla, *novl;
    uint32_t *flags;
    // lasta as an interval list per pair: la_off[p] / la_cap[p] name the pair's slots in the arena (la_low, la_hgh, la_ae);
    // a full list moves to a region twice as large taken from la_top
    uint32_t *la_off, *la_cap; unsigned int *la_top; uint32_t la_arena;
    int32_t *la_low, *la_hgh, *la_ae;
};
struct CandTable {          // one slot per local alignment of the current round
    uint32_t *pair;         // active pair slot
    int64_t *hidx;          // hit index of the seed
    int32_t *ar, *br, *apos, *bpos;
    wc::TaskOut *res;       // [2 * slots]: forward pass, reverse pass
};
struct PipeArgs {
    AlignArgs A;                          // hits, blocks, parameters, output buffers (shared with the strided kernel)
    const uint32_t *aw[2], *bw[2];        // [0] = the orientation that is aligned, [1] = its reverse complement
    PairTable pt; CandTable ct;
    const uint32_t *list_in; uint32_t n_in;
    uint32_t *flag;                       // filter_pairs: per work pair
    uint32_t *cand0;                      // filter_pairs: the pair's first seed candidate, as an offset from the pair's first hit
    unsigned int *ctr;                    // device counters, see CTR_*
    uint32_t *wide_list, *wide_ready; unsigned int wide_cap; int wide_ctas, linger;
    wc::Lane *wide_lane; uint32_t *wide_ring;   // state of the passes that outgrew the 32-diagonal ring: [task][32][FRONT_WORDS]
    const uint32_t *task_order;           // tasks of the round, longest expected pass first (or null: index order)
    wc::Cell *pool; unsigned int pool_cap;
    unsigned long long *wlog; unsigned int *wlog_n; unsigned int wlog_cap;   // DAMAR_WAVE_LOG: per-warp residency log (debug)
};
enum { CTR_NCAND = 0, CTR_NEXT = 1, CTR_NWIDE = 2, CTR_POOL = 3, CTR_NEXT_WIDE = 4, CTR_RETRY = 5, CTR_RETIRED = 6, CTR_N = 8 };

struct ScanState { int64_t pend, nidx, lidx, pe, j; int amark2, alen; };

// first index >= from whose pair differs (warp-collective)
__device__ int64_t pair_extent(const AlignArgs &P, int64_t from, uint64_t pair) {
    const int lane = threadIdx.x & 31;
    int64_t pend = from;
    for (;;) {
        const int64_t j = pend + lane;
        const bool same = (j < P.nhits) && (hit_pair(P, j) == pair);
        const unsigned m = __ballot_sync(FULL, same);
        if (m == FULL) { pend += 32; continue; }
        pend += __ffs(~m) - 1;
        break;
    }
    return pend;
}

// pass 1 of a panel: covered bases per diagonal bucket (filter.c:2268-2277)
__device__ void panel_score(const AlignArgs &P, int64_t lidx, int64_t pe, int32_t *score, int32_t *lastp) {
    const int lane = threadIdx.x & 31;
    for (int64_t j0 = lidx; j0 < pe; j0 += 32) {
        const int64_t j = j0 + lane;
        const bool v = j < pe;
        int apos = 0, d = IMAX - lane;                 // distinct dummies so inactive lanes never match
        if (v) { apos = hit_apos(P, j); d = (apos - hit_bpos(P, j)) >> P.W; }
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
}
// pass 3: clear the touched buckets (filter.c:2407-2411)
__device__ void panel_clear(const AlignArgs &P, int64_t lidx, int64_t pe, int32_t *score, int32_t *lastp) {
    const int lane = threadIdx.x & 31;
    for (int64_t j0 = lidx; j0 < pe; j0 += 32) {
        const int64_t j = j0 + lane;
        if (j < pe) {
            const int apos = hit_apos(P, j);
            const int d = (apos - hit_bpos(P, j)) >> P.W;
            score[d] = 0; lastp[d] = 0;
        }
    }
    __syncwarp();
}

// Walks the pair's A-panels (filter.c:2251-2266) from the saved position to the next hit that passes the -h
// test (filter.c:2289-2290) and is not covered by one of the pair's alignments so far.  Warp-collective; the
// bucket arrays are left clean.  Returns the hit index, or -1 when the pair is exhausted.
__device__ int64_t find_candidate(const AlignArgs &P, ScanState &S, int32_t *score, int32_t *lastp,
                                  int nla, const int32_t *la_low, const int32_t *la_hgh, const int32_t *la_ae) {
    const int lane = threadIdx.x & 31;
    bool scored = false;
    for (;;) {
        if (S.j < 0) {                                         // open the next panel
            int64_t lidx, pe;
            if (!open_panel(P, S.pend, S.alen, S.nidx, S.amark2, lidx, pe)) return -1;
            S.lidx = lidx; S.pe = pe;
            if (pe - lidx < P.minhit) continue;
            S.j = lidx;
        }
        if (!scored) { panel_score(P, S.lidx, S.pe, score, lastp); scored = true; }
        // pass 2 (filter.c:2283-2290) from S.j
        for (int64_t j0 = S.j; j0 < S.pe; j0 += 32) {
            const int64_t j = j0 + lane;
            const bool v = j < S.pe;
            int apos = 0, d = 0;
            bool cand = false;
            if (v) {
                apos = hit_apos(P, j); d = (apos - hit_bpos(P, j)) >> P.W;
                const int s = score[d];
                cand = (s + score[d + 1] >= P.H) || (s + score[d - 1] >= P.H);
            }
            unsigned cm = __ballot_sync(FULL, cand);
            while (cm) {
                const int l = __ffs(cm) - 1; cm &= cm - 1;
                const int capos = __shfl_sync(FULL, apos, l), cd = __shfl_sync(FULL, d, l);
                bool blocked = false;                          // apos > lasta[d] (filter.c:2289)
                for (int q = lane; q < nla; q += 32)
                    if (la_low[q] <= cd && cd <= la_hgh[q] && capos <= la_ae[q]) blocked = true;
                if (__any_sync(FULL, blocked)) continue;
                panel_clear(P, S.lidx, S.pe, score, lastp);
                S.j = j0 + l + 1;
                return j0 + l;
            }
        }
        panel_clear(P, S.lidx, S.pe, score, lastp);
        scored = false;
        S.j = -1;
    }
}

__device__ __forceinline__ int32_t *warp_buckets(const AlignArgs &P, int gw) {
    return P.bscore + (size_t)gw * 2 * P.nbuck - P.dmin + 2;          // +2 margin for d-1
}

// Stage 3 proper: which pairs have any candidate at all?
__global__ void __launch_bounds__(256) filter_pairs(PipeArgs Q) {
    const AlignArgs &P = Q.A;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    int32_t *score = warp_buckets(P, gw), *lastp = score + P.nbuck;
    for (uint32_t w = gw; w < P.nwork; w += nw) {
        const int64_t sidx = P.work[w];
        const uint64_t pair = hit_pair(P, sidx);
        ScanState S;
        S.pend = pair_extent(P, sidx, pair); S.nidx = sidx; S.j = -1; S.amark2 = 0; S.lidx = S.pe = sidx;
        S.alen = P.arlen[(int)(pair & ((1ull << P.abits) - 1ull))];
        const int64_t c = find_candidate(P, S, score, lastp, 0, nullptr, nullptr, nullptr);
        if (lane == 0) { Q.flag[w] = (c >= 0) ? 1u : 0u; Q.cand0[w] = (c >= 0) ? (uint32_t)(c - sidx) : 0u; }
    }
}
__global__ void init_pairs(PipeArgs Q, const uint32_t *__restrict__ flag, const uint64_t *__restrict__ off) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= Q.A.nwork || !flag[w]) return;
    const uint64_t p = off[w];
    Q.pt.wslot[p] = w;
    const int64_t sidx = Q.A.work[w];
    Q.pt.sidx[p] = sidx; Q.pt.pend[p] = -1; Q.pt.nidx[p] = sidx; Q.pt.lidx[p] = sidx; Q.pt.pe[p] = sidx; Q.pt.j[p] = -1;
    Q.pt.amark2[p] = 0; Q.pt.nla[p] = 0; Q.pt.novl[p] = 0; Q.pt.flags[p] = 0;
    Q.pt.la_off[p] = (uint32_t)p * LA_CAP0; Q.pt.la_cap[p] = LA_CAP0;
    // Round 1 needs no scan: filter_pairs already found the pair's first candidate.  The saved walk restarts at the pair's first
    // panel in round 2 -- seeds before this one were no candidates and this one is then covered by its own alignment's lasta
    // interval (filter.c:2289, 2318-2333), so the restart finds what a continued walk would.
    const int64_t c = sidx + (int64_t)Q.cand0[w];
    const uint64_t pair = hit_pair(Q.A, c);
    Q.ct.pair[p] = (uint32_t)p; Q.ct.hidx[p] = c;
    Q.ct.ar[p] = (int)(pair & ((1ull << Q.A.abits) - 1ull)); Q.ct.br[p] = (int)(pair >> Q.A.abits);
    Q.ct.apos[p] = hit_apos(Q.A, c); Q.ct.bpos[p] = hit_bpos(Q.A, c);
}

// One round of the lasta-ordered candidate walk: every listed pair that is still alive proposes its next seed.
__global__ void __launch_bounds__(256) scan_round(PipeArgs Q) {
    const AlignArgs &P = Q.A;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    int32_t *score = warp_buckets(P, gw), *lastp = score + P.nbuck;
    unsigned long long n_multi = 0;
    for (uint32_t i = gw; i < Q.n_in; i += nw) {
        const uint32_t p = Q.list_in ? Q.list_in[i] : i;
        if (Q.pt.flags[p] & (PF_DONE | PF_REDO)) continue;
        const int64_t sidx = Q.pt.sidx[p];
        const uint64_t pair = hit_pair(P, sidx);
        ScanState S;
        S.pend = Q.pt.pend[p];
        if (S.pend < 0) S.pend = pair_extent(P, sidx, pair);
        S.nidx = Q.pt.nidx[p]; S.lidx = Q.pt.lidx[p]; S.pe = Q.pt.pe[p]; S.j = Q.pt.j[p]; S.amark2 = Q.pt.amark2[p];
        S.alen = P.arlen[(int)(pair & ((1ull << P.abits) - 1ull))];
        const int nla = Q.pt.nla[p];
        const size_t lo = Q.pt.la_off[p];
        const int64_t c = find_candidate(P, S, score, lastp, nla, Q.pt.la_low + lo, Q.pt.la_hgh + lo, Q.pt.la_ae + lo);
        if (lane == 0) {
            Q.pt.pend[p] = S.pend; Q.pt.nidx[p] = S.nidx; Q.pt.lidx[p] = S.lidx; Q.pt.pe[p] = S.pe; Q.pt.j[p] = S.j; Q.pt.amark2[p] = S.amark2;
            if (c < 0) {
                Q.pt.flags[p] |= PF_DONE;
                if (Q.pt.novl[p] > 1) n_multi++;
            } else {
                const unsigned slot = atomicAdd(&Q.ctr[CTR_NCAND], 1u);
                const int apos = hit_apos(P, c);
                Q.ct.pair[slot] = p; Q.ct.hidx[slot] = c;
                Q.ct.ar[slot] = (int)(pair & ((1ull << P.abits) - 1ull)); Q.ct.br[slot] = (int)(pair >> P.abits);
                Q.ct.apos[slot] = apos; Q.ct.bpos[slot] = hit_bpos(P, c);
            }
        }
        __syncwarp();
    }
    if (lane == 0 && n_multi) atomicAdd(&P.stats[1], n_multi);
}

// Longest-processing-time-first order for the task queue: a pass can run at most to the nearer read end.
__global__ void task_length_keys(PipeArgs Q, unsigned ntasks, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntasks) return;
    const unsigned slot = t >> 1;
    const int ar = Q.ct.ar[slot], br = Q.ct.br[slot], apos = Q.ct.apos[slot], bpos = Q.ct.bpos[slot];
    int len;
    if (t & 1u) len = min(apos, bpos);
    else len = min(Q.A.arlen[ar] - apos, Q.A.brlen[br] - bpos);
    if (len < 0) len = 0;
    if (len > 0xffffff) len = 0xffffff;
    keys[t] = (uint64_t)(0xffffff - len);
    vals[t] = t;
}

// The sorted queue hands the 32 longest passes to the first warp that asks, and that warp -- 32 long passes in 32 different
// phases of the wave step -- is the last to finish.  Deal the head of the queue out like cards instead: chunk c (one warp's
// first batch) gets sorted[c], sorted[chunks + c], sorted[2 chunks + c], ...: one pass of every length band.
__global__ void spread_order(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, unsigned n, unsigned chunks) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned c = i >> 5, l = i & 31u;
    out[i] = (c < chunks) ? in[l * chunks + c] : in[i];
}

// ---- stage 4: one thread per wave pass ----
struct DevPool {
    wc::Cell *cells; unsigned int *top; unsigned int cap;
    __device__ uint32_t take(int n) {
        const unsigned int b = atomicAdd(top, (unsigned int)n);
        return (b + (unsigned int)n <= cap) ? b : wc::NOCELL;
    }
    __device__ void store(uint32_t id, const wc::Cell &c) { *reinterpret_cast<int4 *>(cells + id) = make_int4(c.prev, c.diag, c.diff, c.mark); }
};

template <int WARPS_>
struct WaveCfg {
    static constexpr int WARPS = WARPS_;                      // 16 KB of front per warp: 32 lanes x 32 diagonals or 16 lanes x 64
    static constexpr int SMEM = WARPS * wc::FRONT_WORDS * wc::FrontRT::ARR * 4;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// DAMAR_WAVE_LOG=<file>: every warp of every wave_tasks launch of the process logs {start, first idle lane, end, SM | shape |
// grid} (ns of %globaltimer); written as raw u64 quadruples at exit.  A debugging aid for the step's SM occupancy.
struct WaveLog {
    unsigned long long *buf = nullptr; unsigned int *n = nullptr; unsigned int cap = 0; std::string path;
    static WaveLog &get() {
        static WaveLog w;
        static std::once_flag once;
        std::call_once(once, [] {
            const char *e = getenv("DAMAR_WAVE_LOG");
            if (!e || !*e) return;
            w.path = e; w.cap = 4u << 20;
            CUDA_CHECK(cudaMalloc(&w.buf, (size_t)w.cap * 32)); CUDA_CHECK(cudaMalloc(&w.n, 4)); CUDA_CHECK(cudaMemset(w.n, 0, 4));
            atexit([] {
                WaveLog &g = WaveLog::get();
                unsigned int cnt = 0;
                if (cudaMemcpy(&cnt, g.n, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return;
                if (cnt > g.cap) cnt = g.cap;
                std::vector<unsigned long long> h((size_t)cnt * 4);
                if (cnt && cudaMemcpy(h.data(), g.buf, (size_t)cnt * 32, cudaMemcpyDeviceToHost) != cudaSuccess) return;
                if (FILE *f = fopen(g.path.c_str(), "wb")) { fwrite(h.data(), 32, cnt, f); fclose(f); }
            });
        });
        return w;
    }
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) { return *reinterpret_cast<const volatile unsigned *>(p); }

// Task t of [0, 2 * ncand) = pass (t & 1) of alignment t >> 1.  A warp works in one of two SHAPES on the same 16 KB of
// shared memory: NARROW -- 32 lanes, a 32-diagonal ring each, tasks drawn from the round's queue -- or WIDE -- 16 lanes, a
// 64-diagonal ring each, resuming the passes that narrow lanes had to PARK because their front outgrew 32 diagonals
// (state + ring copied to global memory, a ready flag raised).  A few CTAs of a launch are wide from the start, so a parked
// pass continues within microseconds while the bulk of the round is still running, and the narrow CTAs leave as soon as
// the queue is empty.  (Round 1 ran the parked passes in a second kernel after the first, with a host wait in between.)  A wide lane leaves when every task has left the narrow queue's lanes
// (CTR_RETIRED == number of tasks: no pass can be parked any more) and no parked pass is unclaimed.
template <int NWAY, int WARPS, int MINB = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB) wave_tasks(PipeArgs Q) {
    extern __shared__ uint32_t front_smem[];
    __shared__ int16_t s_t7[128], s_t8[256];                  // the trim test tables (wave_core.cuh)
    const AlignArgs &P = Q.A;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 384; i += blockDim.x) {
        if (i < 128) s_t7[i] = (int16_t)wc::min_suffix_score(i, 7, P.dscore);
        else s_t8[i - 128] = (int16_t)wc::min_suffix_score(i - 128, 8, P.dscore);
    }
    __syncthreads();
    const wc::TrimTabs tabs{s_t7, s_t8, P.dscore};
    uint32_t *region = front_smem + (size_t)wib * wc::FRONT_WORDS * wc::FrontRT::ARR;
    DevPool pool{Q.pool, &Q.ctr[CTR_POOL], Q.pool_cap};
    const unsigned ntasks = 2u * Q.ctr[CTR_NCAND];
    unsigned retired = 0;                                     // tasks this lane took off the narrow queue and finished or parked

    wc::Lane L;
    L.status = wc::ST_IDLE; L.k = 0; L.low = 0;
    unsigned task = 0;
    unsigned long long cells_done = 0;
    unsigned long long t_begin = 0, t_dry = 0;                // debug log: start, first lane without a task
    if (Q.wlog) t_begin = globaltimer_ns();
    // Warp-uniform shape.  A full-size round is TWO launches of this kernel, back to back with no host wait between them:
    // every CTA narrow, leaving when the queue is empty (their SMs go to the next kernel, usually another comparison's),
    // then a few CTAs that are wide from the start (wide_ctas == gridDim.x) for the parked passes -- about 2 % of the
    // passes, but long ones, which would otherwise keep every SM of the first launch waiting.  A small round is one
    // launch whose narrow warps turn wide at the end (linger).
    const bool linger = Q.linger != 0;
    bool wide = (int)blockIdx.x < Q.wide_ctas;
    wc::FrontRT F = wide ? wc::FrontRT{region + (lane & 15), 63, 4} : wc::FrontRT{region + lane, 31, 5};
    // a launch of wide CTAs only starts when the narrow launch is over: the number of parked passes is final, and a warp
    // takes no more of them than its share (a warp with few lanes steps through fewer phases per iteration)
    int wide_lanes = 16;
    if (wide) {
        const unsigned nw = ld_volatile_u32(&Q.ctr[CTR_NWIDE]), warps = gridDim.x * WARPS;
        wide_lanes = (int)min(16u, max(1u, (nw + warps - 1) / warps));
    }
    if (wide && lane >= wide_lanes) L.status = wc::ST_EXIT;
    unsigned await = 0;
    bool waiting = false;
    for (unsigned iter = 0;; iter++) {
        bool any_run = true;
        if (!wide) {
            // ---------------- NARROW: 32 lanes x 32 diagonals, tasks from the round's queue ----------------
            if (L.status != wc::ST_RUN && L.status != wc::ST_EXIT) {
                if (L.status != wc::ST_IDLE) {                    // retire the finished task
                    wc::TaskOut o;
                    wc::finish(L, o);
                    Q.ct.res[task] = o;
                    retired++;
                    if (L.status != wc::ST_WIDE) cells_done += o.cells;
                    else {                                        // park the pass: a wide lane resumes it where it stands
                        const unsigned w = atomicAdd(&Q.ctr[CTR_NWIDE], 1u);
                        Q.wide_list[w] = task;
                        Q.wide_lane[w] = L;
                        uint32_t *dst = Q.wide_ring + (size_t)w * 32 * wc::FRONT_WORDS;
                        for (int k = L.low; k <= L.hgh; k++) {
                            const uint32_t *sl = F.slot(k);
#pragma unroll
                            for (int a = 0; a < wc::FRONT_WORDS; a++) dst[(k - L.low) * wc::FRONT_WORDS + a] = sl[a * wc::FrontRT::ARR];
                        }
                        __threadfence();
                        *reinterpret_cast<volatile uint32_t *>(Q.wide_ready + w) = 1u;
                    }
                }
                // fetch the next task (one atomic per warp-wide batch of idle lanes)
                const unsigned idle = __activemask();
                const int leader = __ffs(idle) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(&Q.ctr[CTR_NEXT], (unsigned)__popc(idle));
                base = __shfl_sync(idle, base, leader);
                const unsigned t = base + __popc(idle & ((1u << lane) - 1u));
                if (t >= ntasks) { L.status = wc::ST_EXIT; if (Q.wlog && !t_dry) t_dry = globaltimer_ns(); }
                else {
                    task = Q.task_order ? Q.task_order[t] : t;
                    const unsigned slot = task >> 1;
                    const int ar = Q.ct.ar[slot], br = Q.ct.br[slot], apos = Q.ct.apos[slot], bpos = Q.ct.bpos[slot];
                    wc::TaskIn in;
                    in.rev = (int)(task & 1u);
                    in.aw = Q.aw[in.rev]; in.bw = Q.bw[in.rev];
                    in.abase = P.aboff[ar]; in.bbase = P.bboff[br];
                    in.alen = P.arlen[ar]; in.blen = P.brlen[br];
                    in.diag = apos - bpos; in.anti = apos + bpos;
                    in.aoff = 0; in.boff = P.comp ? (in.blen % P.TS) : 0;                     // align.c:1970-1984
                    const bool selfie = P.self && !P.comp && ar == br;                       // align.c:1949
                    in.minp = (selfie && in.diag >= 0) ? 1 : -IMAX;                          // align.c:1951-1968
                    in.maxp = (selfie && in.diag <= 0) ? -1 : IMAX;
                    in.TS = P.TS; in.path_ave = P.ave_path; in.reach = P.reach;
                    wc::setup(L, in, F, pool);
                }
            }
            if (__all_sync(FULL, L.status == wc::ST_EXIT)) {      // the queue is empty
                __threadfence();                                  // this lane's parked passes are visible before its tasks count as retired
                if (retired) atomicAdd(&Q.ctr[CTR_RETIRED], retired);
                __syncwarp();
                if (!linger) break;                               // the wide CTAs take care of the parked passes
                wide = true;
                F = wc::FrontRT{region + (lane & 15), 63, 4};
                L.status = (lane < 16) ? wc::ST_IDLE : wc::ST_EXIT;
                continue;
            }
        } else {
            // ---------------- WIDE: 16 lanes x 64 diagonals, parked passes ----------------
            const bool mine = lane < wide_lanes;
            if (mine && !waiting && L.status != wc::ST_RUN && L.status != wc::ST_EXIT) {
                if (L.status != wc::ST_IDLE) {                    // retire; a front beyond 64 diagonals stays ST_WIDE: the pair goes to the strided kernel
                    wc::TaskOut o;
                    wc::finish(L, o);
                    Q.ct.res[task] = o;
                    cells_done += o.cells;
                }
                await = atomicAdd(&Q.ctr[CTR_NEXT_WIDE], 1u);
                waiting = true;
                L.status = wc::ST_IDLE;
            }
            any_run = __any_sync(FULL, L.status == wc::ST_RUN);
            if (mine && waiting && (!any_run || (iter & 7u) == 0)) {   // polling costs an L2 round trip: rarely while neighbours run
                if (await < Q.wide_cap && ld_volatile_u32(Q.wide_ready + await) != 0u) {
                    __threadfence();
                    task = Q.wide_list[await];
                    L = Q.wide_lane[await];
                    const uint32_t *src = Q.wide_ring + (size_t)await * 32 * wc::FRONT_WORDS;
                    for (int k = L.low; k <= L.hgh; k++) {
                        uint32_t *sl = F.slot(k);
#pragma unroll
                        for (int a = 0; a < wc::FRONT_WORDS; a++) sl[a * wc::FrontRT::ARR] = src[(k - L.low) * wc::FRONT_WORDS + a];
                    }
                    // transition() runs again from the top: clipping and pruning are idempotent, only its counters are not
                    L.waves -= 1; L.cells -= (uint32_t)(L.hgh - L.low + 1);
                    L.k = L.low - 1; L.status = wc::ST_RUN;
                    waiting = false;
                } else if (ld_volatile_u32(&Q.ctr[CTR_RETIRED]) >= ntasks) {
                    __threadfence();
                    if (await >= ld_volatile_u32(&Q.ctr[CTR_NWIDE])) { L.status = wc::ST_EXIT; waiting = false; }
                }
            }
            if (__all_sync(FULL, L.status == wc::ST_EXIT)) break;
        }
        if (L.status == wc::ST_RUN) {
            if (L.k < L.low) wc::transition(L, F);
            if (L.status == wc::ST_RUN && L.k >= L.low) wc::cellsN<NWAY>(L, F, tabs, pool);
        } else if (!any_run) __nanosleep(256);
    }
    if (cells_done) atomicAdd(&P.stats[2], cells_done);
    if (Q.wlog) {
        if (!t_dry) t_dry = ~0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(FULL, t_dry, o); if (v < t_dry) t_dry = v; }
        if (lane == 0) {
            const unsigned i = atomicAdd(Q.wlog_n, 1u);
            if (i < Q.wlog_cap) {
                unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                Q.wlog[4 * (size_t)i] = t_begin; Q.wlog[4 * (size_t)i + 1] = t_dry; Q.wlog[4 * (size_t)i + 2] = globaltimer_ns();
                Q.wlog[4 * (size_t)i + 3] = (unsigned long long)smid | ((unsigned long long)(Q.wide_ctas ? 1 : 0) << 16) | ((unsigned long long)gridDim.x << 32);
            }
        }
    }
}

// ---- stage 4 epilogue: one warp per alignment of the round ----
__global__ void __launch_bounds__(256) finish_las(PipeArgs Q) {
    const AlignArgs &P = Q.A;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    uint16_t *tbuf = P.tbuf + (size_t)gw * 8 * P.maxtp;
    uint16_t *atr = tbuf + 2 * P.maxtp, *btr = tbuf + 6 * P.maxtp;
    const unsigned ncand = Q.ctr[CTR_NCAND];
    unsigned long long n_la = 0;
    for (unsigned slot = gw; slot < ncand; slot += nw) {
        const uint32_t p = Q.ct.pair[slot];
        const wc::TaskOut fo = Q.ct.res[2 * slot], ro = Q.ct.res[2 * slot + 1];
        const uint32_t wi = Q.pt.wslot[p];
        if (fo.status != wc::ST_DONE || ro.status != wc::ST_DONE) {
            const bool pool_only = (fo.status == wc::ST_DONE || fo.status == wc::ST_POOL) && (ro.status == wc::ST_DONE || ro.status == wc::ST_POOL);
            if (lane == 0) {
                if (pool_only) { Q.pt.j[p] = Q.ct.hidx[slot]; atomicAdd(&Q.ctr[CTR_RETRY], 1u); }     // pebble pool ran dry: same seed next round
                else {                                                                                 // too wide / out of range: strided kernel
                    Q.pt.flags[p] |= PF_REDO;
                    atomicAdd(&P.stats[0], (unsigned long long)(-(long long)Q.pt.nla[p]));              // it recounts the pair's alignments
                    const unsigned q = atomicAdd(P.redo_count, 1u);
                    P.redo_sidx[q] = (uint32_t)Q.pt.sidx[p]; P.redo_wi[q] = wi; P.pair_redo[wi] = 1;
                }
            }
            continue;
        }
        n_la++;
        const int ar = Q.ct.ar[slot], br = Q.ct.br[slot], apos = Q.ct.apos[slot], bpos = Q.ct.bpos[slot];
        const int alen = P.arlen[ar], blen = P.brlen[br];
        const int anti = apos + bpos;
        const int aoff = 0, boff = P.comp ? (blen % P.TS) : 0;
        int flen = 0, pre = 0, rroot = -1;
        if (lane < 2)                                          // lane 0: the A trace points of the chains, lane 1: the B trace points
            flen = wc::emit_chain_fwd<true>(Q.pool, fo.ha, lane == 0, anti, fo.ex, fo.ey, fo.ediff, lane == 0 ? atr : btr);
        if (lane == 0) rroot = wc::reverse_chain(Q.pool, ro.ha);
        __syncwarp();
        rroot = __shfl_sync(FULL, rroot, 0);
        if (lane < 2)
            pre = wc::emit_chain_rev<true>(Q.pool, rroot, lane == 0, ro.ex, ro.ey, ro.ediff, P.TS, lane == 0 ? aoff : boff, flen, lane == 0 ? atr : btr);
        __syncwarp();
        const int flenA = __shfl_sync(FULL, flen, 0), flenB = __shfl_sync(FULL, flen, 1);
        const int preA = __shfl_sync(FULL, pre, 0), preB = __shfl_sync(FULL, pre, 1);
        Path ap, bp;
        ap.aepos = fo.ex; ap.bepos = fo.ey; ap.abpos = ro.ex; ap.bbpos = ro.ey; ap.diffs = fo.ediff + ro.ediff;
        ap.tlen = flenA + preA; ap.trace = atr - preA;
        bp.tlen = flenB + preB; bp.trace = btr - preB; bp.diffs = ap.diffs;
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
        // Diagonal_Span (filter.c:2079-2110) on the uncompressed A trace, then lasta (filter.c:2318-2333)
        const int cd = (apos - bpos) >> P.W;
        int lo, hi;
        {
            int dd = ap.abpos - ap.bbpos; lo = hi = dd;
            dd = ap.aepos - ap.bepos; if (dd < lo) lo = dd; else if (dd > hi) hi = dd;
            const int nint = (ap.tlen - 2) / 2;
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
        const int nla = Q.pt.nla[p], novl = Q.pt.novl[p];
        if (nla < MAX_LA_PER_PAIR) {
            uint32_t off = Q.pt.la_off[p];
            const uint32_t cap = Q.pt.la_cap[p];
            bool room = true;
            if ((uint32_t)nla == cap) {                            // the pair's list is full: move it to a region twice as large
                uint32_t noff = 0;
                if (lane == 0) noff = atomicAdd(Q.pt.la_top, 2u * cap);
                noff = __shfl_sync(FULL, noff, 0);
                if ((uint64_t)noff + 2ull * cap > (uint64_t)Q.pt.la_arena) { room = false; if (lane == 0) atomicOr(&P.stats[3], 32ull); }
                else {
                    for (int q = lane; q < nla; q += 32) {
                        Q.pt.la_low[noff + q] = Q.pt.la_low[off + q]; Q.pt.la_hgh[noff + q] = Q.pt.la_hgh[off + q]; Q.pt.la_ae[noff + q] = Q.pt.la_ae[off + q];
                    }
                    __syncwarp();
                    if (lane == 0) { Q.pt.la_off[p] = noff; Q.pt.la_cap[p] = 2u * cap; }
                    off = noff;
                }
            }
            if (room && lane == 0) {
                const size_t o = (size_t)off + nla;
                Q.pt.la_low[o] = lo; Q.pt.la_hgh[o] = hi; Q.pt.la_ae[o] = ap.aepos; Q.pt.nla[p] = nla + 1;
            }
        } else if (lane == 0) atomicOr(&P.stats[3], 2ull);
        if ((ap.aepos - ap.abpos) + (ap.bepos - ap.bbpos) >= P.minover) {   // filter.c:2335
            const bool doB = !P.tandem && P.symmetric && (ar != br || !P.self || !P.comp);   // filter.c:2301; datander: tandem.c:1155
            const uint64_t seq = ((uint64_t)wi << SEQ_PAIR_SHIFT) | (uint64_t)(novl & MAX_LA_PER_PAIR);
            emit_record(P, ap, ar + P.afirst, br + P.bfirst, seq);
            if (doB) emit_record(P, bp, br + P.bfirst, ar + P.afirst, seq | SEQ_B);
            if (lane == 0) Q.pt.novl[p] = novl + 1;
        }
        __syncwarp();
    }
    if (lane == 0 && n_la) atomicAdd(&P.stats[0], n_la);
}

// New_Align_Spec + set_table (align.c:226-318): floating point only here, evaluated exactly as the
// reference does (float sum of freq[0]+freq[3], double afterwards).
double Bias_Factor[10] = {.690, .690, .690, .690, .780, .850, .900, .933, .966, 1.000};
void build_trim_tables(const damar_options &opt, int &ave_path, int &dscore_out, std::vector<int16_t> &score, std::vector<int16_t> &table) {
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
    dscore_out = dscore;
    score.resize(TRIM_MASK + 1); table.resize(TRIM_MASK + 1);
    for (int x = 0; x <= TRIM_MASK; x++) {
        int s = 0, mx = 0;
        for (int bit = TRIM_LEN - 1; bit >= 0; bit--) { if (s > mx) mx = s; s += ((x >> bit) & 1) ? mscore : -dscore; }
        table[x] = (int16_t)(s - mx); score[x] = (int16_t)s;
    }
}

template <typename T> void set_smem(T kernel, int bytes) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
}

}  // namespace


// ---- Handle_Redundancies (filter.c:1804-2077, called at :2438-2449) on the device ----
// Runs after the last round over the finished record arrays: records are sorted by emission rank (= pair, A before B,
// alignment order), one THREAD per pair with more than one alignment applies path_fuse_core.cuh to the pair's paths and
// rewrites the pair's records in place (new traces are appended to the trace pool, dropped records are flagged dead).
namespace {

struct HrArgs {
    damar_overlap *recs; uint64_t *toff, *seq; uint8_t *dead; unsigned long long nrec, nrec_fast;
    const uint8_t *pair_redo;                 // pairs the task pipeline gave up on: what it had emitted for them is void
    const uint32_t *order;                    // record indices sorted by emission rank
    void *trace; int tbytes; unsigned long long trace_cap; unsigned long long *trace_count;
    const uint32_t *aw, *bw; const uint32_t *aboff, *bboff; const int32_t *arlen, *brlen; int afirst, bfirst;
    int comp, ts, bridge;                     // bridge == 0: datander's Handle_Redundancies (tandem.c:767), fusion and containment only
    uint32_t *pool; unsigned long long pool_cap; unsigned long long *pool_top;
    unsigned long long *stats;                // [3] |= 1 trace pool full, |= 8 value > 255 in an 8-bit trace, |= 16 scratch pool dry
    unsigned long long *nfused;
};

struct PackedSeq {                            // base i of one read of a packed block
    const uint32_t *words; uint32_t base;
    __device__ int operator()(int i) const { const uint32_t p = base + (uint32_t)i; return (int)((words[p >> 4] >> (30 - 2 * (int)(p & 15u))) & 3u); }
};
struct PoolArena {                            // bump allocation from one global pool, 4-byte units
    uint32_t *pool; unsigned long long cap; unsigned long long *top; bool dry;
    __device__ uint32_t *grab(unsigned long long words) {
        words = (words + 1ull) & ~1ull;                     // 8-byte granules: Path arrays hold pointers
        const unsigned long long o = atomicAdd(top, words);
        if (o + words > cap) { dry = true; return nullptr; }
        return pool + o;
    }
    __device__ uint16_t *take(int n) { return reinterpret_cast<uint16_t *>(grab((unsigned long long)(n + 1) / 2 + 1)); }
    __device__ int *take_ints(int n) { return reinterpret_cast<int *>(grab((unsigned long long)n + 1)); }
};

__global__ void hr_sort_keys(HrArgs H, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H.nrec) return;
    const uint64_t sq = H.seq[i];
    const bool dead = (i < H.nrec_fast) && H.pair_redo != nullptr && H.pair_redo[sq >> SEQ_PAIR_SHIFT];
    H.dead[i] = dead ? 1 : 0;
    keys[i] = dead ? ~0ull : sq;
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(64) handle_redundancies_pairs(HrArgs H) {
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= H.nrec) return;
    const uint32_t i0 = H.order[j];
    if (H.dead[i0]) return;
    const uint64_t wi = H.seq[i0] >> SEQ_PAIR_SHIFT;
    if (j > 0) { const uint32_t ip = H.order[j - 1]; if (!H.dead[ip] && (H.seq[ip] >> SEQ_PAIR_SHIFT) == wi) return; }   // not the pair's first record
    // the pair's records are contiguous in rank order: its A records (alignment order), then its B records
    int na = 0, nb = 0;
    for (unsigned long long q = j; q < H.nrec; q++) {
        const uint32_t i = H.order[q];
        if (H.dead[i] || (H.seq[i] >> SEQ_PAIR_SHIFT) != wi) break;
        if (H.seq[i] & SEQ_B) nb++; else na++;
    }
    const uint32_t *ia = H.order + j, *ib = H.order + j + na;
    if (na < 2 && nb < 2) return;

    PoolArena al{H.pool, H.pool_cap, H.pool_top, false};
    pf::Path *am = reinterpret_cast<pf::Path *>(al.grab((unsigned long long)(na + nb + 2) * (sizeof(pf::Path) / 4)));
    if (am == nullptr) { atomicOr(H.stats + 3, 16ull); return; }
    pf::Path *bm = am + na + 1;
    auto load = [&](uint32_t i, pf::Path &p) {
        const damar_overlap r = H.recs[i];
        p.abpos = r.abpos; p.bbpos = r.bbpos; p.aepos = r.aepos; p.bepos = r.bepos; p.diffs = r.diffs; p.tlen = r.tlen;
        p.trace = al.take(r.tlen);
        if (p.trace == nullptr) return;
        const uint64_t o = H.toff[i];
        if (H.tbytes == 1) { const uint8_t *t = reinterpret_cast<const uint8_t *>(H.trace) + o; for (int k = 0; k < r.tlen; k++) p.trace[k] = t[k]; }
        else { const uint16_t *t = reinterpret_cast<const uint16_t *>(H.trace) + o; for (int k = 0; k < r.tlen; k++) p.trace[k] = t[k]; }
    };
    for (int k = 0; k < na; k++) load(ia[k], am[k]);
    for (int k = 0; k < nb; k++) load(ib[k], bm[k]);
    if (al.dry) { atomicOr(H.stats + 3, 16ull); return; }

    const damar_overlap first = H.recs[na > 0 ? ia[0] : ib[0]];
    const int ar = (na > 0 ? first.aread - H.afirst : first.bread - H.afirst);
    const int br = (na > 0 ? first.bread - H.bfirst : first.aread - H.bfirst);
    const PackedSeq sa{H.aw, H.aboff[ar]}, sb{H.bw, H.bboff[br]};
    const int alen = H.arlen[ar], blen = H.brlen[br];
    int sa_n = na, sb_n = nb;
    if (na > 1) {
        const bool both = nb > 1;
        const int no = pf::handle_redundancies(am, both ? bm : nullptr, na, H.comp, H.ts, sa, alen, sb, blen, al, H.bridge != 0);
        if (no < 0) { atomicOr(H.stats + 3, 16ull); return; }
        sa_n = no; if (both) sb_n = no;
    } else {
        const int no = pf::handle_redundancies(bm, (pf::Path *)nullptr, nb, H.comp, H.ts, sa, alen, sb, blen, al, H.bridge != 0);
        if (no < 0) { atomicOr(H.stats + 3, 16ull); return; }
        sb_n = no;
    }
    if (sa_n != na || sb_n != nb) atomicAdd(H.nfused, 1ull);

    auto store = [&](uint32_t i, const pf::Path &p, uint64_t sq) {
        const unsigned long long to = atomicAdd(H.trace_count, (unsigned long long)p.tlen);
        if (to + (unsigned long long)p.tlen > H.trace_cap) { atomicOr(H.stats + 3, 1ull); return; }
        damar_overlap r = H.recs[i];
        r.abpos = p.abpos; r.bbpos = p.bbpos; r.aepos = p.aepos; r.bepos = p.bepos; r.diffs = p.diffs; r.tlen = p.tlen;
        H.recs[i] = r; H.toff[i] = to; H.seq[i] = sq;
        if (H.tbytes == 1) {                                 // Compress_TraceTo8 (align.c:3375-3396)
            uint8_t *t = reinterpret_cast<uint8_t *>(H.trace) + to;
            bool big = false;
            for (int k = 0; k < p.tlen; k++) { big |= p.trace[k] > 255; t[k] = (uint8_t)p.trace[k]; }
            if (big) atomicOr(H.stats + 3, 8ull);
        } else {
            uint16_t *t = reinterpret_cast<uint16_t *>(H.trace) + to;
            for (int k = 0; k < p.tlen; k++) t[k] = p.trace[k];
        }
    };
    for (int k = 0; k < na; k++) { if (k < sa_n) store(ia[k], am[k], (wi << SEQ_PAIR_SHIFT) | (uint64_t)k); else H.dead[ia[k]] = 1; }
    for (int k = 0; k < nb; k++) { if (k < sb_n) store(ib[k], bm[k], (wi << SEQ_PAIR_SHIFT) | SEQ_B | (uint64_t)k); else H.dead[ib[k]] = 1; }
}

}  // namespace

__global__ void add_seq_base(uint64_t *__restrict__ seq, unsigned long long n, uint64_t add) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) seq[i] += add;
}

int report_and_align(const DeviceBlock *ab, const DeviceBlock *bb, int comp, int self, const SeedHits &hits, HostOverlaps &out, bool tandem,
                     const HitSpan *span) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const FilterParams &fp = c.fp;
    const int64_t n = hits.n;
    const int64_t hit_base = span ? span->base : 0, hit_total = span ? span->total : n;
    const uint64_t pair_base = span ? span->pair_base : 0;
    // datander walks every read and has neither the minhit rule nor report slices (scrub/tandem.c:895-1000)
    const int minhit = tandem ? 1 : (fp.hitmin - 1) / fp.kmer + 1;
    if (n >= (int64_t)0xffffffffll) { set_error("more than 2^32 seed hits in one comparison (%lld): lower -M or -t", (long long)n); return 1; }

    // ---- work list: first hit of every pair with at least minhit hits (filter.c:2210-2221) ----
    HostProf *hp1 = new HostProf("ra.worklist");
    Timer t_pairs(st);
    const int nslices = tandem ? 0 : (fp.nthreads_pow2 > 1024 ? 1024 : fp.nthreads_pow2);
    DevBuf<int64_t> ends; ends.reserve(fp.nthreads_pow2);
    DevBuf<uint32_t> flag; flag.reserve((size_t)n);
    DevBuf<uint64_t> off; off.reserve((size_t)n);
    DevBuf<uint64_t> total; total.reserve(1);
    const int kshift = hits.bpos ? 0 : hits.qbits;
    if (nslices > 0) { slice_ends<<<1, 1024, 0, st>>>(hits.keys, n, nslices, fp.nshift, hits.pbits + hits.abits + kshift, ends.p, hit_base, hit_total); KERNEL_CHECK(); }
    const unsigned hb = (unsigned)((n + 255) / 256);
    flag_work_pairs<<<hb, 256, 0, st>>>(hits.keys, n, hits.pbits + kshift, minhit, ends.p, nslices, flag.p);
    KERNEL_CHECK();
    exclusive_scan_u32_to_u64(flag.p, off.p, n, total.p, st);
    uint64_t nwork = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nwork, total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    DevBuf<uint32_t> work; work.reserve((size_t)nwork + 1);
    collect_work_pairs<<<hb, 256, 0, st>>>(flag.p, off.p, n, work.p);
    KERNEL_CHECK();
    count_launch(3);
    flag.release(); off.release();
    if (nwork == 0) { c.tm.pairs_ms = t_pairs.stop(); delete hp1; return 0; }

    // ---- common arguments ----
    int ave_path, dscore; std::vector<int16_t> hs, ht;
    build_trim_tables(c.opt, ave_path, dscore, hs, ht);
    DevBuf<int16_t> dscoret, dtable; dscoret.reserve(hs.size()); dtable.reserve(ht.size());
    CUDA_CHECK(cudaMemcpyAsync(dscoret.p, hs.data(), hs.size() * 2, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(dtable.p, ht.data(), ht.size() * 2, cudaMemcpyHostToDevice, st));

    const int TS = c.opt.trace_spacing;
    const int lblocks = NUM_SMS_B200 * 3;                    // strided kernel
    const int sblocks = NUM_SMS_B200 * 8;                    // scan / filter / finish kernels: 8 warps per CTA
    const int nwarps = std::max(lblocks * AL_WARPS, sblocks * 8);
    const int ring = 2048;
    const int maxlen = ab->maxlen > bb->maxlen ? ab->maxlen : bb->maxlen;
    const int maxtp = 2 * (maxlen / TS + 2) + 2;
    const int dmin = (-bb->maxlen) >> fp.binshift, dmax = ab->maxlen >> fp.binshift;
    const int nbuck = dmax - dmin + 1 + 4;
    const int tbytes = (TS <= 125) ? 1 : 2;                  // TRACE_XOVR, align.h:58

    DevBuf<int32_t> bscore; DevBuf<uint16_t> tbuf;
    bscore.reserve((size_t)nwarps * 2 * nbuck); tbuf.reserve((size_t)nwarps * 8 * maxtp);
    CUDA_CHECK(cudaMemsetAsync(bscore.p, 0, sizeof(int32_t) * (size_t)nwarps * 2 * nbuck, st));
    DevBuf<unsigned long long> counters; counters.reserve(8);
    DevBuf<unsigned int> next; next.reserve(4);
    DevBuf<unsigned int> ctr; ctr.reserve(CTR_N);
    DevBuf<uint32_t> redo_sidx, redo_wi; DevBuf<uint8_t> pair_redo;
    redo_sidx.reserve((size_t)nwork + 1); redo_wi.reserve((size_t)nwork + 1); pair_redo.reserve((size_t)nwork + 1);
    DevBuf<uint32_t> pflag; pflag.reserve((size_t)nwork + 1);
    DevBuf<uint32_t> pcand; pcand.reserve((size_t)nwork + 1);
    DevBuf<uint64_t> poff; poff.reserve((size_t)nwork + 1);

    PipeArgs Q{};
    AlignArgs &P = Q.A;
    P.keys = hits.keys; P.bpos = hits.bpos; P.nhits = n; P.pbits = hits.pbits; P.abits = hits.abits;
    P.kshift = kshift; P.qmask = kshift ? ((1u << kshift) - 1u) : 0u;
    P.work = work.p; P.nwork = (uint32_t)nwork; P.next_work = next.p; P.work_wi = nullptr;
    P.redo_sidx = redo_sidx.p; P.redo_wi = redo_wi.p; P.redo_count = next.p + 1; P.pair_redo = pair_redo.p;
    const int bsel = (comp && !bb->as_is) ? 1 : 0;
    P.awords = ab->packed[0]; P.bwords = bb->packed[bsel];
    Q.aw[0] = ab->packed[0]; Q.aw[1] = ab->packed[1]; Q.bw[0] = bb->packed[bsel]; Q.bw[1] = bb->packed[1 - bsel];
    if (!Q.aw[1] || !Q.bw[1]) { set_error("report_and_align: reverse-complemented reads are not resident"); delete hp1; return 1; }
    P.aboff = ab->boff; P.bboff = bb->boff; P.arlen = ab->rlen; P.brlen = bb->rlen;
    P.afirst = ab->ufirst; P.bfirst = bb->ufirst;
    P.K = fp.kmer; P.W = fp.binshift; P.H = fp.hitmin; P.minhit = minhit; P.TS = TS; P.minover = c.opt.minover;
    P.comp = comp ? 1 : 0; P.self = self; P.symmetric = c.opt.symmetric;
    P.tandem = tandem ? 1 : 0; P.reach = tandem ? 0 : 1;
    P.ave_path = ave_path; P.dscore = dscore; P.score_tab = dscoret.p; P.table_tab = dtable.p;
    P.tbuf = tbuf.p; P.maxtp = maxtp;
    P.bscore = bscore.p; P.nbuck = nbuck; P.dmin = dmin;
    P.tbytes = tbytes; P.stats = counters.p;
    Q.flag = pflag.p; Q.cand0 = pcand.p; Q.ctr = ctr.p;

    // ---- stage 3: pairs with at least one seed that passes the -h test ----
    filter_pairs<<<sblocks, 256, 0, st>>>(Q);
    KERNEL_CHECK();
    exclusive_scan_u32_to_u64(pflag.p, poff.p, (int64_t)nwork, total.p, st);
    uint64_t nactive = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nactive, total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    count_launch(1);
    c.tm.pairs_ms = t_pairs.stop();
    delete hp1;

    HostProf *hp3 = new HostProf("ra.align_run");
    Timer t_align(st);
    // pair and alignment tables
    const size_t na = (size_t)nactive + 1;
    DevBuf<uint32_t> pt_wslot, pt_flags, ct_pair, list_buf, wide_list;
    DevBuf<int64_t> pt_i64, ct_hidx;
    DevBuf<int32_t> pt_i32, pt_la, ct_i32, la_scratch;
    DevBuf<uint32_t> pt_laoff; DevBuf<unsigned int> la_top;
    size_t la_arena = na * LA_CAP0 + 2 * na + (1u << 16);      // grows (x4) when a round runs out of interval slots
    DevBuf<wc::TaskOut> ct_res;
    pt_wslot.reserve(na); pt_flags.reserve(na); pt_i64.reserve(6 * na); pt_i32.reserve(3 * na); pt_laoff.reserve(2 * na); la_top.reserve(1);
    ct_pair.reserve(na); list_buf.reserve(na); wide_list.reserve(2 * na); ct_hidx.reserve(na); ct_i32.reserve(4 * na); ct_res.reserve(2 * na);
    Q.pt.wslot = pt_wslot.p; Q.pt.flags = pt_flags.p;
    Q.pt.sidx = pt_i64.p; Q.pt.pend = pt_i64.p + na; Q.pt.nidx = pt_i64.p + 2 * na; Q.pt.lidx = pt_i64.p + 3 * na;
    Q.pt.pe = pt_i64.p + 4 * na; Q.pt.j = pt_i64.p + 5 * na;
    Q.pt.amark2 = pt_i32.p; Q.pt.nla = pt_i32.p + na; Q.pt.novl = pt_i32.p + 2 * na;
    Q.pt.la_off = pt_laoff.p; Q.pt.la_cap = pt_laoff.p + na; Q.pt.la_top = la_top.p;
    Q.ct.pair = ct_pair.p; Q.ct.hidx = ct_hidx.p; Q.ct.res = ct_res.p;
    Q.ct.ar = ct_i32.p; Q.ct.br = ct_i32.p + na; Q.ct.apos = ct_i32.p + 2 * na; Q.ct.bpos = ct_i32.p + 3 * na;
    Q.wide_list = wide_list.p;
    DevBuf<wc::Lane> wide_lane; DevBuf<uint32_t> wide_ring, wide_ready;
    wide_lane.reserve(2 * na); wide_ring.reserve(2 * na * 32 * wc::FRONT_WORDS); wide_ready.reserve(2 * na);
    Q.wide_lane = wide_lane.p; Q.wide_ring = wide_ring.p; Q.wide_ready = wide_ready.p; Q.wide_cap = (unsigned int)(2 * na);
    { WaveLog &wl = WaveLog::get(); Q.wlog = wl.buf; Q.wlog_n = wl.n; Q.wlog_cap = wl.cap; }
    DevBuf<uint64_t> okeys; DevBuf<uint32_t> ovals;
    okeys.reserve(4 * na); ovals.reserve(4 * na);
    static const bool lpt = !(getenv("DAMAR_LPT") && atoi(getenv("DAMAR_LPT")) == 0);
    // Two latency options for a host that runs ONE comparison at a time (the drop-in daligner process): deal the longest
    // passes out over all warps (narrow launch 28 -> 25 ms on a C2 block pair) and run the parked passes one per warp all
    // over the device (wide launch 16 -> 10 ms).  With several comparisons in flight both cost throughput (every CTA then
    // holds its SM to the end: C2 step 590 -> 700-750 ms), so they are off unless asked for.
    static const bool spread = getenv("DAMAR_SPREAD") && atoi(getenv("DAMAR_SPREAD")) != 0;
    static const bool wide_spread = getenv("DAMAR_WIDE_SPREAD") && atoi(getenv("DAMAR_WIDE_SPREAD")) != 0;
    static const bool timeline = getenv("DAMAR_TIMELINE") && atoi(getenv("DAMAR_TIMELINE")) != 0;   // per-round kernel times on stderr

    // Pebble pool, sized from the measured need instead of the worst case: a pass logs one cell per trace point it crosses
    // on either axis (~ 2 x length / TS per pass, the alive paths of a front share prefixes) plus up to one partly used
    // 32-cell chunk per thread and pass -- about 300 cells per alignment on 10 kbp reads.  A dry pool is not an error: the
    // alignments that found no cell are retried in the next round (CTR_RETRY), when most pairs are finished.
    const int64_t avglen = std::max<int64_t>(ab->totlen / std::max(1, ab->nreads), bb->totlen / std::max(1, bb->nreads));
    size_t pool_cells = (size_t)nactive * (size_t)(6 * (avglen / TS + 2) + 128) + (1u << 16);
    {
        const size_t worst = (size_t)nactive * (size_t)(2 * (maxlen / TS + 2) + 512) + (1u << 16);
        if (pool_cells > worst) pool_cells = worst;
    }
    if (pool_cells > 0x7ffffff0ull) pool_cells = 0x7ffffff0ull;
    DevBuf<wc::Cell> pool; pool.reserve(pool_cells);
    Q.pool = pool.p; Q.pool_cap = (unsigned int)pool_cells;

    // warps per CTA of the wave kernel (16 KB of shared memory each)
    static const int wwarps = getenv("DAMAR_WAVE_WARPS") ? atoi(getenv("DAMAR_WAVE_WARPS")) : 12;
    // CTAs of the second launch of a round (parked wide passes); 0: one launch, its narrow warps turn wide at the end
    static const int wide_ctas_max = getenv("DAMAR_WIDE_CTAS") ? std::max(0, std::min(64, atoi(getenv("DAMAR_WIDE_CTAS")))) : 8;
    // CTAs (= SMs, one CTA fills an SM's shared memory) of one comparison's wave kernel; fewer than 148 lets the kernels of
    // concurrent comparisons run side by side so that the tail of one overlaps the bulk of another
    static const int wave_ctas = getenv("DAMAR_WAVE_CTAS") ? std::max(1, std::min(NUM_SMS_B200, atoi(getenv("DAMAR_WAVE_CTAS")))) : NUM_SMS_B200;
    static std::once_flag attr;
    std::call_once(attr, [] {
        set_smem(wave_tasks<4, 8>, WaveCfg<8>::SMEM);
        set_smem(wave_tasks<4, 10>, WaveCfg<10>::SMEM);
        set_smem(wave_tasks<4, 12>, WaveCfg<12>::SMEM);
        set_smem(wave_tasks<4, 13>, WaveCfg<13>::SMEM);
        set_smem(wave_tasks<4, 14>, WaveCfg<14>::SMEM);
        set_smem(wave_tasks<4, 1, 12>, 2 * WaveCfg<1>::SMEM);
    });
    // DAMAR_WAVE_CTA_WARPS=1 (experiment): one-warp CTAs, twelve to an SM.  A CTA of all the SM's warps holds the SM until its
    // LAST pass ends (the longest pass of a round runs ~10 ms on one lane after the queue is empty); one-warp CTAs hand
    // their share of the SM to the next kernel as they finish.  Measured on the C2 step: no gain (620-646 ms against 585-620).
    static const int cta_warps = getenv("DAMAR_WAVE_CTA_WARPS") ? atoi(getenv("DAMAR_WAVE_CTA_WARPS")) : 0;

    unsigned int nredo = 0; unsigned long long nrec_fast = 0; int64_t nwide_total = 0;
    DevBuf<uint8_t> dead; DevBuf<uint64_t> hr_keys; DevBuf<uint32_t> hr_vals, hr_pool; DevBuf<unsigned long long> hr_top;
    bool have_dead = false;
    unsigned long long rec_cap = 4 * nactive + 1024;
    unsigned long long trace_cap = (unsigned long long)rec_cap * (unsigned long long)(2 * (maxlen / TS + 2)) / 4 + (1ull << 20);
    DevBuf<damar_overlap> recs; DevBuf<uint64_t> rtoff, rseq; DevBuf<uint8_t> trace;
    unsigned long long hc[8];
    int rounds = 0;
    for (int attempt = 0;; attempt++) {
        recs.reserve((size_t)rec_cap); rtoff.reserve((size_t)rec_cap); rseq.reserve((size_t)rec_cap);
        trace.reserve((size_t)trace_cap * tbytes);
        CUDA_CHECK(cudaMemsetAsync(counters.p, 0, sizeof(unsigned long long) * 8, st));
        CUDA_CHECK(cudaMemsetAsync(next.p, 0, sizeof(unsigned int) * 4, st));
        CUDA_CHECK(cudaMemsetAsync(pair_redo.p, 0, (size_t)nwork + 1, st));
        if (la_arena > 0xfffffff0ull) la_arena = 0xfffffff0ull;
        pt_la.reserve(3 * la_arena);
        Q.pt.la_low = pt_la.p; Q.pt.la_hgh = pt_la.p + la_arena; Q.pt.la_ae = pt_la.p + 2 * la_arena; Q.pt.la_arena = (uint32_t)la_arena;
        { const unsigned int top0 = (unsigned int)(na * LA_CAP0); CUDA_CHECK(cudaMemcpyAsync(la_top.p, &top0, sizeof(top0), cudaMemcpyHostToDevice, st)); }
        P.recs = recs.p; P.rec_toff = rtoff.p; P.rec_seq = rseq.p; P.rec_cap = rec_cap; P.rec_count = counters.p + 5;
        P.trace = trace.p; P.trace_cap = trace_cap; P.trace_count = counters.p + 6;

        if (nactive > 0) {
            init_pairs<<<(unsigned)((nwork + 255) / 256), 256, 0, st>>>(Q, pflag.p, poff.p);
            KERNEL_CHECK(); count_launch();
        }
        Q.list_in = nullptr; Q.n_in = (uint32_t)nactive;
        while (Q.n_in > 0) {
            rounds++;
            CUDA_CHECK(cudaMemsetAsync(ctr.p, 0, sizeof(unsigned int) * CTR_N, st));
            unsigned int ncand = 0;
            if (Q.list_in == nullptr) {                          // round 1: init_pairs filled the candidate table
                ncand = (unsigned int)nactive;
                CUDA_CHECK(cudaMemcpyAsync(ctr.p + CTR_NCAND, &ncand, sizeof(unsigned int), cudaMemcpyHostToDevice, st));
                CUDA_CHECK(stream_wait(st));
            } else {
                const unsigned sb = (unsigned)std::min<uint64_t>((uint64_t)sblocks, ((uint64_t)Q.n_in + 7) / 8);
                scan_round<<<sb, 256, 0, st>>>(Q);
                KERNEL_CHECK(); count_launch();
                CUDA_CHECK(cudaMemcpyAsync(&ncand, ctr.p + CTR_NCAND, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
                CUDA_CHECK(stream_wait(st));
            }
            if (ncand == 0) break;
            Timer t_k(st);
            float t_narrow = 0;
            {
                const bool small_ctas = cta_warps == 1;
                const unsigned per = (unsigned)(small_ctas ? cta_warps : wwarps) * 32;
                const unsigned per_sm = small_ctas ? (unsigned)(wwarps / cta_warps) : 1u;          // CTAs of one SM
                const unsigned wb = std::min<unsigned>((unsigned)wave_ctas * per_sm, (2 * ncand + per - 1) / per);
                const bool two_launches = wb >= 64 * per_sm && wide_ctas_max > 0;
                // small CTAs: shared memory padded so that exactly per_sm of them share an SM
                const int pad_smem = small_ctas ? std::max((int)per / 32 * WaveCfg<1>::SMEM, (int)((216 * 1024) / per_sm) - 2048) : 0;
                Q.task_order = nullptr;
                if (ncand >= 2048 && lpt) {                      // order the queue: longest passes first
                    const int64_t nt = 2 * (int64_t)ncand;
                    uint64_t *k0 = okeys.p, *k1 = okeys.p + 2 * na; uint32_t *v0 = ovals.p, *v1 = ovals.p + 2 * na;
                    task_length_keys<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(Q, (unsigned)nt, k0, v0);
                    KERNEL_CHECK(); count_launch();
                    radix_sort_pairs(k0, v0, k1, v1, nt, 24, st);
                    Q.task_order = v0;
                    if (spread) {
                        const unsigned chunks = (unsigned)std::min<int64_t>(nt / 32, (int64_t)wb * (per / 32));
                        spread_order<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(v0, v1, (unsigned)nt, chunks);
                        KERNEL_CHECK(); count_launch();
                        Q.task_order = v1;
                    }
                }
                CUDA_CHECK(cudaMemsetAsync(wide_ready.p, 0, sizeof(uint32_t) * 2 * na, st));   // all of it: tickets run past this round's parks
                auto launch = [&](unsigned grid) {
                    if (small_ctas) wave_tasks<4, 1, 12><<<grid, per, pad_smem, st>>>(Q);
                    else if (wwarps == 8) wave_tasks<4, 8><<<grid, per, WaveCfg<8>::SMEM, st>>>(Q);
                    else if (wwarps == 10) wave_tasks<4, 10><<<grid, per, WaveCfg<10>::SMEM, st>>>(Q);
                    else if (wwarps == 13) wave_tasks<4, 13><<<grid, per, WaveCfg<13>::SMEM, st>>>(Q);
                    else if (wwarps == 14) wave_tasks<4, 14><<<grid, per, WaveCfg<14>::SMEM, st>>>(Q);
                    else wave_tasks<4, 12><<<grid, per, WaveCfg<12>::SMEM, st>>>(Q);
                    KERNEL_CHECK(); count_launch();
                };
                Q.wide_ctas = 0; Q.linger = two_launches ? 0 : 1;
                launch(wb);
                if (timeline) t_narrow = t_k.lap();
                if (two_launches && wide_spread) {            // one-warp CTAs all over the device, as few parked passes per warp as possible
                    const unsigned grid = (unsigned)wave_ctas * 12u;
                    Q.wide_ctas = (int)grid;
                    wave_tasks<4, 1, 12><<<grid, 32, WaveCfg<1>::SMEM, st>>>(Q);
                    KERNEL_CHECK(); count_launch();
                } else if (two_launches) { Q.wide_ctas = wide_ctas_max * (int)per_sm; launch((unsigned)wide_ctas_max * per_sm); }
            }
            const float t_round = t_k.stop();
            c.tm.align_kernel_ms += t_round; c.tm.align_launches += 1;
            { unsigned int nw = 0; CUDA_CHECK(cudaMemcpyAsync(&nw, ctr.p + CTR_NWIDE, sizeof(nw), cudaMemcpyDeviceToHost, st)); CUDA_CHECK(stream_wait(st)); nwide_total += nw;
              if (timeline) fprintf(stderr, "[timeline] round %d: %u candidates, narrow %.3f ms, wide %.3f ms (%u parked)\n", rounds, ncand, t_narrow, t_round - t_narrow, nw); }
            const unsigned fb = (unsigned)std::min<uint64_t>((uint64_t)sblocks, ((uint64_t)ncand + 7) / 8);
            finish_las<<<fb, 256, 0, st>>>(Q);
            KERNEL_CHECK(); count_launch();
            CUDA_CHECK(cudaMemcpyAsync(list_buf.p, ct_pair.p, sizeof(uint32_t) * ncand, cudaMemcpyDeviceToDevice, st));
            Q.list_in = list_buf.p; Q.n_in = ncand;
        }
        CUDA_CHECK(cudaMemcpyAsync(&nredo, next.p + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(&nrec_fast, counters.p + 5, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_wait(st));
        if (nredo > 0) {                               // strided warp kernel for the pairs with very wide fronts
            DevBuf<int32_t> V, HA, HB, HM, HN, NA, NB; DevBuf<uint64_t> T; DevBuf<int4> cells;
            const int lw = lblocks * AL_WARPS;
            const int cells_cap = 4 * (2 * (maxlen / TS + 2)) * 24 + 4096;
            V.reserve((size_t)lw * 2 * ring); HA.reserve((size_t)lw * 2 * ring); HB.reserve((size_t)lw * 2 * ring);
            HM.reserve((size_t)lw * 2 * ring); HN.reserve((size_t)lw * 2 * ring); T.reserve((size_t)lw * 2 * ring);
            NA.reserve((size_t)lw * ring); NB.reserve((size_t)lw * ring); cells.reserve((size_t)lw * cells_cap);
            P.ring = ring; P.V = V.p; P.HA = HA.p; P.HB = HB.p; P.HM = HM.p; P.HN = HN.p; P.NA = NA.p; P.NB = NB.p; P.T = T.p;
            P.cells = cells.p; P.cells_cap = cells_cap;
            la_scratch.reserve((size_t)lw * 3 * (MAX_LA_PER_PAIR + 1));
            P.la_scratch = la_scratch.p;
            CUDA_CHECK(cudaMemsetAsync(next.p, 0, sizeof(unsigned int), st));
            AlignArgs R = P;
            R.work = redo_sidx.p; R.work_wi = redo_wi.p; R.nwork = nredo;
            Timer t_r(st);
            align_pairs_strided<<<lblocks, AL_WARPS * 32, 0, st>>>(R);
            KERNEL_CHECK(); count_launch();
            const float t_redo = t_r.stop();
            c.tm.redo_kernel_ms += t_redo;
            if (timeline) fprintf(stderr, "[timeline] strided kernel: %u pairs, %.3f ms\n", nredo, t_redo);
        }
        CUDA_CHECK(cudaMemcpyAsync(hc, counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_wait(st));
        have_dead = false;
        if (!(hc[3] & 1ull) && hc[5] > 0 && (hc[1] > 0 || nredo > 0)) {
            // records the task pipeline emitted for pairs it later handed to the strided kernel are void; pairs with several
            // alignments go through Handle_Redundancies (filter.c:2438-2449), one thread per pair
            HostProf hpf("ra.fuse");
            const unsigned long long nrec = hc[5];
            dead.reserve((size_t)nrec + 1);
            hr_keys.reserve(2 * (size_t)nrec + 4); hr_vals.reserve(2 * (size_t)nrec + 4);
            HrArgs H{};
            H.recs = recs.p; H.toff = rtoff.p; H.seq = rseq.p; H.dead = dead.p; H.nrec = nrec; H.nrec_fast = nrec_fast;
            H.pair_redo = nredo > 0 ? pair_redo.p : nullptr;
            H.trace = trace.p; H.tbytes = tbytes; H.trace_cap = trace_cap; H.trace_count = counters.p + 6;
            H.aw = ab->packed[0]; H.bw = bb->packed[bsel]; H.aboff = ab->boff; H.bboff = bb->boff; H.arlen = ab->rlen; H.brlen = bb->rlen;
            H.afirst = ab->ufirst; H.bfirst = bb->ufirst; H.comp = comp ? 1 : 0; H.ts = TS; H.bridge = tandem ? 0 : 1;
            H.stats = counters.p; H.nfused = counters.p + 7;
            uint64_t *k0 = hr_keys.p, *k1 = hr_keys.p + nrec + 2; uint32_t *v0 = hr_vals.p, *v1 = hr_vals.p + nrec + 2;
            hr_sort_keys<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(H, k0, v0);
            KERNEL_CHECK(); count_launch();
            have_dead = true;
            if (hc[1] > 0) {
                static const size_t pool_mb = getenv("DAMAR_HR_POOL_MB") ? (size_t)atoll(getenv("DAMAR_HR_POOL_MB")) : 0;
                size_t pool_words = pool_mb ? (pool_mb << 18) : std::min<size_t>((size_t)2 << 30, ((size_t)hc[1] << 16) + ((size_t)16 << 20));
                hr_pool.reserve(pool_words); hr_top.reserve(1);
                CUDA_CHECK(cudaMemsetAsync(hr_top.p, 0, sizeof(unsigned long long), st));
                int kb = SEQ_PAIR_SHIFT + 1;                                     // rank = pair | B flag | alignment number
                while ((1ull << (kb - SEQ_PAIR_SHIFT)) < (unsigned long long)nwork + 2) kb++;
                radix_sort_pairs(k0, v0, k1, v1, (int64_t)nrec, kb, st);
                H.order = v0; H.pool = hr_pool.p; H.pool_cap = pool_words; H.pool_top = hr_top.p;
                handle_redundancies_pairs<<<(unsigned)((nrec + 63) / 64), 64, 0, st>>>(H);
                KERNEL_CHECK(); count_launch();
            }
            CUDA_CHECK(cudaMemcpyAsync(hc, counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(stream_wait(st));
        }
        if ((hc[3] & (1ull | 32ull)) && attempt < 6) {  // output buffers or the interval arena too small: grow and rerun
            if (hc[5] > rec_cap) rec_cap = hc[5] + 1024;
            if (hc[6] > trace_cap) trace_cap = hc[6] + 1024;
            if (hc[3] & 32ull) la_arena *= 4;
            CUDA_CHECK(cudaMemsetAsync(bscore.p, 0, sizeof(int32_t) * (size_t)nwarps * 2 * nbuck, st));
            continue;
        }
        break;
    }
    c.tm.align_ms = t_align.stop();
    if (timeline) fprintf(stderr, "[timeline] stage 4 total %.3f ms, %d rounds\n", c.tm.align_ms, rounds);
    delete hp3;
    // capacity errors are reported to the caller (damar_match_filter returns non-zero), never a silent truncation
    if (hc[3] & 4ull) { set_error("alignment scratch overflow: wave wider than %d diagonals or pebble arena exhausted", ring); return 1; }
    if (hc[3] & 2ull) { set_error("a read pair has more than %d local alignments in one orientation", MAX_LA_PER_PAIR); return 1; }
    if (hc[3] & (1ull | 32ull)) { set_error("alignment output buffers overflowed repeatedly"); return 1; }
    if (hc[3] & 16ull) { set_error("Handle_Redundancies scratch pool exhausted; raise DAMAR_HR_POOL_MB"); return 1; }
    if (hc[3] & 8ull) { set_error("Compression of trace to bytes fails, value too big"); return 1; }   // align.c:3385-3388

    if (pair_base && hc[5] > 0) {                     // a later chunk of a comparison: its pairs rank after those of the chunks before
        add_seq_base<<<(unsigned)((hc[5] + 255) / 256), 256, 0, st>>>(rseq.p, hc[5], pair_base << SEQ_PAIR_SHIFT);
        KERNEL_CHECK(); count_launch();
        CUDA_CHECK(stream_wait(st));                   // the records may be picked up by another worker's stream
    }
    // ---- download into pinned staging; records stay in device emission order, `seq` carries the
    //      order the reference's thread buffers would have had ----
    if (out.keep != nullptr) {                         // the device-side writer takes the buffers as they are (las_device.cu)
        DeviceResult &K = *out.keep;
        const uint64_t nrec = hc[5];
        K.recs.take(recs); K.toff.take(rtoff); K.seq.take(rseq); K.trace.take(trace);
        K.nrec = nrec; K.ntrace = hc[6]; K.tbytes = tbytes; K.have_dead = have_dead && nrec > 0;
        uint64_t nkeep = nrec;
        if (K.have_dead) {
            K.dead.take(dead);
            std::vector<uint8_t> h_dead(nrec);
            CUDA_CHECK(cudaMemcpyAsync(h_dead.data(), K.dead.p, nrec, cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(stream_wait(st));
            for (uint64_t i = 0; i < nrec; i++) nkeep -= h_dead[i] ? 1 : 0;
        }
        out.ovl = nullptr; out.toff = nullptr; out.seq = nullptr; out.trace = nullptr;
        out.novl = (int64_t)nkeep; out.ntrace = (int64_t)hc[6]; out.tbytes = tbytes; out.redo = (int64_t)nredo;
        c.tm.wide_passes = nwide_total;
        out.nfilt = (int64_t)hc[0]; out.multi = (int64_t)hc[1]; out.cells = (int64_t)hc[2]; out.pairs = (int64_t)nwork;
        c.tm.download_ms = 0;
        return 0;
    }
    HostProf hp4("ra.download");
    Timer t_dl(st);
    WorkerScratch &ws = scratch();
    PinnedBuf<damar_overlap> &h_recs = ws.h_recs; PinnedBuf<uint64_t> &h_toff = ws.h_toff, &h_seq = ws.h_seq; PinnedBuf<uint8_t> &h_trace = ws.h_trace;
    const uint64_t nrec = hc[5], ntr = hc[6];
    h_recs.reserve(nrec + 1); h_toff.reserve(nrec + 1); h_seq.reserve(nrec + 1); h_trace.reserve(ntr * tbytes + 16);
    if (nrec) {
        CUDA_CHECK(cudaMemcpyAsync(h_recs.p, recs.p, sizeof(damar_overlap) * nrec, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(h_toff.p, rtoff.p, sizeof(uint64_t) * nrec, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(h_seq.p, rseq.p, sizeof(uint64_t) * nrec, cudaMemcpyDeviceToHost, st));
    }
    if (ntr) CUDA_CHECK(cudaMemcpyAsync(h_trace.p, trace.p, (size_t)ntr * tbytes, cudaMemcpyDeviceToHost, st));
    std::vector<uint8_t> h_dead;
    if (have_dead && nrec) { h_dead.resize(nrec); CUDA_CHECK(cudaMemcpyAsync(h_dead.data(), dead.p, nrec, cudaMemcpyDeviceToHost, st)); }
    CUDA_CHECK(stream_wait(st));
    uint64_t nkeep = nrec;
    if (have_dead) {                                   // drop the flagged records (pure bookkeeping; the decisions were made on the device)
        nkeep = 0;
        for (uint64_t i = 0; i < nrec; i++) {
            if (h_dead[i]) continue;
            if (nkeep != i) { h_recs.p[nkeep] = h_recs.p[i]; h_toff.p[nkeep] = h_toff.p[i]; h_seq.p[nkeep] = h_seq.p[i]; }
            nkeep++;
        }
    }
    const uint64_t ntr_out = ntr;
    const void *trace_out = h_trace.p;
    out.ovl = h_recs.p; out.toff = h_toff.p; out.seq = h_seq.p; out.trace = trace_out;
    out.novl = (int64_t)nkeep; out.ntrace = (int64_t)ntr_out; out.tbytes = tbytes; out.redo = (int64_t)nredo;
    c.tm.wide_passes = nwide_total;
    out.nfilt = (int64_t)hc[0]; out.multi = (int64_t)hc[1]; out.cells = (int64_t)hc[2]; out.pairs = (int64_t)nwork;
    c.tm.download_ms = t_dl.stop();
    (void)rounds;
    return 0;
}

}  // namespace damar
