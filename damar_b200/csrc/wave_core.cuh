// wave_core.cuh -- the O(ND) wave extension of Local_Alignment as *scalar* code: one task (one
// forward or one reverse pass of one local alignment) per thread.
//
// Replaces forward_wave / reverse_wave (reference dalign/align.c:409-1122, 1126-1898).  Both passes
// run through ONE code path: the reverse pass is a forward pass over the reverse-complemented reads
// in mirrored coordinates  xt = alen - x,  yt = blen - y  (so  kt = (alen - blen) - k,
// ct = alen + blen - c).  The reference's reverse_wave is the exact mirror image of forward_wave
// (minima for maxima, visiting order low->hgh for hgh->low, `<=` trace tests for `>=`, clip roles
// kept), so every tie-break of align.c:1499-1535/1620-1684 falls out of the forward rules
// (align.c:793-825/911-975) applied to the mirrored problem; comparing complemented bases is the
// same as comparing bases.  Only wave 0's root pebbles (align.c:1230-1249) and the trace emission
// differ, and those are handled by the `rev` flag.
//
// A thread sweeps its wave front hgh -> low exactly like the reference's inner loop, carrying the
// previous wave's values of diagonal k+1 in registers, so the reference's sequential semantics
// (best / trim / clip bookkeeping in visiting order) need no reformulation.  The front lives in a
// ring of FRONT_RING diagonals per thread ([slot][thread] interleaved in shared memory on the device:
// every thread keeps to its own bank).  Trace-point pebbles go to a global pool in 16-byte cells.
//
// Everything here is __host__ __device__: tests/wave_core_host.cu runs the very same code on the
// CPU against the oracle (test infrastructure); the product only ever calls it from kernels.
#pragma once
#include <cstdint>
#ifdef WC_DEBUG
#include <cstdio>
#endif

#if defined(__CUDACC__)
#define WC_HD __host__ __device__ __forceinline__
#else
#define WC_HD inline
#endif

namespace damar {
namespace wc {

constexpr int TRIM_LEN = 15;           // align.c:188
constexpr int PATH_LEN = 60;           // align.c:192
constexpr uint64_t PATH_INT = 0x0fffffffffffffffull;
constexpr uint64_t PATH_MASK61 = 0x1fffffffffffffffull;   // M = popcount of bits 0..60 (align.c:827-855)
constexpr int TRIM_MASK = 0x7fff;
constexpr int TRIM_MLAG = 200;
constexpr int WAVE_LAG = 30;
constexpr int IMAX = 0x7fffffff;
constexpr uint32_t NOCELL = 0xffffffffu;
constexpr int CELL_CHUNK = 32;         // pebble cells are taken from the pool in chunks of this many

enum Status : int { ST_IDLE = 0, ST_RUN = 1, ST_DONE = 2, ST_WIDE = 3, ST_POOL = 4, ST_RANGE = 5, ST_EXIT = 6 };

// ---- bit helpers -------------------------------------------------------------------------------
WC_HD int clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
WC_HD int popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// upper 32 bits of (hi:lo) << s, 0 <= s < 32
WC_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, int s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    return s ? ((hi << s) | (lo >> (32 - s))) : hi;
#endif
}
WC_HD uint32_t ldg32(const uint32_t *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
WC_HD int16_t ldg16(const int16_t *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// 32 bases (first base most significant) starting at base index s of a packed array
WC_HD uint64_t window_at(const uint32_t *__restrict__ words, uint32_t s) {
    const uint32_t *w = words + (s >> 4);
    const int o = (int)(s & 15u) * 2;
    const uint32_t x0 = ldg32(w), x1 = ldg32(w + 1), x2 = ldg32(w + 2);
    return ((uint64_t)funnel_l(x1, x0, o) << 32) | funnel_l(x2, x1, o);
}
// number of equal bases A[ax+i] == B[by+i], i = 0.., capped at maxn
WC_HD int slide(const uint32_t *__restrict__ aw, uint32_t ax, const uint32_t *__restrict__ bw, uint32_t by, int maxn) {
    int n = 0;
    while (n < maxn) {
        const uint64_t d = window_at(aw, ax + (uint32_t)n) ^ window_at(bw, by + (uint32_t)n);
        if (d != 0) { n += clz64(d) >> 1; break; }
        n += 32;
    }
    return n < maxn ? n : maxn;
}
WC_HD uint64_t push_matches(uint64_t b, int n) {
    return (n >= 64) ? ~0ull : ((b << n) | ((1ull << n) - 1ull));
}

// ---- pebble cell (align.c:399-405): {prev, diag, diff, mark}, original coordinates ----------------
struct Cell { int32_t prev, diag, diff, mark; };

// chunked bump allocation from the global pool
struct CellPool {
    Cell *cells;
    unsigned int *top;          // next free cell (device: atomic)
    unsigned int cap;
};

// ---- per-thread view of its wave-front ring ------------------------------------------------------
// word arrays: V (furthest anti-diagonal), T lo/hi (match history), HA, HB (chain heads, pool index),
// HMN (TP index+1 of the heads' marks, A low 16 | B high 16), NAB (next TP index+1 to cross, A | B)
template <int RING, int STRIDE>
struct Front {
    uint32_t *base;
    WC_HD uint32_t &at(int arr, int k) const { return base[(arr * RING + (k & (RING - 1))) * STRIDE]; }
    WC_HD int32_t getV(int k) const { return (int32_t)at(0, k); }
    WC_HD void setV(int k, int32_t v) const { at(0, k) = (uint32_t)v; }
};
constexpr int FRONT_WORDS = 7;

// ---- one task -----------------------------------------------------------------------------------
struct TaskIn {
    const uint32_t *aw, *bw;        // packed reads of this pass (forward or reverse-complemented arrays)
    uint32_t abase, bbase;          // base index of the two reads inside aw / bw
    int alen, blen;
    int diag, anti;                 // seed in ORIGINAL coordinates
    int aoff, boff;                 // trace-point offsets in original coordinates (align.c:1970-1984)
    int minp, maxp;                 // original coordinates (align.c:1951-1968)
    int rev;                        // 0 = forward pass, 1 = reverse pass
    int TS, path_ave;
    const int16_t *score_tab, *table_tab;
};
struct TaskOut {                    // original coordinates
    int32_t ex, ey, ediff;
    int32_t ha, hb;                 // heads of the chosen A / B pebble chains
    int32_t status;
    int32_t waves, cells;           // work counters (wave cells after pruning, as the reference's WAVE_STATS)
};

struct Lane {
    // task constants
    const uint32_t *aw, *bw;
    uint32_t abase, bbase;
    int alen, blen, minp, maxp, TS, path_ave, rev;
    int goffA, goffB;               // trace grid offsets in pass coordinates
    const int16_t *score_tab, *table_tab;
    // wave state (pass coordinates)
    int low, hgh, k, dif, more, aclip, bclip;
    int besta, besty, lasta, trima, trimy, trimd;
    int morea, morey, mored, morem;
    uint32_t trimha, trimhb, moreha, morehb;
    // sweep registers: previous-wave values around k
    int ap, ac, am;                 // V[k+1], V[k], V[k-1] of the previous wave (after the shift at the top of a cell)
    uint64_t pT, cT;                // T of k+1 / k (previous wave)
    uint32_t pHA, pHB, pHMN, cHA, cHB, cHMN;
    // pebble allocation
    uint32_t cell_next, cell_left;
    int status;
    uint32_t waves, cells;
};

WC_HD int grid(int goff, int idx1, int TS) { return goff + (idx1 - 1) * TS; }   // idx1 = TP index + 1

template <class POOL>
WC_HD uint32_t push_cell(Lane &L, POOL &pool, uint32_t prev, int k, int dif, int markA_or_B_pass, bool isA) {
    if (L.cell_left == 0) {
        const uint32_t b = pool.take(CELL_CHUNK);
        if (b == NOCELL) { L.status = ST_POOL; return prev; }
        L.cell_next = b; L.cell_left = CELL_CHUNK;
    }
    const uint32_t id = L.cell_next++;
    L.cell_left--;
    Cell c;
    c.prev = (int32_t)prev;
    c.diff = dif;
    if (L.rev) {                       // back to original coordinates
        c.diag = (L.alen - L.blen) - k;
        c.mark = (isA ? L.alen : L.blen) - markA_or_B_pass;
    } else { c.diag = k; c.mark = markA_or_B_pass; }
    pool.store(id, c);
    return id;
}

// floor((v - goff) / TS) + 1 for v - goff >= -TS
WC_HD int tp_index1(int v, int goff, int TS) { return (v - goff + TS) / TS; }

// Wave 0 (align.c:491-626 forward, 1206-1340 reverse) and the lane's constants.
template <class FRONT, class POOL>
WC_HD void setup(Lane &L, const TaskIn &t, const FRONT &F, POOL &pool) {
    L.aw = t.aw; L.bw = t.bw; L.abase = t.abase; L.bbase = t.bbase;
    L.alen = t.alen; L.blen = t.blen; L.TS = t.TS; L.path_ave = t.path_ave; L.rev = t.rev;
    L.score_tab = t.score_tab; L.table_tab = t.table_tab;
    L.cell_left = 0; L.cell_next = 0;
    L.waves = 0; L.cells = 0;
    const int TS = t.TS;
    const int y0 = (t.anti - t.diag) >> 1, x0 = y0 + t.diag;
    int x, y, k;
    if (!t.rev) {
        x = x0; y = y0; k = t.diag;
        L.goffA = t.aoff; L.goffB = t.boff;
        L.minp = t.minp; L.maxp = t.maxp;
    } else {
        const int D = t.alen - t.blen;
        x = t.alen - x0; y = t.blen - y0; k = x - y;
        L.goffA = (t.alen - t.aoff) % TS; L.goffB = (t.blen - t.boff) % TS;
        L.minp = (t.maxp == IMAX) ? -IMAX : D - t.maxp;
        L.maxp = (t.minp == -IMAX) ? IMAX : D - t.minp;
    }
    L.low = L.hgh = k; L.dif = 0; L.more = 1; L.aclip = IMAX; L.bclip = -IMAX;
    const int anti = x + y;
    L.besta = L.trima = L.morea = L.lasta = anti;
    L.besty = L.trimy = L.morey = y;
    L.trimd = L.mored = 0; L.morem = -1;
    L.status = ST_RUN;

    // root pebbles: forward marks the trace point at or below the seed, reverse the seed itself
    int naI = tp_index1(x, L.goffA, TS), nbI = tp_index1(y, L.goffB, TS);     // index+1 of floor grid point
    uint32_t ha = push_cell(L, pool, NOCELL, k, 0, t.rev ? x : grid(L.goffA, naI, TS), true);
    uint32_t hb = push_cell(L, pool, NOCELL, k, 0, t.rev ? y : grid(L.goffB, nbI, TS), false);
    L.trimha = L.moreha = ha; L.trimhb = L.morehb = hb;
    naI += 1; nbI += 1;
    if (naI > 0xfffe || nbI > 0xfffe) { L.status = ST_RANGE; return; }

    {
        const int ra = L.alen - x, rb = L.blen - y;
        const int n = slide(L.aw, L.abase + (uint32_t)x, L.bw, L.bbase + (uint32_t)y, ra < rb ? ra : rb);
        x += n; y += n;
    }
    if (y == L.blen) { L.more = 0; if (L.bclip < k) L.bclip = k; }
    else if (x == L.alen) { L.more = 0; L.aclip = k; }
    while (x >= grid(L.goffA, naI, TS)) { ha = push_cell(L, pool, ha, k, 0, grid(L.goffA, naI, TS), true); naI++; }
    while (y >= grid(L.goffB, nbI, TS)) { hb = push_cell(L, pool, hb, k, 0, grid(L.goffB, nbI, TS), false); nbI++; }
    if (naI > 0xfffe || nbI > 0xfffe) { if (L.status == ST_RUN) L.status = ST_RANGE; return; }
    const int c = (y << 1) + k;
    if (c > L.besta) { L.besta = L.trima = L.lasta = c; L.besty = L.trimy = y; L.trimha = ha; L.trimhb = hb; }
    F.setV(k, c);
    F.at(1, k) = (uint32_t)PATH_INT; F.at(2, k) = (uint32_t)(PATH_INT >> 32);
    F.at(3, k) = ha; F.at(4, k) = hb;
    F.at(5, k) = (uint32_t)(naI - 1) | ((uint32_t)(nbI - 1) << 16);
    F.at(6, k) = (uint32_t)naI | ((uint32_t)nbI << 16);
    L.k = k - 1;                                  // wave 0 is complete: next step is a transition
}

// End of a wave: boundary clipping (align.c:628-658, 943-975), pruning (:977-986), the loop test
// (:667) and the widening of the next wave (:756-772).  Leaves the lane either finished
// (status != ST_RUN) or positioned on k = hgh of the new wave.
template <int RING, class FRONT>
WC_HD void transition(Lane &L, const FRONT &F) {
    if (L.more == 0) {
        if (L.besty < L.blen && L.besta - L.besty < L.alen) L.more = 1;
        if (L.hgh >= L.aclip) {
            const int kc = L.aclip;
            L.hgh = kc - 1;
            const uint64_t tc = (uint64_t)F.at(1, kc) | ((uint64_t)F.at(2, kc) << 32);
            const int mc = popc64(tc & PATH_MASK61);
            if (L.morem <= mc) {
                L.morem = mc; L.morea = F.getV(kc); L.morey = (L.morea - kc) / 2; L.mored = L.dif;
                L.moreha = F.at(3, kc); L.morehb = F.at(4, kc);
            }
        }
        if (L.low <= L.bclip) {
            const int kc = L.bclip;
            L.low = kc + 1;
            const uint64_t tc = (uint64_t)F.at(1, kc) | ((uint64_t)F.at(2, kc) << 32);
            const int mc = popc64(tc & PATH_MASK61);
            if (L.morem <= mc) {
                L.morem = mc; L.morea = F.getV(kc); L.morey = (L.morea - kc) / 2; L.mored = L.dif;
                L.moreha = F.at(3, kc); L.morehb = F.at(4, kc);
            }
        }
        L.aclip = IMAX; L.bclip = -IMAX;
    }
    if (L.dif > 0) {
        const int n = L.besta - WAVE_LAG;
        while (L.hgh >= L.low)
            if (F.getV(L.hgh) < n) L.hgh -= 1;
            else { while (F.getV(L.low) < n) L.low += 1; break; }
        L.waves += 1;
        if (L.hgh >= L.low) L.cells += (uint32_t)(L.hgh - L.low + 1);
    }
    if (!(L.more && L.lasta >= L.besta - TRIM_MLAG) || L.hgh < L.low) { L.status = ST_DONE; return; }
    // an empty front (every diagonal ran off a read end) ends the pass; the reference reads stale memory there

    int low = L.low - 1, hgh = L.hgh + 1;
#ifdef WC_DEBUG
    printf("dif %d low %d hgh %d besta %d lasta %d more %d\n", L.dif, L.low, L.hgh, L.besta, L.lasta, L.more);
#endif
    if (hgh - low + 1 > RING) { L.status = ST_WIDE; return; }
    if (low >= L.minp) { F.at(6, low) = F.at(6, low + 1); F.setV(low, -1); } else low += 1;
    if (hgh <= L.maxp) { F.at(6, hgh) = F.at(6, hgh - 1); F.setV(hgh, -1); L.am = -1; } else { hgh -= 1; L.am = F.getV(hgh); }
    L.low = low; L.hgh = hgh;
    L.dif += 1;
    L.ac = -1;
    L.cT = PATH_INT; L.cHA = L.cHB = NOCELL; L.cHMN = 0;      // becomes the (never chosen) k+1 state of the first cell
    L.k = hgh;
}

// One furthest-reaching-point update on diagonal L.k (align.c:789-938).
template <class FRONT, class POOL>
WC_HD void cell(Lane &L, const FRONT &F, POOL &pool) {
    const int k = L.k, TS = L.TS;
    // shift the sweep window: (ap, ac, am) <- (ac, am, V[k-1]); path state likewise
    L.ap = L.ac; L.ac = L.am; L.am = (k - 1 >= L.low) ? F.getV(k - 1) : -1;
    L.pT = L.cT; L.pHA = L.cHA; L.pHB = L.cHB; L.pHMN = L.cHMN;
    L.cT = (uint64_t)F.at(1, k) | ((uint64_t)F.at(2, k) << 32); L.cHA = F.at(3, k); L.cHB = F.at(4, k); L.cHMN = F.at(5, k);
    const int ap = L.ap, ac = L.ac, am = L.am;
    int c; uint64_t b; uint32_t ha, hb, hmn;
    if (ac < am) {
        if (am < ap) { c = ap + 1; b = L.pT; ha = L.pHA; hb = L.pHB; hmn = L.pHMN; }
        else {
            c = am + 1;
            b = (uint64_t)F.at(1, k - 1) | ((uint64_t)F.at(2, k - 1) << 32); ha = F.at(3, k - 1); hb = F.at(4, k - 1); hmn = F.at(5, k - 1);
        }
    } else if (ac < ap) { c = ap + 1; b = L.pT; ha = L.pHA; hb = L.pHB; hmn = L.pHMN; }
    else { c = ac + 2; b = L.cT; ha = L.cHA; hb = L.cHB; hmn = L.cHMN; }

    b <<= 1;
    int y = (c - k) >> 1, x = y + k;
    {
        const int ra = L.alen - x, rb = L.blen - y;
        const int n = slide(L.aw, L.abase + (uint32_t)x, L.bw, L.bbase + (uint32_t)y, ra < rb ? ra : rb);
        x += n; y += n;
        b = push_matches(b, n);
    }
    if (y == L.blen) { L.more = 0; if (L.bclip < k) L.bclip = k; }
    else if (x == L.alen) { L.more = 0; L.aclip = k; }
    c = (y << 1) + k;
    const int m = popc64(b & PATH_MASK61);

    const uint32_t nab = F.at(6, k);
    int naI = (int)(nab & 0xffffu), nbI = (int)(nab >> 16);
    int hmI = (int)(hmn & 0xffffu), hnI = (int)(hmn >> 16);
    if (x >= grid(L.goffA, naI, TS) || y >= grid(L.goffB, nbI, TS)) {      // align.c:859-909
        while (x >= grid(L.goffA, naI, TS)) {
            if (hmI < naI) { ha = push_cell(L, pool, ha, k, L.dif, grid(L.goffA, naI, TS), true); hmI = naI; }
            naI++;
        }
        while (y >= grid(L.goffB, nbI, TS)) {
            if (hnI < nbI) { hb = push_cell(L, pool, hb, k, L.dif, grid(L.goffB, nbI, TS), false); hnI = nbI; }
            nbI++;
        }
        if (naI > 0xfffe || nbI > 0xfffe) { if (L.status == ST_RUN) L.status = ST_RANGE; }
    }
    if (c > L.besta) {                                                       // align.c:911-928
        L.besta = c; L.besty = y;
        if (m >= L.path_ave) {
            L.lasta = c;
            const int lo15 = (int)(b & TRIM_MASK), hi15 = (int)((b >> TRIM_LEN) & TRIM_MASK);
            if (ldg16(L.table_tab + lo15) >= 0 && ldg16(L.table_tab + hi15) + ldg16(L.score_tab + lo15) >= 0) {
                L.trima = c; L.trimy = y; L.trimd = L.dif; L.trimha = ha; L.trimhb = hb;
            }
        }
    }
    F.setV(k, c);
    F.at(1, k) = (uint32_t)b; F.at(2, k) = (uint32_t)(b >> 32);
    F.at(3, k) = ha; F.at(4, k) = hb;
    F.at(5, k) = (uint32_t)hmI | ((uint32_t)hnI << 16);
    F.at(6, k) = (uint32_t)naI | ((uint32_t)nbI << 16);
    L.k = k - 1;
}

// end point of the pass in original coordinates (REACH = 1, daligner.c:951; align.c:1001-1012)
WC_HD void finish(const Lane &L, TaskOut &o) {
    int ex, ey;
    if (L.morem >= 0) { ex = L.morea - L.morey; ey = L.morey; o.ediff = L.mored; o.ha = (int32_t)L.moreha; o.hb = (int32_t)L.morehb; }
    else { ex = L.trima - L.trimy; ey = L.trimy; o.ediff = L.trimd; o.ha = (int32_t)L.trimha; o.hb = (int32_t)L.trimhb; }
    if (L.rev) { ex = L.alen - ex; ey = L.blen - ey; }
    o.ex = ex; o.ey = ey;
    o.status = L.status;
    o.waves = (int32_t)L.waves; o.cells = (int32_t)L.cells;
}

// ---- trace emission from the pebble chains (align.c:1001-1111 forward, 1711-1880 reverse) --------
// `cells` is the pool; chains end with prev = -1.  Forward: pairs are appended at tr[0..); returns the
// number of values written.
WC_HD int emit_chain_fwd(const Cell *cells, int head, bool isA, int anti, int ex, int ey, int ediff, uint16_t *tr) {
    int n = 0;
    for (int h = head; h >= 0; h = cells[h].prev) n++;
    // cells in forward order are c_0 (root) .. c_{n-1} (head); pair i (i>=1) = (diff_i - diff_{i-1}, p_i - p_{i-1})
    int len = 2 * (n - 1);
    Cell cur = cells[head];
    const int lastk = cur.diag;
    int lastp, laste;
    for (int i = n - 1; i >= 1; i--) {
        const Cell prv = cells[cur.prev];
        const int pc = isA ? (cur.mark - cur.diag) : (cur.mark + cur.diag);
        int pp, pe;
        if (i == 1) { pp = isA ? ((anti - prv.diag) / 2) : ((anti + prv.diag) / 2); pe = 0; }
        else { pp = isA ? (prv.mark - prv.diag) : (prv.mark + prv.diag); pe = prv.diff; }
        tr[2 * (i - 1)] = (uint16_t)(cur.diff - pe);
        tr[2 * (i - 1) + 1] = (uint16_t)(pc - pp);
        cur = prv;
    }
    {
        const Cell hc = cells[head];
        if (n == 1) { lastp = isA ? ((anti - hc.diag) / 2) : ((anti + hc.diag) / 2); laste = 0; }
        else { lastp = isA ? (hc.mark - hc.diag) : (hc.mark + hc.diag); laste = hc.diff; }
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

// Reverse pass: `flen` = number of values the forward pass left at tr[0..); pairs are prepended before
// tr[0] and the first partial segment is merged into tr[0], tr[1].  Returns the number of values prepended.
WC_HD int emit_chain_rev(Cell *cells, int head, bool isA, int ex, int ey, int ediff, int TS, int off, int flen, uint16_t *tr) {
    int a = -1, b;
    for (int h = head; h >= 0; h = b) { b = cells[h].prev; cells[h].prev = a; a = h; }   // reverse the links in place
    int h = a;                                       // root
    Cell c = cells[h];
    int k = c.diag;
    int pb = isA ? (c.mark - k) : (c.mark + k);      // "other" coordinate of the running point
    int e = 0, at = 0;
    const int own = isA ? (pb + k) : (pb - k);       // coordinate on this chain's own axis
    if (own % TS != off) {                           // seed not on a trace boundary: merge first segment
        h = c.prev;
        int pa, d;
        if (h < 0) { pa = isA ? ey : ex; d = ediff; }
        else { c = cells[h]; k = c.diag; pa = isA ? (c.mark - k) : (c.mark + k); d = c.diff; }
        if (flen == 0) {
            tr[--at] = (uint16_t)(pb - pa);
            tr[--at] = isA ? (uint16_t)(d - e) : (uint16_t)(pb - pa);   // sic (align.c:1843-1844)
        } else { tr[1] = (uint16_t)(tr[1] + (pb - pa)); tr[0] = (uint16_t)(tr[0] + (d - e)); }
        pb = pa; e = d;
    }
    if (h >= 0) {
        for (h = c.prev; h >= 0; h = c.prev) {
            c = cells[h]; k = c.diag;
            const int pa = isA ? (c.mark - k) : (c.mark + k);
            tr[--at] = (uint16_t)(pb - pa);
            tr[--at] = (uint16_t)(c.diff - e);
            pb = pa; e = c.diff;
        }
        const int endo = isA ? ey : ex;              // end point on the "other" axis
        const int endown = isA ? ex : ey;
        const int lastown = isA ? (pb + k) : (pb - k);
        if (lastown != endown) { tr[--at] = (uint16_t)(pb - endo); tr[--at] = (uint16_t)(ediff - e); }
        else if (pb != endo) { tr[at + 1] = (uint16_t)(tr[at + 1] + (pb - endo)); tr[at] = (uint16_t)(tr[at] + (ediff - e)); }
    }
    return -at;
}

}  // namespace wc
}  // namespace damar
