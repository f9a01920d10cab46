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

#if defined(__CUDACC__)
#define WC_HD __host__ __device__ __forceinline__
#define WC_RARE inline __host__ __device__ __noinline__      // one copy of the code paths only few cells take
#else
#define WC_HD inline
#define WC_RARE inline
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
WC_RARE int slide(const uint32_t *__restrict__ aw, uint32_t ax, const uint32_t *__restrict__ bw, uint32_t by, int maxn) {
    int n = 0;
    while (n < maxn) {
        const uint64_t d = window_at(aw, ax + (uint32_t)n) ^ window_at(bw, by + (uint32_t)n);
        if (d != 0) { n += clz64(d) >> 1; break; }
        n += 32;
    }
    return n < maxn ? n : maxn;
}
WC_HD uint64_t push_matches(uint64_t b, int n) {             // b << n with n ones shifted in
    return (n >= 64) ? ~0ull : ~(~b << n);
}

// ---- pebble cell (align.c:399-405): {prev, diag, diff, mark}, original coordinates ----------------
// The reference keeps two chains per path, the A trace points and the B trace points.  Here a path has ONE chain in
// creation order, each cell tagged with its axis (bit 30 of diff): the front then carries one head per diagonal instead
// of two (16 bytes per diagonal, 14 warps per SM instead of 11), and the two chains of the reference are the two
// tag-filtered subsequences of it.
struct Cell { int32_t prev, diag, diff, mark; };
constexpr int32_t CELL_B = 1 << 30;                  // tag in Cell::diff: a B-axis trace point
constexpr int32_t CELL_DIFF = CELL_B - 1;

// ---- per-thread view of its wave-front ring ------------------------------------------------------
// word arrays: V (furthest anti-diagonal), T lo/hi (match history), H (head of the path's pebble chain, pool index).
// The reference also keeps, per diagonal, the next trace point to cross (NA, NB) and relies on the mark of a
// path's head pebble; both are functions of the previous wave here (see cellsN): 16 bytes per diagonal.
enum { F_V = 0, F_TLO = 1, F_THI = 2, F_H = 3, FRONT_WORDS = 4 };
template <int RING, int STRIDE>
struct Front {
    uint32_t *base;
    static constexpr int ARR = RING * STRIDE;                   // words between two arrays of one slot
    WC_HD uint32_t *slot(int k) const { return base + (k & (RING - 1)) * STRIDE; }
    WC_HD int ring() const { return RING; }
};

// The same with the ring size chosen at run time: 32 lanes x 32 diagonals or 16 lanes x 64 diagonals occupy the same
// 4 x 1024 words of a warp, so a warp can change between the two shapes (pair_align.cu::wave_tasks).
struct FrontRT {
    uint32_t *base; int rmask, sshift;                           // ring - 1, log2(threads that share the region)
    static constexpr int ARR = 1024;
    WC_HD uint32_t *slot(int k) const { return base + ((k & rmask) << sshift); }
    WC_HD int ring() const { return rmask + 1; }
};

struct PathState { uint64_t T; uint32_t h; };
template <class FRONT> WC_HD PathState load_path(const uint32_t *s) {
    PathState p;
    p.T = (uint64_t)s[F_TLO * FRONT::ARR] | ((uint64_t)s[F_THI * FRONT::ARR] << 32);
    p.h = s[F_H * FRONT::ARR];
    return p;
}

// ---- the trim test tables (align.c:286-318) in 768 bytes ------------------------------------------
// TABLE[x] (x = 15 match bits, oldest first) is the minimum score over the non-empty suffixes of x and
// SCORE[x] its total, with +mscore per match and -dscore per difference, mscore + dscore = 1000.  A suffix
// of x either lies inside the newest 7 bits or is all of them plus a suffix of the older 8, hence
//   TABLE[x] = min(T7[x & 127], S7(x & 127) + T8[x >> 7]),   S(x) = 1000 * popcount(x) - dscore * bits.
struct TrimTabs { const int16_t *t7, *t8; int dscore; };
WC_HD int min_suffix_score(int x, int bits, int dscore) {
    int best = IMAX;
    for (int j = 1; j <= bits; j++) {
        const int ones = popc64((uint64_t)(x & ((1 << j) - 1)));
        const int v = 1000 * ones - dscore * j;
        if (v < best) best = v;
    }
    return best;
}
WC_HD int table15(const TrimTabs &t, int x) {
    const int a0 = x & 127, a1 = x >> 7;
    const int u = t.t7[a0], v = t.t8[a1] + 1000 * popc64((uint64_t)a0) - 7 * t.dscore;
    return u < v ? u : v;
}
WC_HD int score15(const TrimTabs &t, int x) { return 1000 * popc64((uint64_t)x) - 15 * t.dscore; }
// the trim test of align.c:922-925:  TABLE[lo15] >= 0 && TABLE[hi15] + SCORE[lo15] >= 0  for the newest 30 bits of b,
// i.e. no suffix of the 30-bit window scores below zero; the window is walked in chunks of 7, 8, 7, 8 bits
WC_HD bool trim_ok(const TrimTabs t, uint64_t b) {
    const uint32_t w = (uint32_t)b & 0x3fffffffu;
    const int a0 = (int)(w & 127u), a1 = (int)((w >> 7) & 255u), a2 = (int)((w >> 15) & 127u), a3 = (int)(w >> 22);
    const int s0 = 1000 * popc64((uint64_t)a0) - 7 * t.dscore;                    // score of the newest 7
    const int s1 = 1000 * popc64((uint64_t)(w & 0x7fffu)) - 15 * t.dscore;        // ... newest 15
    const int s2 = 1000 * popc64((uint64_t)(w & 0x3fffffu)) - 22 * t.dscore;      // ... newest 22
    int m = t.t7[a0];
    const int m1 = s0 + t.t8[a1], m2 = s1 + t.t7[a2], m3 = s1 + (1000 * popc64((uint64_t)a2) - 7 * t.dscore) + t.t8[a3];
    (void)s2;
    // TABLE[lo15] = min(m, m1); TABLE[hi15] + SCORE[lo15] = min(m2, m3)
    if (m1 < m) m = m1;
    if (m2 < m) m = m2;
    if (m3 < m) m = m3;
    return m >= 0;
}

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
    int reach = 1;                  // Align_Spec.reach (align.c:1009): end at the boundary candidate when there is one (daligner: 1, datander: 0)
};
struct TaskOut {                    // original coordinates
    int32_t ex, ey, ediff;
    int32_t ha, hb;                 // head of the chosen pebble chain (both: the merged chain)
    int32_t status;
    int32_t waves, cells;           // work counters (wave cells after pruning, as the reference's WAVE_STATS)
};

struct Lane {
    // task constants
    const uint32_t *aw, *bw;
    uint32_t abase, bbase;
    int alen, blen, minp, maxp, TS, path_ave, rev, reach;
    int goffA, goffB;               // trace grid offsets in pass coordinates
    uint64_t magic;                 // ceil(2^40 / TS)
    uint32_t magic32;               // ceil(2^32 / TS), or 0 when that is not exact for these reads
    // wave state (pass coordinates)
    int low, hgh, k, dif, more, aclip, bclip;
    int besta, besty, lasta, trima, trimy, trimd;
    int morea, morey, mored, morem;
    uint32_t trimh, moreh;
    // sweep registers: previous-wave values of the diagonals above the cursor
    int vp, vc;                     // V[k+1], V[k]
    PathState pp;                   // path state of k+1
    // pebble allocation
    uint32_t cell_next, cell_left;
    int status;
    uint32_t waves, cells;
};

WC_HD int grid(int goff, int idx1, int TS) { return goff + (idx1 - 1) * TS; }   // idx1 = TP index + 1

// The one copy of the pebble store: takes a fresh chunk from the pool when the thread's current one is used up.
// Returns the cell's pool index, or NOCELL when the pool is dry.  Everything is passed by value so that the
// caller's state stays in registers across the call.
template <class POOL>
WC_RARE uint32_t push_raw(POOL pool, uint32_t next, uint32_t left, uint32_t prev, int diag, int diff, int mark) {
    if (left == 0) {
        next = pool.take(CELL_CHUNK);
        if (next == NOCELL) return NOCELL;
    }
    Cell c;
    c.prev = (int32_t)prev; c.diag = diag; c.diff = diff; c.mark = mark;
    pool.store(next, c);
    return next;
}
template <class POOL>
WC_HD uint32_t push_cell(Lane &L, POOL &pool, uint32_t prev, int k, int dif, int markA_or_B_pass, bool isA) {
    // back to original coordinates
    const int diag = L.rev ? (L.alen - L.blen) - k : k;
    const int mark = L.rev ? (isA ? L.alen : L.blen) - markA_or_B_pass : markA_or_B_pass;
    const uint32_t id = push_raw(pool, L.cell_next, L.cell_left, prev, diag, isA ? dif : (dif | CELL_B), mark);
    if (id == NOCELL) { L.status = ST_POOL; return prev; }
    L.cell_left = (L.cell_left ? L.cell_left : (uint32_t)CELL_CHUNK) - 1u;
    L.cell_next = id + 1u;
    return id;
}

// floor((v - goff) / TS) + 1 for v - goff >= -TS
WC_HD int tp_index1(int v, int goff, int TS) { return (v - goff + TS) / TS; }
// the same with the lane's reciprocal ceil(2^40 / TS): exact for reads shorter than 2^24 bases
WC_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// magic32 = ceil(2^32 / TS) is exact while u * TS < 2^32, which setup() checks against the read lengths
WC_HD int tp_index1(const Lane &L, int v, int goff) {
    const uint32_t u = (uint32_t)(v - goff + L.TS);
    return L.magic32 ? (int)umulhi32(u, L.magic32) : (int)(((uint64_t)u * L.magic) >> 40);
}

// Wave 0 (align.c:491-626 forward, 1206-1340 reverse) and the lane's constants.
template <class FRONT, class POOL>
WC_HD void setup(Lane &L, const TaskIn &t, const FRONT &F, POOL &pool) {
    L.aw = t.aw; L.bw = t.bw; L.abase = t.abase; L.bbase = t.bbase;
    L.alen = t.alen; L.blen = t.blen; L.TS = t.TS; L.path_ave = t.path_ave; L.rev = t.rev; L.reach = t.reach;
    L.cell_left = 0; L.cell_next = 0;
    L.waves = 0; L.cells = 0;
    const int TS = t.TS;
    L.magic = ((1ull << 40) + (uint64_t)TS - 1) / (uint64_t)TS;
    {
        const uint64_t umax = (uint64_t)(t.alen > t.blen ? t.alen : t.blen) + 2 * (uint64_t)TS;
        L.magic32 = (TS > 1 && umax * (uint64_t)TS < (1ull << 32)) ? (uint32_t)(((1ull << 32) + (uint64_t)TS - 1) / (uint64_t)TS) : 0u;
    }
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
    uint32_t h = push_cell(L, pool, NOCELL, k, 0, t.rev ? x : grid(L.goffA, naI, TS), true);
    h = push_cell(L, pool, h, k, 0, t.rev ? y : grid(L.goffB, nbI, TS), false);
    L.trimh = L.moreh = h;
    naI += 1; nbI += 1;
    {
        const int ra = L.alen - x, rb = L.blen - y;
        const int n = slide(L.aw, L.abase + (uint32_t)x, L.bw, L.bbase + (uint32_t)y, ra < rb ? ra : rb);
        x += n; y += n;
    }
    if (y == L.blen) { L.more = 0; if (L.bclip < k) L.bclip = k; }
    else if (x == L.alen) { L.more = 0; L.aclip = k; }
    while (x >= grid(L.goffA, naI, TS)) { h = push_cell(L, pool, h, k, 0, grid(L.goffA, naI, TS), true); naI++; }
    while (y >= grid(L.goffB, nbI, TS)) { h = push_cell(L, pool, h, k, 0, grid(L.goffB, nbI, TS), false); nbI++; }
    const int c = (y << 1) + k;
    if (c > L.besta) { L.besta = L.trima = L.lasta = c; L.besty = L.trimy = y; L.trimh = h; }
    uint32_t *s = F.slot(k);
    s[F_V * FRONT::ARR] = (uint32_t)c;
    s[F_TLO * FRONT::ARR] = (uint32_t)PATH_INT; s[F_THI * FRONT::ARR] = (uint32_t)(PATH_INT >> 32);
    s[F_H * FRONT::ARR] = h;
    L.k = k - 1;                                  // wave 0 is complete: next step is a transition
}

template <class FRONT>
WC_HD void reach_candidate(Lane &L, const FRONT &F, int kc) {          // align.c:947-974
    const uint32_t *s = F.slot(kc);
    const PathState p = load_path<FRONT>(s);
    const int mc = popc64(p.T & PATH_MASK61);
    if (L.morem <= mc) {
        L.morem = mc; L.morea = (int32_t)s[F_V * FRONT::ARR]; L.morey = (L.morea - kc) / 2; L.mored = L.dif;
        L.moreh = p.h;
    }
}

// End of a wave: boundary clipping (align.c:628-658, 943-975), pruning (:977-986), the loop test
// (:667) and the widening of the next wave (:756-772).  Leaves the lane either finished
// (status != ST_RUN) or positioned on k = hgh of the new wave.
template <class FRONT>
WC_HD void transition(Lane &L, const FRONT &F) {
    if (L.more == 0) {
        if (L.besty < L.blen && L.besta - L.besty < L.alen) L.more = 1;
        if (L.hgh >= L.aclip) { const int kc = L.aclip; L.hgh = kc - 1; reach_candidate(L, F, kc); }
        if (L.low <= L.bclip) { const int kc = L.bclip; L.low = kc + 1; reach_candidate(L, F, kc); }
        L.aclip = IMAX; L.bclip = -IMAX;
    }
    if (L.dif > 0) {
        // align.c:977-986: drop lagging diagonals from the top, then from the bottom.  Mostly 0 - 2 go on either side:
        // the four front values that decide that are fetched at once instead of one dependent load per test.
        const int n = L.besta - WAVE_LAG;
        int hgh = L.hgh, low = L.low;
        int vh[4], vl[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { vh[i] = (int32_t)F.slot(hgh - i)[F_V * FRONT::ARR]; vl[i] = (int32_t)F.slot(low + i)[F_V * FRONT::ARR]; }
        {
            const int w = hgh - low + 1;
            int cut = 0;
#pragma unroll
            for (int i = 3; i >= 0; i--) cut = (i < w && vh[i] < n) ? cut + 1 : 0;      // leading run of lagging diagonals
            hgh -= cut;
            if (cut == 4) while (hgh >= low && (int32_t)F.slot(hgh)[F_V * FRONT::ARR] < n) hgh -= 1;
        }
        if (hgh >= low) {                                 // V[hgh] >= n stops the walk at the latest
            int cut = 0;
#pragma unroll
            for (int i = 3; i >= 0; i--) cut = (vl[i] < n && low + i < hgh) ? cut + 1 : 0;
            low += cut;
            if (cut == 4) while ((int32_t)F.slot(low)[F_V * FRONT::ARR] < n) low += 1;
        }
        L.hgh = hgh; L.low = low;
        L.waves += 1;
        if (L.hgh >= L.low) L.cells += (uint32_t)(L.hgh - L.low + 1);
    }
    // an empty front (every diagonal ran off a read end) ends the pass; the reference reads stale memory there
    if (!(L.more && L.lasta >= L.besta - TRIM_MLAG) || L.hgh < L.low) { L.status = ST_DONE; return; }

    int low = L.low - 1, hgh = L.hgh + 1;
    if (hgh - low + 1 > F.ring()) { L.status = ST_WIDE; return; }
    if (low >= L.minp) F.slot(low)[F_V * FRONT::ARR] = (uint32_t)-1; else low += 1;
    if (hgh <= L.maxp) { F.slot(hgh)[F_V * FRONT::ARR] = (uint32_t)-1; L.vc = -1; }
    else { hgh -= 1; L.vc = (int32_t)F.slot(hgh)[F_V * FRONT::ARR]; }
    L.low = low; L.hgh = hgh;
    L.dif += 1;
    L.vp = -1;
    L.pp.T = PATH_INT; L.pp.h = NOCELL;                    // the (never chosen) state above hgh
    L.k = hgh;
}

// The three-way furthest-reaching choice (align.c:793-825) with its tie rules.
WC_HD int choose(int vp, int vc, int vm, int &c) {               // returns +1 / 0 / -1: predecessor diagonal offset
    const bool low = vc < vm;                                     // k - 1 beats k (ties go to k)
    const int m = low ? vm : vc;
    const bool up = m < vp;                                       // k + 1 beats both (ties go to the lower diagonal)
    c = up ? vp + 1 : (low ? vm + 1 : vc + 2);
    return up ? 1 : (low ? -1 : 0);
}

// What follows a cell's predecessor choice and slide (align.c:836-928) is done for the N cells of a cellsN call in
// three converged phases instead of cell by cell: (1) clips, best / lasta and the trace points crossed, (2) the
// pebbles, (3) the trim test.  Pebbles and the trim test are rare per cell but not per warp: inlined per cell they
// made the warp run 4 x 32 mostly idle copies of that code per call.
//
// Trace points.  The reference keeps NA[k] / NB[k] = first trace point beyond diagonal k's previous furthest
// point, and logs a pebble for every NA[k] <= x that the arriving path has not logged yet (its head pebble's
// mark < NA[k]).  A path has always logged every trace point up to its own previous position (xp, yp), and the
// chosen predecessor is never behind diagonal k's previous point on either axis (the three-way rule picks the
// larger anti-diagonal), so the pebbles logged are exactly the trace points in (xp, x] and (yp, y]: nothing
// but the previous wave's V is needed.  One step from the predecessor costs x and/or y one base:
// d = +1 (from k + 1): y only; d = -1 (from k - 1): x only; d = 0: both.
//
// Trim point.  The reference tests every new best point in visiting order and keeps the last one that passes
// (align.c:918-927); the candidates of a call are therefore tested newest first until one passes.
struct CellOut { uint64_t b; int c, y, ia, ja, ib, jb; };

// first probe of a slide: equal bases in the first 16-base window (16 = all equal).  At 12-15 % error a slide passes 16
// bases in a few percent of the cells only, and 16 bases need two words per read instead of three.
constexpr int PROBE = 16;
WC_HD int clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
WC_HD uint32_t window16_at(const uint32_t *__restrict__ words, uint32_t s) {
    const uint32_t *w = words + (s >> 4);
    return funnel_l(ldg32(w + 1), ldg32(w), (int)(s & 15u) * 2);
}
WC_HD int probe16(const uint32_t *__restrict__ aw, uint32_t ax, const uint32_t *__restrict__ bw, uint32_t by) {
    const uint32_t d = window16_at(aw, ax) ^ window16_at(bw, by);
    return d ? (clz32(d) >> 1) : PROBE;
}

// N furthest-reaching-point updates per call, on diagonals L.k, L.k - 1, ... (align.c:789-938).  All cells of a
// wave read only the previous wave, so the N predecessor choices and slides are independent instruction streams:
// with a few warps per SM this instruction-level parallelism is what hides latency.  The bookkeeping that the
// reference does in visiting order (clips, best, trim) is then applied to k, k - 1, ... in turn.
template <int N, class FRONT, class POOL>
WC_HD void cellsN(Lane &L, const FRONT &F, const TrimTabs &tabs, POOL &pool) {
    const int k0 = L.k;
    const int cnt = (k0 - L.low + 1 < N) ? (k0 - L.low + 1) : N;      // cells handled by this call (>= 1)
    // slot j holds diagonal k0 - j; slots beyond the front alias slot 0 and are never used
    uint32_t *s[N + 1];
    int v[N + 2];                                                       // v[j + 1] = V[k0 - j] of the previous wave, v[0] = V[k0 + 1]
    PathState P[N + 2];                                                 // likewise
    v[0] = L.vp; v[1] = L.vc; P[0] = L.pp;
#pragma unroll
    for (int j = 0; j <= N; j++) {
        const bool in = (k0 - j >= L.low) && (j <= cnt);
        s[j] = F.slot(in ? k0 - j : k0);
        if (j >= 1) v[j + 1] = in ? (int32_t)s[j][F_V * FRONT::ARR] : -1;
    }
#pragma unroll
    for (int j = 0; j < N; j++) P[j + 1] = load_path<FRONT>(s[j]);
    P[N + 1] = P[N];                                                    // loaded below only if it is chosen
    int c[N], n[N], mx[N], dd[N];
    PathState q[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        const int d = choose(v[j], v[j + 1], v[j + 2], c[j]);
        dd[j] = d;
        q[j] = (d > 0) ? P[j] : (d == 0 ? P[j + 1] : P[j + 2]);
        if (j == N - 1 && j < cnt && d < 0) q[j] = load_path<FRONT>(s[N]);
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
        const int jj = (j < cnt) ? j : 0;
        const int kk = k0 - jj, y = ((j < cnt ? c[j] : c[0]) - kk) >> 1, x = y + kk;
        mx[j] = (L.alen - x < L.blen - y) ? L.alen - x : L.blen - y;
        n[j] = probe16(L.aw, L.abase + (uint32_t)x, L.bw, L.bbase + (uint32_t)y);
    }
    // phase 1: end of the slide, clips, best / lasta, trace points crossed (visiting order)
    CellOut o[N];
    uint32_t pm = 0, slow = 0;                   // bit 2j / 2j+1: cell j crosses exactly one A / B trace point; slow: more than one
    uint32_t cm = 0;                             // bit j: cell j is a trim candidate (new best with a dense enough history)
#pragma unroll
    for (int j = 0; j < N; j++) {
        o[j].ia = o[j].ja = o[j].ib = o[j].jb = 0; o[j].b = 0; o[j].c = 0; o[j].y = 0;
        if (j < cnt) {
            const int kk = k0 - j, y0 = (c[j] - kk) >> 1, x0 = y0 + kk;
            if (n[j] == PROBE && mx[j] > PROBE) n[j] = PROBE + slide(L.aw, L.abase + (uint32_t)(x0 + PROBE), L.bw, L.bbase + (uint32_t)(y0 + PROBE), mx[j] - PROBE);
            if (n[j] > mx[j]) n[j] = mx[j];
            const int d = dd[j];
            const uint64_t b = push_matches(q[j].T << 1, n[j]);
            const int y = y0 + n[j], x = x0 + n[j];
            if (y == L.blen) { L.more = 0; if (L.bclip < kk) L.bclip = kk; }
            else if (x == L.alen) { L.more = 0; L.aclip = kk; }
            const int cc = (y << 1) + kk;
            // trace points in (xp, x]: the last one at or below x lies rem below it, so it is crossed iff x - xp > rem
            // and a second one iff x - xp > rem + TS (long slides only: counted exactly there)
            const int ja = tp_index1(L, x, L.goffA), jb = tp_index1(L, y, L.goffB);
            const int rema = x - grid(L.goffA, ja, L.TS), remb = y - grid(L.goffB, jb, L.TS);
            const int sa = n[j] + (d <= 0 ? 1 : 0), sb = n[j] + (d >= 0 ? 1 : 0);          // x - xp, y - yp
            o[j].ja = ja; o[j].jb = jb;
            if (sa > rema) pm |= 1u << (2 * j);
            if (sb > remb) pm |= 2u << (2 * j);
            if (sa > rema + L.TS || sb > remb + L.TS) {
                slow |= 1u << j;
                o[j].ia = tp_index1(L, x - sa, L.goffA); o[j].ib = tp_index1(L, y - sb, L.goffB);
            }
            o[j].b = b; o[j].c = cc; o[j].y = y;
            if (cc > L.besta) {                                                  // align.c:911-917
                L.besta = cc; L.besty = y;
                if (popc64(b & PATH_MASK61) >= L.path_ave) { L.lasta = cc; cm |= 1u << j; }
            }
        }
    }
    // phase 2: pebbles (align.c:859-909), one at a time through the single copy of the store
    while (slow) {                               // cells that crossed several trace points of one axis in one slide: A chain, then B chain
        const int js = 63 - clz64((uint64_t)(slow & (0u - slow)));
        slow &= slow - 1u;
        pm &= ~(3u << (2 * js));
        int ia = 0, ja = 0, ib = 0, jb = 0; uint32_t h = 0;
#pragma unroll
        for (int j = 0; j < N; j++) if (j == js) { ia = o[j].ia; ja = o[j].ja; ib = o[j].ib; jb = o[j].jb; h = q[j].h; }
        for (int i = ia + 1; i <= ja; i++) h = push_cell(L, pool, h, k0 - js, L.dif, grid(L.goffA, i, L.TS), true);
        for (int i = ib + 1; i <= jb; i++) h = push_cell(L, pool, h, k0 - js, L.dif, grid(L.goffB, i, L.TS), false);
#pragma unroll
        for (int j = 0; j < N; j++) if (j == js) q[j].h = h;
    }
    while (pm) {
        const int bit = 63 - clz64((uint64_t)(pm & (0u - pm)));
        pm &= pm - 1u;
        const int js = bit >> 1;
        const bool isA = (bit & 1) == 0;
        int idx = 0; uint32_t head = 0;
#pragma unroll
        for (int j = 0; j < N; j++) if (j == js) { idx = isA ? o[j].ja : o[j].jb; head = q[j].h; }
        const uint32_t id = push_cell(L, pool, head, k0 - js, L.dif, grid(isA ? L.goffA : L.goffB, idx, L.TS), isA);
#pragma unroll
        for (int j = 0; j < N; j++) if (j == js) q[j].h = id;
    }
    // phase 3: the trim point = the newest candidate that passes the test (align.c:918-927); it is nearly always the first tried
    while (cm) {
        const int nj = 63 - clz64((uint64_t)cm);
        cm &= ~(1u << nj);
        uint64_t nb = 0; int nc = 0, ny = 0; uint32_t h = 0;
#pragma unroll
        for (int j = 0; j < N; j++) if (j == nj) { nb = o[j].b; nc = o[j].c; ny = o[j].y; h = q[j].h; }
        if (trim_ok(tabs, nb)) { L.trima = nc; L.trimy = ny; L.trimd = L.dif; L.trimh = h; break; }
    }
    // the new front entries
#pragma unroll
    for (int j = 0; j < N; j++)
        if (j < cnt) {
            uint32_t *t = s[j];
            t[F_V * FRONT::ARR] = (uint32_t)o[j].c;
            t[F_TLO * FRONT::ARR] = (uint32_t)o[j].b; t[F_THI * FRONT::ARR] = (uint32_t)(o[j].b >> 32);
            t[F_H * FRONT::ARR] = q[j].h;
        }
    // carry the previous-wave values of the diagonal just above the next cursor
    L.vp = v[N]; L.vc = v[N + 1]; L.pp = P[N];
    L.k = k0 - cnt;
}

// end point of the pass in original coordinates (align.c:1001-1012; REACH = 1 for daligner, daligner.c:951, 0 for datander)
WC_HD void finish(const Lane &L, TaskOut &o) {
    int ex, ey;
    if (L.morem >= 0 && L.reach) { ex = L.morea - L.morey; ey = L.morey; o.ediff = L.mored; o.ha = o.hb = (int32_t)L.moreh; }
    else { ex = L.trima - L.trimy; ey = L.trimy; o.ediff = L.trimd; o.ha = o.hb = (int32_t)L.trimh; }
    if (L.rev) { ex = L.alen - ex; ey = L.blen - ey; }
    o.ex = ex; o.ey = ey;
    o.status = L.status;
    o.waves = (int32_t)L.waves; o.cells = (int32_t)L.cells;
}

// ---- trace emission from the pebble chains (align.c:1001-1111 forward, 1711-1880 reverse) --------
// `cells` is the pool; chains end with prev = -1.  MERGED: the chain holds the A and the B trace points of the path in
// creation order, tagged (CELL_B); the A (B) chain of the reference is its tag-filtered subsequence.  !MERGED: the chain
// is one axis only (the strided warp kernel keeps two chains).
template <bool MERGED> WC_HD bool on_axis(const Cell &c, bool isA) { return !MERGED || (((c.diff & CELL_B) == 0) == isA); }
template <bool MERGED> WC_HD int cell_diff(const Cell &c) { return MERGED ? (c.diff & CELL_DIFF) : c.diff; }
// first cell of the axis at or behind h (following prev links)
template <bool MERGED> WC_HD int same_axis(const Cell *cells, int h, bool isA) {
    while (h >= 0 && !on_axis<MERGED>(cells[h], isA)) h = cells[h].prev;
    return h;
}

// Forward: pairs are appended at tr[0..); returns the number of values written.
template <bool MERGED = false>
WC_HD int emit_chain_fwd(const Cell *cells, int head, bool isA, int anti, int ex, int ey, int ediff, uint16_t *tr) {
    head = same_axis<MERGED>(cells, head, isA);
    int n = 0;
    for (int h = head; h >= 0; h = same_axis<MERGED>(cells, cells[h].prev, isA)) n++;
    // cells in forward order are c_0 (root) .. c_{n-1} (head); pair i (i>=1) = (diff_i - diff_{i-1}, p_i - p_{i-1})
    int len = 2 * (n - 1);
    Cell cur = cells[head];
    const int lastk = cur.diag;
    int lastp, laste;
    for (int i = n - 1; i >= 1; i--) {
        const Cell prv = cells[same_axis<MERGED>(cells, cur.prev, isA)];
        const int pc = isA ? (cur.mark - cur.diag) : (cur.mark + cur.diag);
        int pp, pe;
        if (i == 1) { pp = isA ? ((anti - prv.diag) / 2) : ((anti + prv.diag) / 2); pe = 0; }
        else { pp = isA ? (prv.mark - prv.diag) : (prv.mark + prv.diag); pe = cell_diff<MERGED>(prv); }
        tr[2 * (i - 1)] = (uint16_t)(cell_diff<MERGED>(cur) - pe);
        tr[2 * (i - 1) + 1] = (uint16_t)(pc - pp);
        cur = prv;
    }
    {
        const Cell hc = cells[head];
        if (n == 1) { lastp = isA ? ((anti - hc.diag) / 2) : ((anti + hc.diag) / 2); laste = 0; }
        else { lastp = isA ? (hc.mark - hc.diag) : (hc.mark + hc.diag); laste = cell_diff<MERGED>(hc); }
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

// Reverses the prev links of a chain in place (the reverse pass emits from its root onward); returns the root.
WC_HD int reverse_chain(Cell *cells, int head) {
    int a = -1, b;
    for (int h = head; h >= 0; h = b) { b = cells[h].prev; cells[h].prev = a; a = h; }
    return a;
}

// Reverse pass, on a chain already reversed by reverse_chain (`root`): `flen` = number of values the forward pass
// left at tr[0..); pairs are prepended before tr[0] and the first partial segment is merged into tr[0], tr[1].
// Returns the number of values prepended.
template <bool MERGED = false>
WC_HD int emit_chain_rev(const Cell *cells, int root, bool isA, int ex, int ey, int ediff, int TS, int off, int flen, uint16_t *tr) {
    int h = same_axis<MERGED>(cells, root, isA);             // root of this axis (links now lead toward the head)
    Cell c = cells[h];
    int k = c.diag;
    int pb = isA ? (c.mark - k) : (c.mark + k);      // "other" coordinate of the running point
    int e = 0, at = 0;
    const int own = isA ? (pb + k) : (pb - k);       // coordinate on this chain's own axis
    if (own % TS != off) {                           // seed not on a trace boundary: merge first segment
        h = same_axis<MERGED>(cells, c.prev, isA);
        int pa, d;
        if (h < 0) { pa = isA ? ey : ex; d = ediff; }
        else { c = cells[h]; k = c.diag; pa = isA ? (c.mark - k) : (c.mark + k); d = cell_diff<MERGED>(c); }
        if (flen == 0) {
            tr[--at] = (uint16_t)(pb - pa);
            tr[--at] = isA ? (uint16_t)(d - e) : (uint16_t)(pb - pa);   // sic (align.c:1843-1844)
        } else { tr[1] = (uint16_t)(tr[1] + (pb - pa)); tr[0] = (uint16_t)(tr[0] + (d - e)); }
        pb = pa; e = d;
    }
    if (h >= 0) {
        for (h = same_axis<MERGED>(cells, c.prev, isA); h >= 0; h = same_axis<MERGED>(cells, c.prev, isA)) {
            c = cells[h]; k = c.diag;
            const int pa = isA ? (c.mark - k) : (c.mark + k);
            tr[--at] = (uint16_t)(pb - pa);
            tr[--at] = (uint16_t)(cell_diff<MERGED>(c) - e);
            pb = pa; e = cell_diff<MERGED>(c);
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
