// seed_merge.cu -- stage 2: A-versus-B merge of two sorted k-mer lists into seed hits, then the
// (bread, aread, apos) sort.
//
// Replaces count_thread (reference dalign/filter.c:1039-1165), the -M limit choice
// (filter.c:2634-2702), merge_thread (filter.c:1170-1358) and the seed-hit lex_sort
// (filter.c:2561-2580, 2776-2781).
//
// Two sweeps over the lists (round 1 made six):
//   merge_bounds   one thread per tile of MT_TA A entries: the tile's first entry moved back to the start of its
//                  equal-code run, and the first B entry whose code is not smaller -- tiles hold whole runs and
//                  the B entries that can match a tile form one window
//   merge_tiles<false>  COUNT: a CTA stages a tile and its B window in shared memory; one thread per A ENTRY finds
//                  the entry's run and the matching B run by bisection in shared memory, its number of partners
//                  (all of the B run, or -- self comparison -- those of earlier reads) and adds it to the run's
//                  match count ct; runs with ct < limit contribute to the tile's hit count, every run to the
//                  10000-bin histogram of the -M rule.  The limit is not known before the histogram is: the sweep
//                  runs with the limit the rule yields whenever memory is plentiful (MAXGRAM) and is repeated
//                  only if the host then picks a smaller one.
//   (scan of the tile counts)
//   merge_tiles<true>   EMIT: the same staging and per-entry work, an exclusive scan of the entries' hit counts,
//                  then one thread per HIT: which entry it belongs to (bisection of the scan), which B partner,
//                  (bread, bpos) of that partner -- all lanes busy and the stores coalesced, whatever the run lengths
// A run (or a window) too large for shared memory is handled by its CTA straight from global memory.
// The B run of a code is found through a bucket table over the window's code range (a lookup and a bisection of the few
// entries of one bucket).  A comparison with more hits than one sort takes is cut into chunks of B reads (api.cu): the
// sweeps then run `ranged`, counting and emitting only the partners whose position lies in the chunk.
// A hit is ONE 64-bit word  bread | aread | apos | bpos  whenever those fit (they do for every configuration of
// BASELINE.json), sorted keys-only on its upper fields; the stable sort keeps ties in merge order (= bpos ascending),
// which report_thread depends on.  Otherwise key + bpos pairs as in round 1.
#include <type_traits>

#include <mutex>

#include "common.cuh"

namespace damar {
namespace {

constexpr int MAXGRAM = 10000;     // filter.c:71
constexpr int MT_THREADS = 512;
constexpr int MT_TA = 1024;        // nominal A entries per tile
constexpr int MT_CAPA = 2048;      // A entries / B entries one shared-memory segment can hold
constexpr int MT_CAPB = 3072;
constexpr uint32_t HAS_RUN = 0x8000u;

// A sorted k-mer list: packed (gpos == nullptr, entry = code << gbits | gpos) or (code, gpos) pairs
struct KList {
    const uint64_t *ent; const uint32_t *gpos; int64_t len; int gbits;
    __device__ __forceinline__ uint64_t ecode(uint64_t e) const { return gpos ? e : (e >> gbits); }
    __device__ __forceinline__ uint64_t code(int64_t i) const { return ecode(ent[i]); }
    __device__ __forceinline__ uint32_t epos(uint64_t e, int64_t i) const { return gpos ? gpos[i] : (uint32_t)(e & ((1ull << gbits) - 1ull)); }
};
// the read whose slot range [boff[r], boff[r+1]) holds gpos p, from the block's lookup table (kmer_index.cu::build_rtab)
struct ReadMap { const uint32_t *boff, *rtab; };
__device__ __forceinline__ int read_of(const ReadMap m, uint32_t p) {
    int r = (int)m.rtab[p >> RTAB_SHIFT];
    while (m.boff[r + 1] <= p) r++;
    return r;
}

// first index in [lo, hi) whose code is >= c (GT: > c)
template <bool GT>
__device__ __forceinline__ int64_t bound_code(const KList &L, int64_t lo, int64_t hi, uint64_t c) {
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; const uint64_t x = L.code(m); if (GT ? (x <= c) : (x < c)) lo = m + 1; else hi = m; }
    return lo;
}
// start / end of the equal-code run that holds entry p (gallop, then bisect: runs are short)
__device__ int64_t run_start(const KList &L, int64_t p) {
    const uint64_t c = L.code(p);
    int64_t lo = p, w = 1;
    while (lo > 0) {
        int64_t q = lo - w; if (q < 0) q = 0;
        if (L.code(q) == c) { lo = q; w <<= 1; }
        else return bound_code<false>(L, q + 1, lo, c);
    }
    return 0;
}
__device__ int64_t run_end(const KList &L, int64_t p) {
    const uint64_t c = L.code(p);
    int64_t hi = p + 1, w = 1;
    while (hi < L.len) {
        int64_t q = hi + w - 1; if (q >= L.len) q = L.len - 1;
        if (L.code(q) == c) { hi = q + 1; w <<= 1; }
        else return bound_code<true>(L, hi, q, c);
    }
    return L.len;
}
// the same inside a staged segment: the codes of its entries, already separated from the positions (CT = uint32_t when a
// code has at most 32 bits -- k <= 16 -- which halves the shared-memory traffic of the bisections)
template <typename CT>
struct Staged {
    const CT *c; int n;
    __device__ __forceinline__ CT code(int i) const { return c[i]; }
};
template <bool GT, typename CT>
__device__ __forceinline__ int sbound(const Staged<CT> &S, int lo, int hi, CT c) {
    while (lo < hi) { const int m = (lo + hi) >> 1; const CT x = S.c[m]; if (GT ? (x <= c) : (x < c)) lo = m + 1; else hi = m; }
    return lo;
}
template <typename CT>
__device__ __forceinline__ int srun_start(const Staged<CT> &S, int p, CT c) {
    int lo = p, w = 1;
    while (lo > 0) {
        int q = lo - w; if (q < 0) q = 0;
        if (S.c[q] == c) { lo = q; w <<= 1; }
        else return sbound<false>(S, q + 1, lo, c);
    }
    return 0;
}
template <typename CT>
__device__ __forceinline__ int srun_end(const Staged<CT> &S, int p, CT c) {
    int hi = p + 1, w = 1;
    while (hi < S.n) {
        int q = hi + w - 1; if (q >= S.n) q = S.n - 1;
        if (S.c[q] == c) { hi = q + 1; w <<= 1; }
        else return sbound<true>(S, hi, q, c);
    }
    return S.n;
}

struct MergeArgs {
    KList A, B;
    ReadMap amap, bmap;                           // amap: the A block (== B block when self)
    int self, comp, identity, same_list;
    uint32_t limit;
    uint32_t ntiles;
    int64_t *tstart, *twin;                       // [ntiles + 1]
    unsigned long long *tcount;                   // [ntiles] hits of the tile (COUNT), their exclusive scan (EMIT)
    unsigned long long *gram;                     // MAXGRAM bins or nullptr
    uint64_t *keys; uint32_t *bpos;               // EMIT: bpos == nullptr -> one packed word per hit
    int pbits, abits, qbits;
    // A comparison with more hits than one sort takes is cut into chunks of B reads (api.cu): only partners whose position
    // lies in [g_lo, g_hi) count and are emitted.  The match count of a run that the -M limit looks at stays that of the
    // whole run.
    int ranged; uint32_t g_lo, g_hi;
};

__global__ void merge_bounds(MergeArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > a.ntiles) return;
    if (t == a.ntiles) { a.tstart[t] = a.A.len; a.twin[t] = a.B.len; return; }
    const int64_t s = (t == 0) ? 0 : run_start(a.A, (int64_t)t * MT_TA);
    a.tstart[t] = s;
    a.twin[t] = a.same_list ? s : bound_code<false>(a.B, 0, a.B.len, a.A.code(s));
}

// Threshold below which a B entry's gpos makes it a partner of the A entry at gpos ga in a self comparison:
//   default      : b.read <  a.read                          (filter.c:1108-1112)
//   -I, N        : (b.read,b.rpos) < (a.read,a.rpos)         (filter.c:1097-1104)
//   -I, C        : b.read <= a.read                          (filter.c:1092-1096)
// All three are "gposB < threshold" because a run is in gpos order and both blocks share boff.
__device__ __forceinline__ uint32_t self_threshold(const MergeArgs &a, uint32_t ga) {
    if (a.identity && !a.comp) return ga;
    const int r = read_of(a.amap, ga);
    return (a.identity && a.comp) ? a.amap.boff[r + 1] : a.amap.boff[r];
}

__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long *sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    unsigned long long t = 0;
    for (int x = 0; x < MT_THREADS / 32; x++) t += sh[x];
    return t;
}

constexpr int MT_NBUCK = 1024;                      // buckets over the code range of a window (lookup instead of an 11-step bisection)
constexpr int MT_SGRAM = 512;                       // histogram bins kept in shared memory; rarer, larger counts go to global atomics
template <typename CT, bool EMIT>
struct MtSmem {
    CT ac[MT_CAPA];                               // codes of the staged A entries
    CT bc[MT_CAPB];                               // ... of the B window
    uint32_t ap[MT_CAPA], bp[MT_CAPB];            // their positions (gpos)
    uint32_t cnt[MT_CAPA + 4];                    // partners per entry, then their exclusive scan
    uint32_t gct[MT_CAPA];                        // match count of the run that starts at this entry
    uint16_t jbw[MT_CAPA];                        // start of the entry's B run inside the window | HAS_RUN
    uint16_t rsw[MT_CAPA];                        // start of the entry's A run inside the segment
    union { uint32_t gram[EMIT ? 1 : MT_SGRAM]; uint64_t klow[EMIT ? MT_CAPA : 1]; } u;   // COUNT: histogram; EMIT: aread << pbits | apos
    unsigned long long red[MT_THREADS / 32];
    uint32_t wsum[MT_THREADS / 32];
    int64_t seg[4];                               // q, we, giant, -
    uint16_t bstart[MT_NBUCK + 2];                // first window entry whose code bucket is >= b: a bucket holds a handful of entries
};

template <bool EMIT, typename CT>
__global__ void __launch_bounds__(MT_THREADS, EMIT ? 2 : 3) merge_tiles(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MtSmem<CT, EMIT> &S = *reinterpret_cast<MtSmem<CT, EMIT> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool histo = !EMIT && a.gram != nullptr;
    const bool packed_out = a.bpos == nullptr;
    if (histo) { for (int i = tid; i < MT_SGRAM; i += MT_THREADS) S.u.gram[i] = 0; }
    __syncthreads();

    for (uint32_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int64_t a0 = a.tstart[tile], a1 = a.tstart[tile + 1];
        if (a1 <= a0) { if (!EMIT && tid == 0) a.tcount[tile] = 0; continue; }
        const int64_t b0 = a.twin[tile], b1 = a.twin[tile + 1];
        unsigned long long tile_hits = 0;                       // COUNT: hits so far; EMIT: output cursor
        if (EMIT) tile_hits = a.tcount[tile];
        int64_t p = a0, wb = b0;
        while (p < a1) {
            // ---- next segment [p, q) x [wb, we): whole runs that fit shared memory, or one oversized run ----
            int64_t q, we; bool giant = false;
            if (a1 - p <= MT_CAPA && (a.same_list || b1 - wb <= MT_CAPB)) { q = a1; we = b1; }
            else {
                __syncthreads();
                if (tid == 0) {
                    int64_t span = MT_CAPA, qq, ww; bool g = false;
                    for (;;) {
                        bool single = (span == 0);
                        if (!single) {
                            const int64_t cand = p + span;
                            qq = (cand >= a1) ? a1 : run_start(a.A, cand);              // whole runs in [p, qq)
                            single = qq <= p;                                            // the run at p alone is longer than `span`
                        }
                        if (single) qq = run_end(a.A, p);                                // <= a1: tiles end on run boundaries
                        ww = a.same_list ? qq : ((qq >= a.A.len) ? b1 : bound_code<false>(a.B, wb, b1, a.A.code(qq)));
                        const bool fits = (qq - p <= MT_CAPA) && (a.same_list || ww - wb <= MT_CAPB);
                        if (fits) break;
                        if (single) { g = true; break; }
                        span >>= 1;                                                      // window too large: fewer A entries
                    }
                    S.seg[0] = qq; S.seg[1] = ww; S.seg[2] = g ? 1 : 0;
                }
                __syncthreads();
                q = S.seg[0]; we = S.seg[1]; giant = S.seg[2] != 0;
            }

            if (!giant) {
                // ================= fast path: the segment and its window in shared memory =================
                const int na = (int)(q - p), nb = a.same_list ? na : (int)(we - wb);
                __syncthreads();
                for (int i = tid; i < na; i += MT_THREADS) {
                    const uint64_t e = a.A.ent[p + i];
                    S.ac[i] = (CT)a.A.ecode(e); S.ap[i] = a.A.epos(e, p + i); S.gct[i] = 0;
                }
                if (!a.same_list)
                    for (int i = tid; i < nb; i += MT_THREADS) {
                        const uint64_t e = a.B.ent[wb + i];
                        S.bc[i] = (CT)a.B.ecode(e); S.bp[i] = a.B.epos(e, wb + i);
                    }
                __syncthreads();
                const Staged<CT> SA{S.ac, na};
                const Staged<CT> SB{a.same_list ? S.ac : S.bc, nb};
                const uint32_t *bpos = a.same_list ? S.ap : S.bp;
                // Bucket table over the window's code range: bstart[b] = first window entry whose bucket is >= b, so the run
                // of a code is found by a lookup and a bisection of the few entries of its bucket instead of the whole window.
                CT c0 = 0; int bshift = 0;
                if (!a.same_list && nb > 0) {
                    c0 = S.bc[0];
                    const CT range = S.bc[nb - 1] - c0;
                    while ((range >> bshift) >= (CT)MT_NBUCK) bshift++;
                    for (int b = tid; b <= MT_NBUCK; b += MT_THREADS) S.bstart[b] = (uint16_t)nb;
                    __syncthreads();
                    for (int j = tid; j < nb; j += MT_THREADS) {
                        const int b = (int)((S.bc[j] - c0) >> bshift);
                        if (j == 0 || (int)((S.bc[j - 1] - c0) >> bshift) != b) S.bstart[b] = (uint16_t)j;   // the bucket's first entry
                    }
                    __syncthreads();
                    if (tid < 32) {                               // empty buckets take the start of the next non-empty one (suffix minimum)
                        constexpr int PERL = (MT_NBUCK + 32) / 32;
                        int run = nb;
                        uint16_t loc[PERL];
                        for (int x = PERL - 1; x >= 0; x--) { const int b = tid * PERL + x; int v = (b <= MT_NBUCK) ? S.bstart[b] : nb; if (v < run) run = v; loc[x] = (uint16_t)run; }
                        int suf = run;                            // minimum over this lane's buckets; now over all later lanes too
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, suf, o); if (tid + o < 32 && t < suf) suf = t; }
                        const int later = __shfl_down_sync(0xffffffffu, suf, 1);
                        const int cap = (tid < 31) ? later : nb;
                        for (int x = 0; x < PERL; x++) { const int b = tid * PERL + x; if (b <= MT_NBUCK) S.bstart[b] = (uint16_t)(loc[x] < cap ? loc[x] : cap); }
                    }
                    __syncthreads();
                }
                for (int i = tid; i < na; i += MT_THREADS) {
                    const CT c = S.ac[i];
                    const int rs = srun_start(SA, i, c);
                    int jb, ib;
                    if (a.same_list) { jb = rs; ib = srun_end(SA, i, c); }
                    else if (nb == 0 || c < c0 || c > S.bc[nb - 1]) { jb = ib = 0; }
                    else {
                        const int b = (int)((c - c0) >> bshift);
                        jb = sbound<false>(SB, (int)S.bstart[b], (int)S.bstart[b + 1], c);
                        ib = (jb < nb && S.bc[jb] == c) ? srun_end(SB, jb, c) : jb;
                    }
                    const int nbr = ib - jb;
                    uint32_t cn = 0, cn_run = 0;                                // partners emitted / partners in the whole run
                    int jr = jb;                                                // first partner
                    if (nbr > 0 && !(a.same_list && a.self && nbr == 1)) {     // a run of one k-mer of the list itself has no partner
                        int pe = ib;                                            // partners = [jb, pe)
                        if (a.self) {
                            const uint32_t thr = self_threshold(a, S.ap[i]);
                            int lo = jb, hi = ib;                               // B run entries with gpos < thr
                            while (lo < hi) { const int m = (lo + hi) >> 1; if (bpos[m] < thr) lo = m + 1; else hi = m; }
                            pe = lo;
                        }
                        cn = cn_run = (uint32_t)(pe - jb);
                        if (a.ranged) {                                         // ... of which those in [g_lo, g_hi): a run is in gpos order
                            int lo = jb, hi = pe;
                            while (lo < hi) { const int m = (lo + hi) >> 1; if (bpos[m] < a.g_lo) lo = m + 1; else hi = m; }
                            jr = lo; hi = pe;
                            while (lo < hi) { const int m = (lo + hi) >> 1; if (bpos[m] < a.g_hi) lo = m + 1; else hi = m; }
                            cn = (uint32_t)(lo - jr);
                        }
                    }
                    S.cnt[i] = cn;
                    S.jbw[i] = (uint16_t)((uint32_t)jr | (nbr > 0 ? HAS_RUN : 0u));
                    S.rsw[i] = (uint16_t)rs;
                    if (a.self) { if (cn_run) atomicAdd(&S.gct[rs], cn_run); }
                    else if (i == rs) {                                          // ct = |A run| x |B run| (filter.c:1148-1149)
                        const unsigned long long ct = (unsigned long long)(srun_end(SA, i, c) - rs) * (unsigned long long)nbr;
                        S.gct[rs] = ct > 0xffffffffull ? 0xffffffffu : (uint32_t)ct;
                    }
                }
                __syncthreads();
                for (int i = tid; i < na; i += MT_THREADS) {
                    const int rs = S.rsw[i];
                    const uint32_t ct = S.gct[rs];
                    const bool has = (S.jbw[i] & HAS_RUN) != 0;
                    if (!(has && ct < a.limit)) S.cnt[i] = 0;
                    if (histo && i == rs && has && ct < MAXGRAM) { if (ct < MT_SGRAM) atomicAdd(&S.u.gram[ct], 1u); else atomicAdd(&a.gram[ct], 1ull); }
                }
                __syncthreads();
                // exclusive scan of cnt[0, na) in place; cnt[na] = total
                uint32_t total;
                {
                    constexpr int PER = MT_CAPA / MT_THREADS;
                    uint32_t v[PER], s = 0;
#pragma unroll
                    for (int x = 0; x < PER; x++) { const int i = tid * PER + x; v[x] = (i < na) ? S.cnt[i] : 0u; s += v[x]; }
                    uint32_t inc = s;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                    if (lane == 31) S.wsum[w] = inc;
                    __syncthreads();
                    uint32_t off = 0, tot = 0;
#pragma unroll
                    for (int x = 0; x < MT_THREADS / 32; x++) { const uint32_t t = S.wsum[x]; if (x < w) off += t; tot += t; }
                    total = tot;
                    uint32_t run = off + inc - s;
#pragma unroll
                    for (int x = 0; x < PER; x++) { const int i = tid * PER + x; if (i < na) S.cnt[i] = run; run += v[x]; }
                    if (tid == 0) S.cnt[na] = tot;
                }
                if (EMIT && total > 0) {
                    __syncthreads();
                    for (int i = tid; i < na; i += MT_THREADS)
                        if (S.cnt[i + 1] > S.cnt[i]) {
                            const uint32_t ga = S.ap[i];
                            const int ar = read_of(a.amap, ga);
                            S.u.klow[i] = ((uint64_t)ar << a.pbits) | (uint64_t)(ga - a.amap.boff[ar]);
                        }
                    __syncthreads();
                    // hit-centric emission: one thread per hit, the entry found by bisecting the segment's scan (all lanes busy
                    // whatever the run lengths; letting an entry's thread write its own few hits measured 20 % slower: 4.2 against 3.5 ms)
                    for (uint32_t h = tid; h < total; h += MT_THREADS) {
                        int lo = 0, hi = na;                                 // last entry whose scan value is <= h
                        while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (S.cnt[m] <= h) lo = m; else hi = m; }
                        const int i = lo;
                        const int j = (int)(S.jbw[i] & (HAS_RUN - 1u)) + (int)(h - S.cnt[i]);
                        const uint32_t gb = bpos[j];
                        const int br = read_of(a.bmap, gb);
                        const uint32_t bp = gb - a.bmap.boff[br];
                        const unsigned long long o = tile_hits + h;
                        if (packed_out) a.keys[o] = ((((uint64_t)br << (a.pbits + a.abits)) | S.u.klow[i]) << a.qbits) | bp;
                        else { a.keys[o] = ((uint64_t)br << (a.pbits + a.abits)) | S.u.klow[i]; a.bpos[o] = bp; }
                    }
                }
                tile_hits += total;
            } else {
                // ================= one oversized run (or window), straight from global memory =================
                const uint64_t c = a.A.code(p);
                __syncthreads();
                if (tid == 0) {
                    int64_t jb, ib;
                    if (a.same_list) { jb = p; ib = q; }
                    else { jb = bound_code<false>(a.B, wb, we, c); ib = bound_code<true>(a.B, jb, we, c); }
                    S.seg[0] = jb; S.seg[1] = ib;
                }
                __syncthreads();
                const int64_t jb = S.seg[0], ib = S.seg[1], nbr = ib - jb, nA = q - p;
                if (nbr > 0) {
                    const bool lonely = a.same_list && a.self && nbr == 1;
                    auto below = [&](int64_t lo, int64_t hi, uint32_t thr) -> int64_t {      // first run entry in [lo, hi) with gpos >= thr
                        while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (a.B.epos(a.B.ent[m], m) < thr) lo = m + 1; else hi = m; }
                        return lo;
                    };
                    auto partners = [&](int64_t i) -> unsigned long long {     // in the whole run
                        if (lonely) return 0ull;
                        if (!a.self) return (unsigned long long)nbr;
                        return (unsigned long long)(below(jb, ib, self_threshold(a, a.A.epos(a.A.ent[i], i))) - jb);
                    };
                    // the partners that count for this chunk: [*first, *first + result)
                    auto chunk_partners = [&](int64_t i, int64_t *first) -> unsigned long long {
                        *first = jb;
                        if (lonely) return 0ull;
                        const int64_t pe = a.self ? below(jb, ib, self_threshold(a, a.A.epos(a.A.ent[i], i))) : ib;
                        if (!a.ranged) return (unsigned long long)(pe - jb);
                        const int64_t jr = below(jb, pe, a.g_lo);
                        *first = jr;
                        return (unsigned long long)(below(jr, pe, a.g_hi) - jr);
                    };
                    unsigned long long ct;
                    if (!a.self) ct = (unsigned long long)nA * (unsigned long long)nbr;
                    else {
                        unsigned long long mine = 0;
                        for (int64_t i = p + tid; i < q; i += MT_THREADS) mine += partners(i);
                        ct = block_sum(mine, S.red);
                    }
                    unsigned long long hits_of_run = a.self ? ct : (lonely ? 0ull : ct);
                    if (a.ranged && !EMIT) {
                        unsigned long long mine = 0; int64_t f;
                        for (int64_t i = p + tid; i < q; i += MT_THREADS) mine += chunk_partners(i, &f);
                        hits_of_run = block_sum(mine, S.red);
                    }
                    const uint32_t ct32 = ct > 0xffffffffull ? 0xffffffffu : (uint32_t)ct;
                    if (histo && tid == 0 && ct32 < MAXGRAM) atomicAdd(&a.gram[ct32], 1ull);
                    if (ct32 < a.limit) {
                        if (EMIT) {
                            for (int64_t base = p; base < q; base += MT_THREADS) {      // entries of the run, a block at a time
                                const int64_t i = base + tid;
                                int64_t jr = jb;
                                const unsigned long long cn = (i < q) ? chunk_partners(i, &jr) : 0ull;
                                unsigned long long inc = cn;
#pragma unroll
                                for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                                __syncthreads();
                                if (lane == 31) S.red[w] = inc;
                                __syncthreads();
                                unsigned long long off = 0, tot = 0;
                                for (int x = 0; x < MT_THREADS / 32; x++) { if (x < w) off += S.red[x]; tot += S.red[x]; }
                                if (cn) {
                                    const uint32_t ga = a.A.epos(a.A.ent[i], i);
                                    const int ar = read_of(a.amap, ga);
                                    const uint64_t klow = ((uint64_t)ar << a.pbits) | (uint64_t)(ga - a.amap.boff[ar]);
                                    unsigned long long o = tile_hits + off + inc - cn;
                                    for (unsigned long long x = 0; x < cn; x++, o++) {
                                        const uint32_t gb = a.B.epos(a.B.ent[jr + (int64_t)x], jr + (int64_t)x);
                                        const int br = read_of(a.bmap, gb);
                                        const uint32_t bp = gb - a.bmap.boff[br];
                                        if (packed_out) a.keys[o] = ((((uint64_t)br << (a.pbits + a.abits)) | klow) << a.qbits) | bp;
                                        else { a.keys[o] = ((uint64_t)br << (a.pbits + a.abits)) | klow; a.bpos[o] = bp; }
                                    }
                                }
                                tile_hits += tot;
                            }
                        } else tile_hits += hits_of_run;
                    }
                }
            }
            p = q; wb = we;
        }
        if (!EMIT && tid == 0) a.tcount[tile] = tile_hits;
    }
    if (histo) {
        __syncthreads();
        for (int i = tid; i < MT_SGRAM; i += MT_THREADS)
            if (S.u.gram[i]) atomicAdd(&a.gram[i], (unsigned long long)S.u.gram[i]);
    }
}

// exclusive scan of m 64-bit counts in place, total to *total (one CTA: m is the number of tiles)
__global__ void __launch_bounds__(1024) scan_counts(unsigned long long *v, uint32_t m, unsigned long long *total) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < m; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long x = (i < m) ? v[i] : 0ull;
        unsigned long long inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        unsigned long long off = 0, tot = 0;
        for (int x2 = 0; x2 < 32; x2++) { if (x2 < w) off += wsum[x2]; tot += wsum[x2]; }
        const unsigned long long carry = carry_s;
        if (i < m) v[i] = carry + off + inc - x;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

int bits_for(int64_t maxval_exclusive) {   // smallest b with 2^b >= maxval_exclusive (>= 1)
    int b = 1;
    while ((1ll << b) < maxval_exclusive) b++;
    return b;
}

}  // namespace

// chunk == nullptr: the whole comparison; returns 2 without emitting anything (out.total and out.limit set) when it has more
// than max_hits hits.  chunk != nullptr: the hits whose B position lies in [g_lo, g_hi), under the limit the whole
// comparison's count found; returns 2 when even the chunk has more than max_hits.
int seed_hits(const DeviceIndex *ai, const DeviceIndex *bi, int self, int comp, SeedHits &out, const HitChunk *chunk, int64_t max_hits) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const DeviceBlock *ab = ai->blk, *bb = bi->blk;
    const int64_t alen = ai->len, blen = bi->len;
    const int same_list = (ai == bi) ? 1 : 0;
    out.release();
    out.limit = INT32_MAX; out.total = 0;

    HostProf *hp1 = new HostProf("sh.count");
    Timer t_count(st);
    const uint32_t ntiles = (uint32_t)((alen + MT_TA - 1) / MT_TA);
    DevBuf<int64_t> tbounds; tbounds.reserve(2 * ((size_t)ntiles + 1));
    DevBuf<unsigned long long> tcount; tcount.reserve((size_t)ntiles + 2);
    DevBuf<unsigned long long> gram;
    const bool use_limit = (c.opt.mem_limit > 0) && chunk == nullptr;
    if (use_limit) { gram.reserve(MAXGRAM); CUDA_CHECK(cudaMemsetAsync(gram.p, 0, sizeof(unsigned long long) * MAXGRAM, st)); }

    MergeArgs m{};
    m.A = KList{ai->codes, ai->gpos, alen, ai->gbits};
    m.B = KList{bi->codes, bi->gpos, blen, bi->gbits};
    m.amap = ReadMap{ab->boff, ab->rtab}; m.bmap = ReadMap{bb->boff, bb->rtab};
    m.self = self; m.comp = comp; m.identity = c.opt.identity; m.same_list = same_list;
    m.ntiles = ntiles; m.tstart = tbounds.p; m.twin = tbounds.p + ntiles + 1; m.tcount = tcount.p;
    m.gram = use_limit ? gram.p : nullptr;
    // the -M rule yields MAXGRAM whenever the hits fit the allowance (filter.c:2650-2656): count with that limit first
    m.limit = use_limit ? (uint32_t)MAXGRAM : (uint32_t)INT32_MAX;
    if (chunk) { m.ranged = 1; m.g_lo = chunk->g_lo; m.g_hi = chunk->g_hi; m.limit = (uint32_t)chunk->limit; }

    static std::once_flag attr;
    std::call_once(attr, [] {
        CUDA_CHECK(cudaFuncSetAttribute(merge_tiles<false, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MtSmem<uint32_t, false>)));
        CUDA_CHECK(cudaFuncSetAttribute(merge_tiles<true, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MtSmem<uint32_t, true>)));
        CUDA_CHECK(cudaFuncSetAttribute(merge_tiles<false, uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MtSmem<uint64_t, false>)));
        CUDA_CHECK(cudaFuncSetAttribute(merge_tiles<true, uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MtSmem<uint64_t, true>)));
    });
    const bool code32 = c.fp.kmer <= 16;                    // a k-mer code has 2k bits
    auto count_sweep = [&](const MergeArgs &mm, unsigned g) {
        if (code32) merge_tiles<false, uint32_t><<<g, MT_THREADS, sizeof(MtSmem<uint32_t, false>), st>>>(mm);
        else merge_tiles<false, uint64_t><<<g, MT_THREADS, sizeof(MtSmem<uint64_t, false>), st>>>(mm);
        KERNEL_CHECK();
    };
    const unsigned grid = ntiles < 3u * NUM_SMS_B200 ? ntiles : 3u * NUM_SMS_B200;
    merge_bounds<<<(ntiles + 1 + 127) / 128, 128, 0, st>>>(m);
    KERNEL_CHECK();
    count_sweep(m, grid);
    scan_counts<<<1, 1024, 0, st>>>(tcount.p, ntiles, tcount.p + ntiles);
    KERNEL_CHECK();
    count_launch(3);

    // the limit (filter.c:2634-2699), evaluated on the host from the histogram like the reference's main thread
    unsigned long long nhits = 0;
    int64_t limit = chunk ? chunk->limit : INT32_MAX;
    std::vector<unsigned long long> histo;
    if (use_limit) { histo.resize(MAXGRAM); CUDA_CHECK(cudaMemcpyAsync(histo.data(), gram.p, sizeof(unsigned long long) * MAXGRAM, cudaMemcpyDeviceToHost, st)); }
    CUDA_CHECK(cudaMemcpyAsync(&nhits, tcount.p + ntiles, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    if (use_limit) {
        int64_t avail = (int64_t)(((int64_t)(c.opt.mem_limit - (uint64_t)(ab->sizeof_db + bb->sizeof_db))) / (uint64_t)16);
        if (same_list || avail > alen + 2 * blen) avail = (avail - alen) / 2;
        else avail = avail - (alen + blen);
        avail *= .98;
        int64_t tom = 0; int j;
        for (j = 0; j < MAXGRAM; j++) { tom += j * (int64_t)histo[j]; if (tom > avail) break; }
        limit = j;
        if (limit <= 1) { set_error("Insufficient memory allocation (%.1fGb), reduce block size or increase allocation", (1. * c.opt.mem_limit) / 0x40000000ll); delete hp1; return 1; }
        if (limit < 10) fprintf(stderr, "\nWarning: Sensitivity hampered by low memory allocation (%.1fGb), reduce block size or increase allocation\n", (1. * c.opt.mem_limit) / 0x40000000ll);
        if (c.opt.verbose) printf("   Capping mutual k-mer matches over %d (effectively -t%d)\n", (int)limit, (int)__builtin_sqrt(1. * limit));
        if (limit != MAXGRAM) {                         // a tighter cap than assumed: count again with it
            m.limit = (uint32_t)limit; m.gram = nullptr;
            count_sweep(m, grid);
            scan_counts<<<1, 1024, 0, st>>>(tcount.p, ntiles, tcount.p + ntiles);
            KERNEL_CHECK();
            count_launch(2);
            CUDA_CHECK(cudaMemcpyAsync(&nhits, tcount.p + ntiles, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(stream_wait(st));
        }
    }
    out.limit = (int32_t)limit; out.total = (int64_t)nhits;
    c.tm.count_ms = t_count.stop();
    delete hp1;
    if (c.opt.verbose && chunk == nullptr) { printf("   Hit count = %lld\n", (long long)nhits); fflush(stdout); }
    if ((int64_t)nhits > max_hits) return 2;

    HostProf *hp2 = new HostProf("sh.merge");
    Timer t_merge(st);
    if (nhits == 0) { c.tm.merge_ms = t_merge.stop(); delete hp2; return 0; }

    out.pbits = bits_for(ab->maxlen);
    out.abits = bits_for(ab->nreads);
    out.bbits = bits_for(bb->nreads);
    out.qbits = bits_for(bb->maxlen);
    const int kbits = out.pbits + out.abits + out.bbits;
    if (kbits > 64) { set_error("seed-hit key needs %d bits (> 64): block too large for this build", kbits); delete hp2; return 1; }
    static const bool pack_ok = !(getenv("DAMAR_PACK") && atoi(getenv("DAMAR_PACK")) == 0);
    const bool packed = pack_ok && kbits + out.qbits <= 64;

    uint64_t *k0, *k1; uint32_t *v0 = nullptr, *v1 = nullptr;
    FreeOnUnwind raw;
    k0 = raw.add((uint64_t *)dev_alloc_bytes(sizeof(uint64_t) * (nhits + 2)));
    k1 = raw.add((uint64_t *)dev_alloc_bytes(sizeof(uint64_t) * (nhits + 2)));
    if (!packed) {
        v0 = raw.add((uint32_t *)dev_alloc_bytes(sizeof(uint32_t) * (nhits + 4)));
        v1 = raw.add((uint32_t *)dev_alloc_bytes(sizeof(uint32_t) * (nhits + 4)));
    }
    m.keys = k0; m.bpos = v0; m.pbits = out.pbits; m.abits = out.abits; m.qbits = out.qbits;
    m.limit = (uint32_t)limit; m.gram = nullptr;
    if (code32) merge_tiles<true, uint32_t><<<grid, MT_THREADS, sizeof(MtSmem<uint32_t, true>), st>>>(m);
    else merge_tiles<true, uint64_t><<<grid, MT_THREADS, sizeof(MtSmem<uint64_t, true>), st>>>(m);
    KERNEL_CHECK(); count_launch();
    c.tm.merge_ms = t_merge.stop();
    delete hp2;

    HostProf hp3("sh.sort+free");
    Timer t_sort(st);
    const int rc = packed ? radix_sort(k0, v0, k1, v1, (int64_t)nhits, out.qbits, out.qbits + kbits, st)
                          : radix_sort(k0, v0, k1, v1, (int64_t)nhits, 0, kbits, st);
    c.tm.hsort_ms = t_sort.stop();
    raw.dismiss();
    dev_free(k1); if (v1) dev_free(v1);
    if (rc != 0) { dev_free(k0); if (v0) dev_free(v0); return 1; }
    out.keys = k0; out.bpos = v0; out.n = (int64_t)nhits;
    return 0;
}

}  // namespace damar
