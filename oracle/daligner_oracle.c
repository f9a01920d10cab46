/*
 * daligner_oracle.c -- TEST INFRASTRUCTURE ONLY (see daligner_oracle.h).
 *
 * Single-threaded CPU restatement of the daligner overlap path of MartinPippel/DAmar.
 * Every function cites the reference lines it restates (paths relative to the reference
 * root).  It is pinned against the reference binary (oracle/_ref/daligner); it is never
 * linked into, imported by, or executed from the product path.
 *
 * Known, documented deviations (all are undefined/racy behaviour in the reference):
 *   - score[d+-1] for the extreme diagonal buckets reads 0 here (reference aliases the
 *     neighbouring array, filter.c:2138-2139, 2289, 2828-2832).
 *   - Handle_Redundancies (filter.c:1804) is not restated yet: pairs with >1 surviving
 *     local alignment are emitted unfused and counted in ora_result.multi_la_pairs.
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <limits.h>
#include "daligner_oracle.h"

#define MAXGRAM        10000      /* filter.c:71 */
#define PANEL_SIZE     50000      /* filter.c:73 */
#define PANEL_OVERLAP  10000      /* filter.c:74 */
#define TRACE_XOVR     125        /* align.h:58  */

#define TRIM_LEN    15            /* align.c:188 */
#define PATH_LEN    60            /* align.c:192 */
#define PATH_TOP    0x1000000000000000ull
#define PATH_INT    0x0fffffffffffffffull
#define TRIM_MASK   0x7fff
#define TRIM_MLAG   200
#define WAVE_LAG    30
#define FRACTION    1000

static void *xmalloc(size_t n)
{
    void *p = malloc(n ? n : 1);
    if (p == NULL) { fprintf(stderr, "oracle: out of memory (%zu bytes)\n", n); exit(1); }
    return p;
}
static void *xrealloc(void *q, size_t n)
{
    void *p = realloc(q, n ? n : 1);
    if (p == NULL) { fprintf(stderr, "oracle: out of memory (%zu bytes)\n", n); exit(1); }
    return p;
}
void ora_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * Stable LSD radix sort of 16-byte records on a 64-bit key (16-bit digits).  Restates the
 * *contract* of lex_sort (filter.c:328-435): stable, ascending on the flagged key bytes.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t lo, hi; } rec16;
typedef uint64_t (*keyfn)(const rec16 *);

static rec16 *radix_sort64(rec16 *src, rec16 *tmp, int64_t n, keyfn key, int nbits)
{
    int64_t *cnt = (int64_t *)xmalloc(sizeof(int64_t) * 65536);
    int shift;
    for (shift = 0; shift < nbits; shift += 16) {
        int64_t i, sum = 0, c;
        memset(cnt, 0, sizeof(int64_t) * 65536);
        for (i = 0; i < n; i++) cnt[(key(src + i) >> shift) & 0xffff]++;
        for (i = 0; i < 65536; i++) { c = cnt[i]; cnt[i] = sum; sum += c; }
        for (i = 0; i < n; i++) tmp[cnt[(key(src + i) >> shift) & 0xffff]++] = src[i];
        { rec16 *x = src; src = tmp; tmp = x; }
    }
    free(cnt);
    return src;
}

static uint64_t key_code(const rec16 *r) { return r->lo; }                       /* KmerPos.code        */
static uint64_t key_apos(const rec16 *r) { return r->lo >> 32; }                 /* SeedPair.apos       */
static uint64_t key_pair(const rec16 *r) { return r->hi; }                       /* bread<<32 | aread   */

/* ------------------------------------------------------------------------------------------
 * Stage 1.  Sort_Kmers (filter.c:753-994) with tuple_thread (filter.c:458-547) and the -t
 * compaction (filter.c:700-751, 890-939).  Result: every admissible k-mer, stably sorted by
 * code (ties in (read,rpos) order), followed by the sentinels {~0},{0} (filter.c:941-942).
 * ------------------------------------------------------------------------------------------ */
/* biased_tuple_thread (filter.c:549-688) for one stretch [p0, q) of a read: variable-length k-mers whose summed
 * -log4 frequency reaches Kmer "bases" of information; `masked` selects the interval-gap variant (:571-631), whose
 * inner growth loop stops at q WITHOUT leaving the outer step, else the whole-read variant (:633-676), which
 * leaves the read at its terminator.  Emits into list (may be NULL: count only); returns the count.             */
static int64_t biased_stretch(const int8_t *s, int p0, int q, int masked, int K, uint64_t kmask, const int *LogBase,
                              int LogNorm, int LogThresh, ora_kmer *list, int read)
{
    const int8_t *t = s + 1;
    uint64_t c = 0; int a = 0, k = 1, p = p0, x; int64_t n = 0;
    const int Kshift = 2 * K;
    while (masked ? (p < q) : (s[p] != 4)) {
        x = s[p];
        a += LogBase[x]; c = (c << 2) | (uint64_t)x;
        while (a < LogNorm && k < K) {
            if (masked) { if (++p >= q) break; x = s[p]; }
            else { if ((x = s[++p]) == 4) return n; }
            k += 1; a += LogBase[x]; c = (c << 2) | (uint64_t)x;
        }
        for (;;) { int u = a - LogBase[(int)t[p - k]]; if (u < LogNorm) break; a = u; k -= 1; }
        if (a > LogThresh) {
            if (list) { list[n].code = (c << (Kshift - 2 * k)) & kmask; list[n].read = read; list[n].rpos = p; }
            n++;
        }
        p += 1;
        a -= LogBase[(int)s[p - k]];
    }
    return n;
}

ora_kmer *ora_sort_kmers(const ora_block *blk, const ora_params *p, int64_t *len)
{
    const int K = p->kmer;
    const uint64_t kmask = (K == 32) ? ~0ull : ((1ull << (2 * K)) - 1);   /* filter.c:182-186 */
    int64_t total = 0, n = 0;
    int i;
    ora_kmer *list, *tmp, *rez;
    int LogBase[4] = {0, 0, 0, 0}, LogNorm = 0, LogThresh = 0;

    if (p->biased) {                               /* filter.c:774-789; MAX_BIAS = 2 (:69).  The reference fills LogBase
                                                      ONCE, at the first Sort_Kmers call of the process = the forward A block
                                                      (daligner.c:982), and reuses it for every later block and orientation. */
        const float *fq = p->freq;
        double scale = -10000. / log(4.);
        LogNorm = 10000 * K; LogThresh = 10000 * (K - 2);
        for (i = 0; i < 4; i++) LogBase[i] = (int)ceil(scale * log(fq[i]));
    }

    for (i = 0; i < blk->nreads; i++) {
        const int8_t *s = blk->bases + blk->boff[i];
        if (blk->manno == NULL) {
            if (p->biased) total += biased_stretch(s, 0, blk->rlen[i], 0, K, kmask, LogBase, LogNorm, LogThresh, NULL, i);
            else if (blk->rlen[i] >= K) total += blk->rlen[i] - K + 1;
        } else {                                   /* gaps between mask intervals, filter.c:483-515 */
            int64_t b = blk->manno[i], f = blk->manno[i + 1], a;
            for (a = b; a <= f; a += 2) {
                int pp = (a == b) ? 0 : blk->mdata[a - 1];
                int q  = (a == f) ? blk->rlen[i] : blk->mdata[a];
                if (pp + K <= q) {
                    if (p->biased) total += biased_stretch(s, pp, q, 1, K, kmask, LogBase, LogNorm, LogThresh, NULL, i);
                    else total += q - (pp + K - 1);
                }
            }
        }
    }
    *len = 0;
    if (total <= 0) return NULL;                   /* filter.c:800-801, 990-993 */

    list = (ora_kmer *)xmalloc(sizeof(ora_kmer) * (total + 2));
    tmp  = (ora_kmer *)xmalloc(sizeof(ora_kmer) * (total + 2));

    for (i = 0; i < blk->nreads; i++) {
        const int8_t *s = blk->bases + blk->boff[i];
        if (blk->manno == NULL) {                  /* filter.c:529-544 */
            uint64_t c = 0; int x, pos = 0;
            if (p->biased) { n += biased_stretch(s, 0, blk->rlen[i], 0, K, kmask, LogBase, LogNorm, LogThresh, list + n, i); continue; }
            if (blk->rlen[i] < K) continue;
            for (x = 1; x < K; x++) c = (c << 2) | (uint64_t)s[pos++];
            while (s[pos] != 4) {
                c = ((c << 2) | (uint64_t)s[pos]) & kmask;
                list[n].code = c; list[n].read = i; list[n].rpos = pos++; n++;
            }
        } else {
            int64_t b = blk->manno[i], f = blk->manno[i + 1], a;
            for (a = b; a <= f; a += 2) {
                int pos = (a == b) ? 0 : blk->mdata[a - 1];
                int q   = (a == f) ? blk->rlen[i] : blk->mdata[a];
                if (pos + K <= q) {
                    uint64_t c = 0; int x;
                    if (p->biased) { n += biased_stretch(s, pos, q, 1, K, kmask, LogBase, LogNorm, LogThresh, list + n, i); continue; }
                    for (x = 1; x < K; x++) c = (c << 2) | (uint64_t)s[pos++];
                    while (pos < q) {
                        c = ((c << 2) | (uint64_t)s[pos]) & kmask;
                        list[n].code = c; list[n].read = i; list[n].rpos = pos++; n++;
                    }
                }
            }
        }
    }

    rez = (ora_kmer *)radix_sort64((rec16 *)list, (rec16 *)tmp, n, key_code, 2 * K);

    if (p->suppress > 0) {                         /* -t: drop runs of length >= TooFrequent */
        ora_kmer *out = (rez == list) ? tmp : list;
        int64_t j = 0, m = 0;
        while (j < n) {
            int64_t e = j + 1;
            while (e < n && rez[e].code == rez[j].code) e++;
            if (e - j < p->suppress) { memcpy(out + m, rez + j, sizeof(ora_kmer) * (e - j)); m += e - j; }
            j = e;
        }
        rez = out; n = m;
    }
    if (rez == list) free(tmp); else free(list);
    if (n <= 0) { free(rez); return NULL; }
    rez[n].code = ~0ull;  rez[n].read = rez[n].rpos = 0;
    rez[n + 1].code = 0;  rez[n + 1].read = rez[n + 1].rpos = 0;
    *len = n;
    return rez;
}

/* ------------------------------------------------------------------------------------------
 * Stage 2.  count_thread (filter.c:1039-1165), limit choice (filter.c:2634-2702),
 * merge_thread (filter.c:1170-1358), seed-hit sort (filter.c:2561-2580, 2776-2781).
 * ------------------------------------------------------------------------------------------ */

/* number of b-entries of a code group admitted for a-entry `a` in a self comparison;
 * *b is the running cursor (monotone over the a-entries of the group)                     */
static inline void self_cursor(const ora_kmer *al, int64_t a, const ora_kmer *bl, int64_t *b,
                               int64_t ib, int identity, int comp)
{
    int ar = al[a].read;
    if (identity) {
        if (comp) {                                   /* filter.c:1092-1096 */
            while (*b < ib && bl[*b].read <= ar) (*b)++;
        } else {                                      /* filter.c:1097-1104 */
            int ap = al[a].rpos;
            while (*b < ib && bl[*b].read < ar) (*b)++;
            while (*b < ib && bl[*b].read == ar && bl[*b].rpos < ap) (*b)++;
        }
    } else                                            /* filter.c:1108-1112 */
        while (*b < ib && bl[*b].read < ar) (*b)++;
}

ora_hit *ora_seed_hits(const ora_block *ab, const ora_block *bb,
                       const ora_kmer *al, int64_t alen, const ora_kmer *bl, int64_t blen,
                       int self, int comp, const ora_params *p, int64_t *nhits_out, int32_t *limit_out)
{
    int64_t *gram = (int64_t *)xmalloc(sizeof(int64_t) * MAXGRAM);
    int64_t ia, ib, nhits = 0, n = 0;
    int64_t limit;
    ora_hit *hits, *tmp, *rez;
    int pass;

    *nhits_out = 0; *limit_out = INT32_MAX;
    if (alen == 0 || blen == 0) { free(gram); return NULL; }
    memset(gram, 0, sizeof(int64_t) * MAXGRAM);

    limit = INT32_MAX;
    hits = tmp = NULL;
    for (pass = 0; pass < 2; pass++) {
        ia = ib = 0;
        while (ia < alen && ib < blen) {
            uint64_t ca = al[ia].code, cb = bl[ib].code;
            if (ca < cb) { ia++; continue; }
            if (cb < ca) { ib++; continue; }
            {
                int64_t ja = ia, jb = ib, a, b, c, ct;
                while (ia < alen && al[ia].code == ca) ia++;
                while (ib < blen && bl[ib].code == cb) ib++;
                if (self) {
                    ct = 0; b = jb;
                    for (a = ja; a < ia; a++) { self_cursor(al, a, bl, &b, ib, p->identity, comp); ct += b - jb; }
                } else
                    ct = (ia - ja) * (ib - jb);
                if (pass == 0) {
                    nhits += ct;
                    if (ct < MAXGRAM) gram[ct]++;
                } else if (ct < limit) {
                    if (self) {
                        b = jb;
                        for (a = ja; a < ia; a++) {
                            self_cursor(al, a, bl, &b, ib, p->identity, comp);
                            for (c = jb; c < b; c++) {
                                hits[n].bread = bl[c].read; hits[n].aread = al[a].read;
                                hits[n].apos = al[a].rpos;  hits[n].diag = al[a].rpos - bl[c].rpos; n++;
                            }
                        }
                    } else
                        for (a = ja; a < ia; a++)
                            for (b = jb; b < ib; b++) {
                                hits[n].bread = bl[b].read; hits[n].aread = al[a].read;
                                hits[n].apos = al[a].rpos;  hits[n].diag = al[a].rpos - bl[b].rpos; n++;
                            }
                }
            }
        }
        if (pass == 0) {
            if (p->mem_limit > 0) {                /* filter.c:2634-2699 */
                int64_t tom = 0, avail; int j;
                avail = (int64_t)(((int64_t)(p->mem_limit - (uint64_t)(p->sizeof_a + p->sizeof_b))) / sizeof(rec16));
                if (al == bl || avail > alen + 2 * blen) avail = (avail - alen) / 2;
                else avail = avail - (alen + blen);
                avail *= .98;
                for (j = 0; j < MAXGRAM; j++) { tom += j * gram[j]; if (tom > avail) break; }
                limit = j;
                if (limit <= 1) { fprintf(stderr, "oracle: insufficient memory allocation (limit<=1)\n"); exit(1); }
                nhits = 0;
                for (j = 1; j < limit; j++) nhits += j * gram[j];
            }
            if (nhits == 0) { free(gram); *limit_out = (int32_t)limit; return NULL; }
            hits = (ora_hit *)xmalloc(sizeof(ora_hit) * (nhits + 1));
            tmp  = (ora_hit *)xmalloc(sizeof(ora_hit) * (nhits + 1));
        }
    }
    free(gram);
    if (n != nhits) { fprintf(stderr, "oracle: internal error, merge emitted %lld != %lld\n", (long long)n, (long long)nhits); exit(1); }

    /* (bread, aread, apos) ascending, stable => merge order survives inside equal keys */
    {
        int nb = 0; int64_t pw;
        for (pw = 1; pw < ab->maxlen; pw <<= 8) nb += 8;
        rez = (ora_hit *)radix_sort64((rec16 *)hits, (rec16 *)tmp, n, key_apos, nb ? nb : 16);
        rez = (ora_hit *)radix_sort64((rec16 *)rez, (rec16 *)(rez == hits ? tmp : hits), n, key_pair, 64);
    }
    if (rez == hits) free(tmp); else free(hits);
    rez[n].aread = 0x7fffffff; rez[n].bread = 0x7fffffff; rez[n].diag = 0x7fffffff; rez[n].apos = 0;
    (void)bb;
    *nhits_out = n; *limit_out = (int32_t)limit;
    return rez;
}

/* ------------------------------------------------------------------------------------------
 * Stage 4.  Local_Alignment (align.c:1904-2097) = forward_wave (align.c:409-1122) +
 * reverse_wave (align.c:1126-1898) + the B-path wrapper.
 * ------------------------------------------------------------------------------------------ */
typedef struct { int32_t prev, diag, diff, mark; } pebble;        /* align.c:399-405 */
typedef struct { int32_t abpos, bbpos, aepos, bepos, diffs, tlen; uint16_t *trace; } opath;

typedef struct {
    int16_t *score, *table;            /* align.c:286-297 */
    int      ave_path, tspace, reach;
    int32_t *V, *M, *HA, *HB, *NA, *NB; uint64_t *T;  int vcap;
    pebble  *cells; int ccap, ctop;
    uint16_t *tbuf; int tcap;
    int64_t  wave_cells, waves;
} la_ctx;

/* New_Align_Spec + set_table (align.c:226-318); FP exactly as the reference evaluates it */
static double Bias_Factor[10] = { .690, .690, .690, .690, .780, .850, .900, .933, .966, 1.000 };

static la_ctx *new_la_ctx(const ora_params *p)
{
    la_ctx *c = (la_ctx *)xmalloc(sizeof(la_ctx));
    double match; int bias, mscore, dscore, x;
    memset(c, 0, sizeof(*c));
    match = p->freq[0] + p->freq[3];
    if (match > .5) match = 1. - match;
    bias = (int)((match + .025) * 20. - 1.);
    if (match < .2) bias = 3;
    c->ave_path = (int)(PATH_LEN * (1. - Bias_Factor[bias] * (1. - p->ave_corr)));
    mscore = (int)(FRACTION * Bias_Factor[bias] * (1. - p->ave_corr));
    dscore = FRACTION - mscore;
    c->score = (int16_t *)xmalloc(sizeof(int16_t) * (TRIM_MASK + 1) * 2);
    c->table = c->score + (TRIM_MASK + 1);
    for (x = 0; x <= TRIM_MASK; x++) {            /* oldest column = bit 14, newest = bit 0 */
        int s = 0, mx = 0, bit;
        for (bit = TRIM_LEN - 1; bit >= 0; bit--) {
            if (s > mx) mx = s;
            s += ((x >> bit) & 1) ? mscore : -dscore;
        }
        c->table[x] = (int16_t)(s - mx);
        c->score[x] = (int16_t)s;
    }
    c->tspace = p->tspace;
    c->reach = 1;                                 /* daligner.c:951 */
    return c;
}
static void free_la_ctx(la_ctx *c)
{
    free(c->score); free(c->V); free(c->M); free(c->HA); free(c->HB); free(c->NA); free(c->NB);
    free(c->T); free(c->cells); free(c->tbuf); free(c);
}

static inline int peb_push(la_ctx *c, int prev, int diag, int diff, int mark)
{
    if (c->ctop >= c->ccap) {
        c->ccap = (int)(c->ccap * 1.5) + 4096;
        c->cells = (pebble *)xrealloc(c->cells, sizeof(pebble) * c->ccap);
    }
    c->cells[c->ctop].prev = prev; c->cells[c->ctop].diag = diag;
    c->cells[c->ctop].diff = diff; c->cells[c->ctop].mark = mark;
    return c->ctop++;
}
/* reverse a pebble chain in place, return its root (align.c:1018-1025) */
static inline int peb_reverse(pebble *cells, int h)
{
    int a = -1, b;
    for (; h >= 0; h = b) { b = cells[h].prev; cells[h].prev = a; a = h; }
    return a;
}

/* forward_wave, align.c:409-1122.  A/B are the read starts (index 0 = first base).       */
static void forward_wave(la_ctx *c, const int8_t *A, const int8_t *B, int *mind, int maxd, int mida,
                         int minp, int maxp, int aoff, int boff, opath *apath, opath *bpath, int voff)
{
    int32_t *V = c->V + voff, *M = c->M + voff, *HA = c->HA + voff, *HB = c->HB + voff;
    int32_t *NA = c->NA + voff, *NB = c->NB + voff; uint64_t *T = c->T + voff;
    pebble *cells;
    const int TS = c->tspace, PATH_AVE = c->ave_path;
    int hgh = maxd, low = *mind, dif = 0;
    int besta, besty, trima, trimy, trimd, trimha, trimhb;
    int morea, morey, mored, moreha, morehb, more, morem, lasta, aclip, bclip;
    int k;

    more = 1; aclip = INT32_MAX; bclip = -INT32_MAX;
    besta = trima = morea = lasta = mida;
    besty = trimy = morey = (mida - hgh) >> 1;
    trimd = mored = 0; trimha = moreha = 0; trimhb = morehb = 1; morem = -1;
    c->ctop = 0;

    for (k = hgh; k >= low; k--) {                            /* wave 0, align.c:491-626 */
        int y = (mida - k) >> 1, cc, ha, hb, na, nb;
        na = (((y + k) + (TS - aoff)) / TS - 1) * TS + aoff;
        ha = peb_push(c, -1, k, 0, na); na += TS;
        nb = ((y + (TS - boff)) / TS - 1) * TS + boff;
        hb = peb_push(c, -1, k, 0, nb); nb += TS;
        for (;;) {
            int cb = B[y], ca;
            if (cb == 4) { more = 0; if (bclip < k) bclip = k; break; }
            ca = A[y + k];
            if (cb != ca) { if (ca == 4) { more = 0; aclip = k; } break; }
            y += 1;
        }
        cc = (y << 1) + k;
        while (y + k >= na) { ha = peb_push(c, ha, k, 0, na); na += TS; }
        while (y >= nb)     { hb = peb_push(c, hb, k, 0, nb); nb += TS; }
        if (cc > besta) { besta = trima = lasta = cc; besty = trimy = y; trimha = ha; trimhb = hb; }
        V[k] = cc; T[k] = PATH_INT; M[k] = PATH_LEN; HA[k] = ha; HB[k] = hb; NA[k] = na; NB[k] = nb;
    }

    for (;;) {
        if (more == 0) {                                      /* align.c:628-658, 943-975 */
            if (B[besty] != 4 && A[besta - besty] != 4) more = 1;
            if (hgh >= aclip) {
                hgh = aclip - 1;
                if (morem <= M[aclip]) { morem = M[aclip]; morea = V[aclip]; morey = (morea - aclip) / 2;
                                         mored = dif; moreha = HA[aclip]; morehb = HB[aclip]; }
            }
            if (low <= bclip) {
                low = bclip + 1;
                if (morem <= M[bclip]) { morem = M[bclip]; morea = V[bclip]; morey = (morea - bclip) / 2;
                                         mored = dif; moreha = HA[bclip]; morehb = HB[bclip]; }
            }
            aclip = INT32_MAX; bclip = -INT32_MAX;
        }
        if (dif > 0) {                                        /* prune, align.c:977-986 */
            int n = besta - WAVE_LAG;
            while (hgh >= low)
                if (V[hgh] < n) hgh -= 1;
                else { while (V[low] < n) low += 1; break; }
            c->waves++; c->wave_cells += (hgh - low) + 1;
        }
        if (!(more && lasta >= besta - TRIM_MLAG)) break;     /* align.c:667 */

        {
            int am, ac, ap, n, ua, ub; uint64_t t;
            low -= 1; hgh += 1;
            if (low >= minp) { NA[low] = NA[low + 1]; NB[low] = NB[low + 1]; V[low] = -1; }
            else low += 1;
            if (hgh <= maxp) { NA[hgh] = NA[hgh - 1]; NB[hgh] = NB[hgh - 1]; V[hgh] = am = -1; }
            else am = V[--hgh];
            dif += 1;
            ac = V[hgh + 1] = V[low - 1] = -1;
            t = PATH_INT; n = PATH_LEN; ua = ub = -1;
            for (k = hgh; k >= low; k--) {
                int y, m, ha, hb, cc; uint64_t b;
                ap = ac; ac = am; am = V[k - 1];
                if (ac < am) {
                    if (am < ap) { cc = ap + 1; m = n; b = t; ha = ua; hb = ub; }
                    else         { cc = am + 1; m = M[k - 1]; b = T[k - 1]; ha = HA[k - 1]; hb = HB[k - 1]; }
                } else if (ac < ap) { cc = ap + 1; m = n; b = t; ha = ua; hb = ub; }
                else                { cc = ac + 2; m = M[k]; b = T[k]; ha = HA[k]; hb = HB[k]; }

                if ((b & PATH_TOP) != 0) m -= 1;
                b <<= 1;
                y = (cc - k) >> 1;
                for (;;) {
                    int cb = B[y], ca;
                    if (cb == 4) { more = 0; if (bclip < k) bclip = k; break; }
                    ca = A[y + k];
                    if (cb != ca) { if (ca == 4) { more = 0; aclip = k; } break; }
                    y += 1;
                    if ((b & PATH_TOP) == 0) m += 1;
                    b = (b << 1) | 1;
                }
                cc = (y << 1) + k;

                while (y + k >= NA[k]) {
                    if (c->cells[ha].mark < NA[k]) ha = peb_push(c, ha, k, dif, NA[k]);
                    NA[k] += TS;
                }
                while (y >= NB[k]) {
                    if (c->cells[hb].mark < NB[k]) hb = peb_push(c, hb, k, dif, NB[k]);
                    NB[k] += TS;
                }
                if (cc > besta) {
                    besta = cc; besty = y;
                    if (m >= PATH_AVE) {
                        lasta = cc;
                        if (c->table[b & TRIM_MASK] >= 0)
                            if (c->table[(b >> TRIM_LEN) & TRIM_MASK] + c->score[b & TRIM_MASK] >= 0) {
                                trima = cc; trimy = y; trimd = dif; trimha = ha; trimhb = hb;
                            }
                    }
                }
                t = T[k]; n = M[k]; ua = HA[k]; ub = HB[k];
                V[k] = cc; T[k] = b; M[k] = m; HA[k] = ha; HB[k] = hb;
            }
        }
    }

    {                                                         /* finish, align.c:1001-1118 */
        uint16_t *atrace = apath->trace, *btrace = bpath->trace;
        int atlen = 0, btlen = 0, trimx, a, b, h, d, e;
        cells = c->cells;
        if (morem >= 0 && c->reach) { trimx = morea - morey; trimy = morey; trimd = mored; trimha = moreha; trimhb = morehb; }
        else trimx = trima - trimy;

        h = peb_reverse(cells, trimha);
        k = cells[h].diag; b = (mida - k) / 2; e = 0;
        for (h = cells[h].prev; h >= 0; h = cells[h].prev) {
            k = cells[h].diag; a = cells[h].mark - k; d = cells[h].diff;
            atrace[atlen++] = (uint16_t)(d - e); atrace[atlen++] = (uint16_t)(a - b);
            b = a; e = d;
        }
        if (b + k != trimx) { atrace[atlen++] = (uint16_t)(trimd - e); atrace[atlen++] = (uint16_t)(trimy - b); }
        else if (b != trimy) {
            atrace[atlen - 1] = (uint16_t)(atrace[atlen - 1] + (trimy - b));
            atrace[atlen - 2] = (uint16_t)(atrace[atlen - 2] + (trimd - e));
        }

        h = peb_reverse(cells, trimhb);
        k = cells[h].diag; b = (mida + k) / 2; e = 0; low = k;
        for (h = cells[h].prev; h >= 0; h = cells[h].prev) {
            k = cells[h].diag; a = cells[h].mark + k; d = cells[h].diff;
            btrace[btlen++] = (uint16_t)(d - e); btrace[btlen++] = (uint16_t)(a - b);
            b = a; e = d;
        }
        if (b - k != trimy) { btrace[btlen++] = (uint16_t)(trimd - e); btrace[btlen++] = (uint16_t)(trimx - b); }
        else if (b != trimx) {
            btrace[btlen - 1] = (uint16_t)(btrace[btlen - 1] + (trimx - b));
            btrace[btlen - 2] = (uint16_t)(btrace[btlen - 2] + (trimd - e));
        }
        apath->aepos = trimx; apath->bepos = trimy; apath->diffs = trimd;
        apath->tlen = atlen; bpath->tlen = btlen;
    }
    *mind = low;
}

/* reverse_wave, align.c:1126-1898.  Bases are read one to the left of the coordinate.     */
static void reverse_wave(la_ctx *c, const int8_t *A0, const int8_t *B0, int mind, int maxd, int mida,
                         int minp, int maxp, int aoff, int boff, opath *apath, opath *bpath, int voff)
{
    const int8_t *A = A0 - 1, *B = B0 - 1;
    int32_t *V = c->V + voff, *M = c->M + voff, *HA = c->HA + voff, *HB = c->HB + voff;
    int32_t *NA = c->NA + voff, *NB = c->NB + voff; uint64_t *T = c->T + voff;
    pebble *cells;
    const int TS = c->tspace, PATH_AVE = c->ave_path;
    int hgh = maxd, low = mind, dif = 0;
    int besta, besty, trima, trimy, trimd, trimha, trimhb;
    int morea, morey, mored, moreha, morehb, more, morem, lasta, aclip, bclip;
    int k;

    more = 1; aclip = -INT32_MAX; bclip = INT32_MAX;
    besta = trima = morea = lasta = mida;
    besty = trimy = morey = (mida - hgh) >> 1;
    trimd = mored = 0; trimha = moreha = 0; trimhb = morehb = 1; morem = -1;
    c->ctop = 0;

    for (k = low; k <= hgh; k++) {                            /* wave 0, align.c:1206-1340 */
        int y = (mida - k) >> 1, cc, ha, hb, na, nb;
        na = (((y + k) + (TS - aoff) - 1) / TS - 1) * TS + aoff;
        ha = peb_push(c, -1, k, 0, y + k);
        nb = ((y + (TS - boff) - 1) / TS - 1) * TS + boff;
        hb = peb_push(c, -1, k, 0, y);
        for (;;) {
            int cb = B[y], ca;
            if (cb == 4) { more = 0; if (bclip > k) bclip = k; break; }
            ca = A[y + k];
            if (cb != ca) { if (ca == 4) { more = 0; aclip = k; } break; }
            y -= 1;
        }
        cc = (y << 1) + k;
        while (y + k <= na) { ha = peb_push(c, ha, k, 0, na); na -= TS; }
        while (y <= nb)     { hb = peb_push(c, hb, k, 0, nb); nb -= TS; }
        if (cc < besta) { besta = trima = lasta = cc; besty = trimy = y; trimha = ha; trimhb = hb; }
        V[k] = cc; T[k] = PATH_INT; M[k] = PATH_LEN; HA[k] = ha; HB[k] = hb; NA[k] = na; NB[k] = nb;
    }

    for (;;) {
        if (more == 0) {                                      /* align.c:1342-1372, 1656-1684 */
            if (B[besty] != 4 && A[besta - besty] != 4) more = 1;
            if (low <= aclip) {
                low = aclip + 1;
                if (morem <= M[aclip]) { morem = M[aclip]; morea = V[aclip]; morey = (morea - aclip) / 2;
                                         mored = dif; moreha = HA[aclip]; morehb = HB[aclip]; }
            }
            if (hgh >= bclip) {
                hgh = bclip - 1;
                if (morem <= M[bclip]) { morem = M[bclip]; morea = V[bclip]; morey = (morea - bclip) / 2;
                                         mored = dif; moreha = HA[bclip]; morehb = HB[bclip]; }
            }
            aclip = -INT32_MAX; bclip = INT32_MAX;
        }
        if (dif > 0) {                                        /* prune, align.c:1686-1695 */
            int n = besta + WAVE_LAG;
            while (hgh >= low)
                if (V[hgh] > n) hgh -= 1;
                else { while (V[low] > n) low += 1; break; }
            c->waves++; c->wave_cells += (hgh - low) + 1;
        }
        if (!(more && lasta <= besta + TRIM_MLAG)) break;     /* align.c:1378 */

        {
            int am, ac, ap, n, ua, ub; uint64_t t;
            low -= 1; hgh += 1;
            if (low >= minp) { NA[low] = NA[low + 1]; NB[low] = NB[low + 1]; V[low] = ap = INT32_MAX; }
            else ap = V[++low];
            if (hgh <= maxp) { NA[hgh] = NA[hgh - 1]; NB[hgh] = NB[hgh - 1]; V[hgh] = INT32_MAX; }
            else hgh -= 1;
            dif += 1;
            ac = V[hgh + 1] = V[low - 1] = INT32_MAX;
            t = PATH_INT; n = PATH_LEN; ua = ub = -1;
            for (k = low; k <= hgh; k++) {
                int y, m, ha, hb, cc; uint64_t b;
                am = ac; ac = ap; ap = V[k + 1];
                if (ac > ap) {
                    if (ap > am) { cc = am - 1; m = n; b = t; ha = ua; hb = ub; }
                    else         { cc = ap - 1; m = M[k + 1]; b = T[k + 1]; ha = HA[k + 1]; hb = HB[k + 1]; }
                } else if (ac > am) { cc = am - 1; m = n; b = t; ha = ua; hb = ub; }
                else                { cc = ac - 2; m = M[k]; b = T[k]; ha = HA[k]; hb = HB[k]; }

                if ((b & PATH_TOP) != 0) m -= 1;
                b <<= 1;
                y = (cc - k) >> 1;
                for (;;) {
                    int cb = B[y], ca;
                    if (cb == 4) { more = 0; if (bclip > k) bclip = k; break; }
                    ca = A[y + k];
                    if (cb != ca) { if (ca == 4) { more = 0; aclip = k; } break; }
                    y -= 1;
                    if ((b & PATH_TOP) == 0) m += 1;
                    b = (b << 1) | 1;
                }
                cc = (y << 1) + k;

                while (y + k <= NA[k]) {
                    if (c->cells[ha].mark > NA[k]) ha = peb_push(c, ha, k, dif, NA[k]);
                    NA[k] -= TS;
                }
                while (y <= NB[k]) {
                    if (c->cells[hb].mark > NB[k]) hb = peb_push(c, hb, k, dif, NB[k]);
                    NB[k] -= TS;
                }
                if (cc < besta) {
                    besta = cc; besty = y;
                    if (m >= PATH_AVE) {
                        lasta = cc;
                        if (c->table[b & TRIM_MASK] >= 0)
                            if (c->table[(b >> TRIM_LEN) & TRIM_MASK] + c->score[b & TRIM_MASK] >= 0) {
                                trima = cc; trimy = y; trimd = dif; trimha = ha; trimhb = hb;
                            }
                    }
                }
                t = T[k]; n = M[k]; ua = HA[k]; ub = HB[k];
                V[k] = cc; T[k] = b; M[k] = m; HA[k] = ha; HB[k] = hb;
            }
        }
    }

    {                                                         /* finish, align.c:1711-1895 */
        uint16_t *atrace = apath->trace, *btrace = bpath->trace;
        int atlen = 0, btlen = 0, trimx, a, b, h, d, e;
        cells = c->cells;
        if (morem >= 0 && c->reach) { trimx = morea - morey; trimy = morey; trimd = mored; trimha = moreha; trimhb = morehb; }
        else trimx = trima - trimy;

        h = peb_reverse(cells, trimha);
        k = cells[h].diag; b = cells[h].mark - k; e = 0;
        if ((b + k) % TS != aoff) {                           /* seed not on an A trace boundary */
            h = cells[h].prev;
            if (h < 0) { a = trimy; d = trimd; }
            else { k = cells[h].diag; a = cells[h].mark - k; d = cells[h].diff; }
            if (apath->tlen == 0) { atrace[--atlen] = (uint16_t)(b - a); atrace[--atlen] = (uint16_t)(d - e); }
            else { atrace[1] = (uint16_t)(atrace[1] + (b - a)); atrace[0] = (uint16_t)(atrace[0] + (d - e)); }
            b = a; e = d;
        }
        if (h >= 0) {
            for (h = cells[h].prev; h >= 0; h = cells[h].prev) {
                k = cells[h].diag; a = cells[h].mark - k;
                atrace[--atlen] = (uint16_t)(b - a);
                d = cells[h].diff;
                atrace[--atlen] = (uint16_t)(d - e);
                b = a; e = d;
            }
            if (b + k != trimx) { atrace[--atlen] = (uint16_t)(b - trimy); atrace[--atlen] = (uint16_t)(trimd - e); }
            else if (b != trimy) {
                atrace[atlen + 1] = (uint16_t)(atrace[atlen + 1] + (b - trimy));
                atrace[atlen]     = (uint16_t)(atrace[atlen] + (trimd - e));
            }
        }

        h = peb_reverse(cells, trimhb);
        k = cells[h].diag; b = cells[h].mark + k; e = 0;
        if ((b - k) % TS != boff) {
            h = cells[h].prev;
            if (h < 0) { a = trimx; d = trimd; }
            else { k = cells[h].diag; a = cells[h].mark + k; d = cells[h].diff; }
            if (bpath->tlen == 0) {                           /* sic: (b-a) twice, align.c:1843-1844 */
                btrace[--btlen] = (uint16_t)(b - a); btrace[--btlen] = (uint16_t)(b - a);
            } else { btrace[1] = (uint16_t)(btrace[1] + (b - a)); btrace[0] = (uint16_t)(btrace[0] + (d - e)); }
            b = a; e = d;
        }
        if (h >= 0) {
            for (h = cells[h].prev; h >= 0; h = cells[h].prev) {
                k = cells[h].diag; a = cells[h].mark + k;
                btrace[--btlen] = (uint16_t)(b - a);
                d = cells[h].diff;
                btrace[--btlen] = (uint16_t)(d - e);
                b = a; e = d;
            }
            if (b - k != trimy) { btrace[--btlen] = (uint16_t)(b - trimx); btrace[--btlen] = (uint16_t)(trimd - e); }
            else if (b != trimx) {
                btrace[btlen + 1] = (uint16_t)(btrace[btlen + 1] + (b - trimx));
                btrace[btlen]     = (uint16_t)(btrace[btlen] + (trimd - e));
            }
        }
        apath->abpos = trimx; apath->bbpos = trimy;
        apath->diffs = apath->diffs + trimd;
        apath->tlen = apath->tlen - atlen; apath->trace = atrace + atlen;
        bpath->tlen = bpath->tlen - btlen; bpath->trace = btrace + btlen;
    }
}

/* Local_Alignment, align.c:1904-2097, for the one call shape daligner uses
 * (low == hgh == seed diagonal, lbord = hbord = -1, flags in {0, COMP_FLAG}).             */
static void local_alignment(la_ctx *c, const int8_t *aseq, int alen, const int8_t *bseq, int blen,
                            int comp, int selfie, int diag, int anti, opath *apath, opath *bpath)
{
    int need = alen + blen + 16, voff = blen + 8;
    int maxtp, minp, maxp, aoff, boff, low;

    if (need > c->vcap) {
        c->vcap = need + need / 4 + 1024;
        c->V = (int32_t *)xrealloc(c->V, sizeof(int32_t) * c->vcap);
        c->M = (int32_t *)xrealloc(c->M, sizeof(int32_t) * c->vcap);
        c->HA = (int32_t *)xrealloc(c->HA, sizeof(int32_t) * c->vcap);
        c->HB = (int32_t *)xrealloc(c->HB, sizeof(int32_t) * c->vcap);
        c->NA = (int32_t *)xrealloc(c->NA, sizeof(int32_t) * c->vcap);
        c->NB = (int32_t *)xrealloc(c->NB, sizeof(int32_t) * c->vcap);
        c->T = (uint64_t *)xrealloc(c->T, sizeof(uint64_t) * c->vcap);
    }
    maxtp = 2 * ((alen < blen ? blen : alen) / c->tspace + 2);
    if (4 * maxtp > c->tcap) { c->tcap = 4 * maxtp + 1024; c->tbuf = (uint16_t *)xrealloc(c->tbuf, sizeof(uint16_t) * c->tcap); }
    apath->trace = c->tbuf + maxtp;                           /* room to prepend, align.c:1941-1942 */
    bpath->trace = apath->trace + 2 * maxtp;

    minp = (selfie && diag >= 0) ? 1 : -INT32_MAX;            /* align.c:1951-1968 */
    maxp = (selfie && diag <= 0) ? -1 : INT32_MAX;
    aoff = 0; boff = comp ? blen % c->tspace : 0;             /* align.c:1970-1984 */

    low = diag;
    forward_wave(c, aseq, bseq, &low, diag, anti, minp, maxp, aoff, boff, apath, bpath, voff);
    reverse_wave(c, aseq, bseq, low, low, anti, minp, maxp, aoff, boff, apath, bpath, voff);

    bpath->diffs = apath->diffs;
    if (comp) {                                               /* align.c:2033-2055 */
        uint16_t *tr = bpath->trace, t; int i, j;
        bpath->abpos = blen - apath->bepos; bpath->bbpos = alen - apath->aepos;
        bpath->aepos = blen - apath->bbpos; bpath->bepos = alen - apath->abpos;
        for (i = bpath->tlen - 2, j = 0; j < i; i -= 2, j += 2) {
            t = tr[i]; tr[i] = tr[j]; tr[j] = t;
            t = tr[i + 1]; tr[i + 1] = tr[j + 1]; tr[j + 1] = t;
        }
    } else {                                                  /* align.c:2056-2062 */
        bpath->aepos = apath->bepos; bpath->bepos = apath->aepos;
        bpath->abpos = apath->bbpos; bpath->bbpos = apath->abpos;
    }
}

void ora_local_alignment(const int8_t *aseq, int alen, const int8_t *bseq, int blen,
                         int comp, int selfie, int diag, int anti, const ora_params *p,
                         int32_t *aout, uint16_t *atrace, int32_t *bout, uint16_t *btrace)
{
    la_ctx *c = new_la_ctx(p);
    opath ap, bp;
    memset(&ap, 0, sizeof(ap)); memset(&bp, 0, sizeof(bp));
    local_alignment(c, aseq, alen, bseq, blen, comp, selfie, diag, anti, &ap, &bp);
    aout[0] = ap.abpos; aout[1] = ap.bbpos; aout[2] = ap.aepos; aout[3] = ap.bepos; aout[4] = ap.diffs; aout[5] = ap.tlen;
    bout[0] = bp.abpos; bout[1] = bp.bbpos; bout[2] = bp.aepos; bout[3] = bp.bepos; bout[4] = bp.diffs; bout[5] = bp.tlen;
    memcpy(atrace, ap.trace, sizeof(uint16_t) * ap.tlen);
    memcpy(btrace, bp.trace, sizeof(uint16_t) * bp.tlen);
    free_la_ctx(c);
}

/* ------------------------------------------------------------------------------------------
 * Stage 3.  report_thread (filter.c:2128-2511) driving stage 4, incl. the thread-slice rule
 * (filter.c:2804-2816, 2210-2214) and Diagonal_Span (filter.c:2079-2110).
 * ------------------------------------------------------------------------------------------ */
ora_result *ora_new_result(void)
{
    ora_result *r = (ora_result *)xmalloc(sizeof(ora_result));
    memset(r, 0, sizeof(*r));
    return r;
}
void ora_reset_result(ora_result *r)
{
    r->novl = 0; r->ttop = 0; r->nhits = r->nfilt = r->ncheck = r->multi_la_pairs = 0; r->wave_cells = r->waves = 0;
}
void ora_free_result(ora_result *r) { if (r) { free(r->ovl); free(r->trace); free(r); } }

static void add_record(ora_result *r, const opath *pth, uint32_t flags, int aread, int bread)
{
    ora_ovl *o;
    if (r->novl >= r->omax) { r->omax = r->omax * 3 / 2 + 4096; r->ovl = (ora_ovl *)xrealloc(r->ovl, sizeof(ora_ovl) * r->omax); }
    if (r->ttop + pth->tlen > r->tmax) { r->tmax = (r->ttop + pth->tlen) * 3 / 2 + 65536; r->trace = (uint16_t *)xrealloc(r->trace, sizeof(uint16_t) * r->tmax); }
    o = r->ovl + r->novl++;
    o->tlen = pth->tlen; o->diffs = pth->diffs; o->abpos = pth->abpos; o->bbpos = pth->bbpos;
    o->aepos = pth->aepos; o->bepos = pth->bepos; o->flags = flags; o->aread = aread; o->bread = bread;
    o->toff = r->ttop;
    memcpy(r->trace + r->ttop, pth->trace, sizeof(uint16_t) * pth->tlen);
    r->ttop += pth->tlen;
}

static inline uint64_t pair_key(const ora_hit *h) { return ((uint64_t)(uint32_t)h->bread << 32) | (uint32_t)h->aread; }

void ora_report(const ora_block *ab, const ora_block *bb, const ora_hit *hits, int64_t nhits,
                int self, int comp, const ora_params *p, ora_result *res)
{
    const int K = p->kmer, W = p->binshift, H = p->hitmin, TS = p->tspace;
    const int maxdiag = (ab->maxlen >> W), mindiag = ((-bb->maxlen) >> W);
    const int span = maxdiag - mindiag + 1;
    int32_t *score0 = (int32_t *)xmalloc(sizeof(int32_t) * 3 * (span + 8));
    int32_t *score = score0 + 4 - mindiag, *lastp = score + span + 8, *lasta = lastp + span + 8;
    int nthreads = 1, nshift = 0, t;
    const int minhit = (H - 1) / K + 1;
    la_ctx *c = new_la_ctx(p);
    /* A and B records of one pair are buffered so that A records precede B records (filter.c:2452-2483) */
    opath *amatch = NULL, *bmatch = NULL; int amax = 0; uint16_t *ptr = NULL; int64_t ptop, pmax = 0;

    memset(score0, 0, sizeof(int32_t) * 3 * (span + 8));
    while (2 * nthreads <= p->nthreads) { nthreads *= 2; nshift++; }      /* filter.c:192-198 */

    for (t = 0; t < nthreads; t++) {
        int64_t beg, end, nidx, eidx;
        {                                                     /* slice bounds, filter.c:2804-2816 */
            int64_t q = (nhits * t) >> nshift;
            if (t > 0 && q > 0) { int d = hits[q - 1].bread; while (hits[q].bread == d) q++; }
            beg = (t == 0) ? 0 : q;
            if (t == nthreads - 1) end = nhits;
            else { q = (nhits * (t + 1)) >> nshift; if (q > 0) { int d = hits[q - 1].bread; while (hits[q].bread == d) q++; } end = q; }
        }
        eidx = end - minhit; nidx = beg;
        while (nidx < eidx) {
            uint64_t cpair = pair_key(hits + nidx);
            if (pair_key(hits + nidx + (minhit - 1)) != cpair) {          /* fewer than minhit hits */
                nidx++; while (pair_key(hits + nidx) == cpair) nidx++;
                continue;
            }
            {
                int ar = hits[nidx].aread, br = hits[nidx].bread;
                int alen = ab->rlen[ar], blen = bb->rlen[br];
                const int8_t *aseq = ab->bases + ab->boff[ar], *bseq = bb->bases + bb->boff[br];
                int doB = (p->symmetric && (ar != br || !self || !comp));
                int amark, amark2 = 0, novl = 0;
                int64_t sidx = nidx, lidx, h2, f;
                ptop = 0;
                while (pair_key(hits + nidx) == cpair) {                  /* A-panels, filter.c:2251-2266 */
                    int apos;
                    amark = amark2 + PANEL_SIZE; amark2 = amark - PANEL_OVERLAP;
                    h2 = lidx = nidx;
                    do {
                        apos = hits[nidx].apos; nidx++;
                        if (apos <= amark2) h2 = nidx;
                    } while (pair_key(hits + nidx) == cpair && apos <= amark);
                    if (nidx - lidx >= minhit) {
                        for (f = lidx; f < nidx; f++) {                   /* pass 1, filter.c:2268-2277 */
                            int d = hits[f].diag >> W; apos = hits[f].apos;
                            score[d] += (apos - lastp[d] >= K) ? K : apos - lastp[d];
                            lastp[d] = apos;
                        }
                        for (f = lidx; f < nidx; f++) {                   /* pass 2, filter.c:2283-2405 */
                            int diag = hits[f].diag, bpos, d;
                            apos = hits[f].apos; bpos = apos - diag; d = diag >> W;
                            if (apos > lasta[d] && (score[d] + score[d + 1] >= H || score[d] + score[d - 1] >= H)) {
                                opath ap, bp; int lo, hi, dd, i;
                                memset(&ap, 0, sizeof(ap)); memset(&bp, 0, sizeof(bp));
                                res->nfilt++;
                                local_alignment(c, aseq, alen, bseq, blen, comp, aseq == bseq, diag, apos + bpos, &ap, &bp);
                                /* Diagonal_Span, filter.c:2079-2110 */
                                dd = ap.abpos - ap.bbpos; lo = hi = dd;
                                dd = ap.aepos - ap.bepos; if (dd < lo) lo = dd; else if (dd > hi) hi = dd;
                                dd = (ap.abpos / TS) * TS - ap.bbpos;
                                for (i = 1; i < ap.tlen - 2; i += 2) { dd += TS - ap.trace[i]; if (dd < lo) lo = dd; else if (dd > hi) hi = dd; }
                                lo = (lo >> W) - 1; hi = (hi >> W) + 1;
                                if (d < lo) lo = d; else if (d > hi) hi = d;
                                for (i = lo; i <= hi; i++) if (ap.aepos > lasta[i]) lasta[i] = ap.aepos;

                                if ((ap.aepos - ap.abpos) + (ap.bepos - ap.bbpos) >= p->minover) {
                                    int need = ap.tlen + bp.tlen;
                                    if (novl >= amax) { amax = amax * 2 + 64; amatch = (opath *)xrealloc(amatch, sizeof(opath) * amax); bmatch = (opath *)xrealloc(bmatch, sizeof(opath) * amax); }
                                    if (ptop + need > pmax) {
                                        pmax = (ptop + need) * 2 + 65536; ptr = (uint16_t *)xrealloc(ptr, sizeof(uint16_t) * pmax);
                                    }
                                    amatch[novl] = ap; amatch[novl].trace = (uint16_t *)(intptr_t)ptop;
                                    memcpy(ptr + ptop, ap.trace, sizeof(uint16_t) * ap.tlen); ptop += ap.tlen;
                                    bmatch[novl] = bp; bmatch[novl].trace = (uint16_t *)(intptr_t)ptop;
                                    memcpy(ptr + ptop, bp.trace, sizeof(uint16_t) * bp.tlen); ptop += bp.tlen;
                                    novl++;
                                }
                            }
                        }
                        for (f = lidx; f < nidx; f++) { int d = hits[f].diag >> W; score[d] = lastp[d] = 0; }
                    }
                    nidx = h2;
                }
                for (f = sidx; f < nidx; f++) {                           /* filter.c:2417-2432 */
                    int d0 = hits[f].diag >> W, d;
                    for (d = d0; d <= maxdiag; d++) { if (lasta[d] == 0) break; lasta[d] = 0; }
                    for (d = d0 - 1; d >= mindiag; d--) { if (lasta[d] == 0) break; lasta[d] = 0; }
                }
                lasta[mindiag - 1] = lasta[mindiag - 2] = lasta[maxdiag + 1] = lasta[maxdiag + 2] = 0;
                if (novl > 1) res->multi_la_pairs++;
                {
                    int i;
                    for (i = 0; i < novl; i++) {
                        opath q = amatch[i]; q.trace = ptr + (intptr_t)amatch[i].trace;
                        add_record(res, &q, (uint32_t)comp, ar + ab->ufirst, br + bb->ufirst);
                    }
                    if (doB)
                        for (i = 0; i < novl; i++) {
                            opath q = bmatch[i]; q.trace = ptr + (intptr_t)bmatch[i].trace;
                            add_record(res, &q, (uint32_t)comp, br + bb->ufirst, ar + ab->ufirst);
                        }
                    res->ncheck += novl + (doB ? novl : 0);
                }
            }
        }
    }
    res->wave_cells += c->wave_cells; res->waves += c->waves;
    free(amatch); free(bmatch); free(ptr);
    free_la_ctx(c);
    free(score0);
}

/* ------------------------------------------------------------------------------------------
 * datander: Match_Self (scrub/tandem.c:1182-1428).  The reference sorts all k-mers of the block by code (tandem.c:428),
 * turns every entry into the distance to the previous occurrence of its k-mer IN THE SAME READ (0: none; count_thread,
 * tandem.c:556-590), sorts back to (read, rpos) order (:1281) and lets report_thread walk every read's positions.  Per read
 * that is: link[apos] = apos - (previous apos with the same k-mer in this read), which is computed here read by read.
 * One deviation, undefined in the reference: count_thread never rewrites the FIRST entry of a thread's range (its zeroing
 * store after the inner loop hits the next run's head, :588), so report_thread reads a raw k-mer code as a distance and
 * indexes its bucket arrays out of range with it; that entry is a run head, i.e. "no link" (0), which is what it gets here.
 * rpos = index of the k-mer's last base + 1 (tandem.c:412), so apos runs over [K, alen].
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t code; int32_t apos; } self_kmer;
static int cmp_self_kmer(const void *x, const void *y)
{
    const self_kmer *a = (const self_kmer *)x, *b = (const self_kmer *)y;
    if (a->code != b->code) return a->code < b->code ? -1 : 1;
    return a->apos - b->apos;
}

void ora_match_self(const ora_block *blk, const ora_params *p, ora_result *res)
{
    const int K = p->kmer, W = p->binshift, H = p->hitmin, TS = p->tspace;
    const int maxdiag = (blk->maxlen >> W), mindiag = ((-blk->maxlen) >> W);
    const int span = maxdiag - mindiag + 1;
    const uint64_t kmask = (K >= 32) ? ~0ull : ((1ull << (2 * K)) - 1);
    int32_t *score0 = (int32_t *)xmalloc(sizeof(int32_t) * 3 * (span + 8));
    int32_t *score = score0 + 4 - mindiag, *lastp = score + span + 8, *lasta = lastp + span + 8;
    int32_t *link = (int32_t *)xmalloc(sizeof(int32_t) * ((size_t)blk->maxlen + 2));
    self_kmer *km = (self_kmer *)xmalloc(sizeof(self_kmer) * ((size_t)blk->maxlen + 2));
    la_ctx *c = new_la_ctx(p);
    int ar;

    c->reach = 0;                                 /* New_Align_Spec(..., reach = 0), scrub/datander.c:246 */
    memset(score0, 0, sizeof(int32_t) * 3 * (span + 8));
    for (ar = 0; ar < blk->nreads; ar++) {
        const int alen = blk->rlen[ar];
        const int8_t *aseq = blk->bases + blk->boff[ar];
        int apos, n = 0, amarkb, amarke, novl0 = (int)res->novl;
        uint64_t code = 0;
        (void)novl0;
        for (apos = 1; apos <= alen; apos++) {                       /* tuple_thread, tandem.c:406-418 */
            code = ((code << 2) | (uint64_t)aseq[apos - 1]) & kmask;
            link[apos] = 0;
            if (apos >= K) { km[n].code = code; km[n].apos = apos; n++; }
        }
        qsort(km, (size_t)n, sizeof(self_kmer), cmp_self_kmer);
        for (apos = 1; apos < n; apos++)                             /* count_thread, tandem.c:566-586 */
            if (km[apos].code == km[apos - 1].code) link[km[apos].apos] = km[apos].apos - km[apos - 1].apos;

        amarkb = K; amarke = PANEL_SIZE;                             /* report_thread, tandem.c:966-971 */
        if (amarke >= alen) amarke = alen + 1;
        for (;;) {
            for (apos = amarkb; apos < amarke; apos++) {             /* :976-985 */
                int d = link[apos];
                if (d == 0) continue;
                d >>= W;
                score[d] += (apos - lastp[d] >= K) ? K : apos - lastp[d];
                lastp[d] = apos;
            }
            for (apos = amarkb; apos < amarke; apos++) {             /* :989-1100 */
                int diag = link[apos], d, bpos;
                if (diag == 0) continue;
                d = diag >> W;
                if (apos > lasta[d] && (score[d] + score[d + 1] >= H || score[d] + score[d - 1] >= H)) {
                    opath ap, bp; int lo, hi, dd, i;
                    bpos = apos - diag;
                    memset(&ap, 0, sizeof(ap)); memset(&bp, 0, sizeof(bp));
                    res->nfilt++;
                    local_alignment(c, aseq, alen, aseq, alen, 0, 1, diag, apos + bpos, &ap, &bp);
                    dd = ap.abpos - ap.bbpos; lo = hi = dd;          /* Diagonal_Span, tandem.c:852-890 */
                    dd = ap.aepos - ap.bepos; if (dd < lo) lo = dd; else if (dd > hi) hi = dd;
                    dd = (ap.abpos / TS) * TS - ap.bbpos;
                    for (i = 1; i < ap.tlen - 2; i += 2) { dd += TS - ap.trace[i]; if (dd < lo) lo = dd; else if (dd > hi) hi = dd; }
                    lo = (lo >> W) - 1; hi = (hi >> W) + 1;
                    if (d < lo) lo = d; else if (d > hi) hi = d;
                    for (i = lo; i <= hi; i++) if (ap.aepos > lasta[i]) lasta[i] = ap.aepos;
                    if ((ap.aepos - ap.abpos) + (ap.bepos - ap.bbpos) >= p->minover) {
                        add_record(res, &ap, 0u, ar + blk->ufirst, ar + blk->ufirst);
                        res->ncheck++;
                    }
                }
            }
            for (apos = amarkb; apos < amarke; apos++) {             /* :1103-1108 */
                int d = link[apos];
                if (d == 0) continue;
                d >>= W; score[d] = lastp[d] = 0;
            }
            if (amarke > alen) break;
            amarkb = amarke - PANEL_OVERLAP; amarke = amarkb + PANEL_SIZE;
            if (amarke > alen) amarke = alen + 1;
        }
        memset(lasta + mindiag - 2, 0, sizeof(int32_t) * (span + 4));   /* :1119-1136 */
    }
    res->wave_cells += c->wave_cells; res->waves += c->waves;
    free_la_ctx(c);
    free(km); free(link); free(score0);
}

void ora_match_filter(const ora_block *ab, const ora_block *bb,
                      const ora_kmer *al, int64_t alen, const ora_kmer *bl, int64_t blen,
                      int self, int comp, const ora_params *p, ora_result *res)
{
    int64_t nhits; int32_t limit;
    ora_hit *hits = ora_seed_hits(ab, bb, al, alen, bl, blen, self, comp, p, &nhits, &limit);
    res->nhits += nhits; res->limit = limit;
    if (hits == NULL) return;
    ora_report(ab, bb, hits, nhits, self, comp, p, res);
    free(hits);
}

/* ------------------------------------------------------------------------------------------
 * complement_DB (daligner.c:511-628): reverse-complement reads, mirror mask intervals.
 * ------------------------------------------------------------------------------------------ */
ora_block *ora_complement_block(const ora_block *blk)
{
    ora_block *c = (ora_block *)xmalloc(sizeof(ora_block));
    int64_t tot = blk->boff[blk->nreads];
    int8_t *seq = (int8_t *)xmalloc(tot + 1);
    int i;
    *c = *blk;
    seq[0] = 4; memcpy(seq + 1, blk->bases, tot);
    for (i = 0; i < blk->nreads; i++) {
        int8_t *s = seq + 1 + blk->boff[i], *e = s + blk->rlen[i] - 1;
        while (s < e) { int8_t x = *s; *s++ = (int8_t)(3 - *e); *e-- = (int8_t)(3 - x); }
        if (s == e) *s = (int8_t)(3 - *s);
    }
    c->bases = seq + 1;
    if (blk->manno) {
        int64_t n = blk->manno[blk->nreads];
        int32_t *d = (int32_t *)xmalloc(sizeof(int32_t) * (n + 1));
        for (i = 0; i < blk->nreads; i++) {
            int64_t lo = blk->manno[i], hi = blk->manno[i + 1] - 1; int rl = blk->rlen[i];
            while (lo < hi) { int x = blk->mdata[hi]; d[hi--] = rl - blk->mdata[lo]; d[lo++] = rl - x; }
            if (lo == hi) d[lo] = rl - blk->mdata[lo];
        }
        c->mdata = d;
    }
    return c;
}
void ora_free_block(ora_block *c)
{
    if (!c) return;
    free((void *)(c->bases - 1));
    if (c->manno) free((void *)c->mdata);
    free(c);
}

/* ------------------------------------------------------------------------------------------
 * Write_Overlap_Buffer (align.c:6104-6367) + Write_Overlap (align.c:3365) + Compress_TraceTo8
 * (align.c:3375).  qsort falls back to element address on full key ties (align.c:6158-6163);
 * with glibc's merge sort that is input order, which is what the index tiebreak gives.
 * ------------------------------------------------------------------------------------------ */
static const ora_ovl *g_sort_base;
static int cmp_ovl(const void *x, const void *y)
{
    const ora_ovl *l = g_sort_base + *(const int64_t *)x, *r = g_sort_base + *(const int64_t *)y;
    int cl, cr;
    if (l->aread != r->aread) return l->aread - r->aread;
    if (l->bread != r->bread) return l->bread - r->bread;
    cl = l->flags & 1; cr = r->flags & 1;
    if (cl != cr) return cl - cr;
    if (l->abpos != r->abpos) return l->abpos - r->abpos;
    if (l->aepos != r->aepos) return l->aepos - r->aepos;
    if (l->bbpos != r->bbpos) return l->bbpos - r->bbpos;
    if (l->bepos != r->bepos) return l->bepos - r->bepos;
    return (l < r) ? -1 : (l > r);
}

static void put_records(const ora_result *r, const ora_params *p, const int64_t *idx, int64_t from, int64_t to,
                        uint8_t **out, int64_t *len)
{
    int tbytes = (p->tspace <= TRACE_XOVR) ? 1 : 2;
    int64_t i, sz = 12, o;
    uint8_t *buf;
    for (i = from; i < to; i++) sz += 40 + (p->no_trace ? 0 : (int64_t)tbytes * r->ovl[idx[i]].tlen);
    buf = (uint8_t *)xmalloc(sz);
    { int64_t n = to - from; int32_t ts = p->tspace; memcpy(buf, &n, 8); memcpy(buf + 8, &ts, 4); }
    o = 12;
    for (i = from; i < to; i++) {
        const ora_ovl *v = r->ovl + idx[i];
        int32_t rec[10]; int j;
        rec[0] = p->no_trace ? 0 : v->tlen; rec[1] = v->diffs; rec[2] = v->abpos; rec[3] = v->bbpos;
        rec[4] = v->aepos; rec[5] = v->bepos; rec[6] = (int32_t)v->flags; rec[7] = v->aread; rec[8] = v->bread; rec[9] = 0;
        memcpy(buf + o, rec, 40); o += 40;
        if (p->no_trace) continue;
        for (j = 0; j < v->tlen; j++) {
            uint16_t x = r->trace[v->toff + j];
            if (tbytes == 1) {
                if (x > 255) { fprintf(stderr, "oracle: Compression of trace to bytes fails, value too big\n"); exit(1); }
                buf[o++] = (uint8_t)x;
            } else { memcpy(buf + o, &x, 2); o += 2; }
        }
    }
    *out = buf; *len = o;
}

void ora_serialise_las(ora_result *r, const ora_params *p, int only_identity, int split_read,
                       uint8_t **out1, int64_t *len1, uint8_t **out2, int64_t *len2)
{
    int64_t *idx = (int64_t *)xmalloc(sizeof(int64_t) * (r->novl + 1));
    int64_t n = 0, i, cut;
    for (i = 0; i < r->novl; i++)
        if (!only_identity || r->ovl[i].aread == r->ovl[i].bread) idx[n++] = i;
    g_sort_base = r->ovl;
    qsort(idx, n, sizeof(int64_t), cmp_ovl);
    if (split_read < 0) {
        put_records(r, p, idx, 0, n, out1, len1);
        if (out2) { *out2 = NULL; *len2 = 0; }
    } else {
        for (cut = 0; cut < n; cut++) if (r->ovl[idx[cut]].aread > split_read) break;
        put_records(r, p, idx, 0, cut, out1, len1);
        put_records(r, p, idx, cut, n, out2, len2);
    }
    free(idx);
}

/* ------------------------------------------------------------------------------------------
 * Compute_Trace_PTS, cost form (align.c:5577-5692).  Segment i of a record is A[ab, ae) x B[bb, bb + b_i) with
 * ae = the next multiple of the trace spacing (the last segment ends at aepos, bepos; align.c:5655-5688); iter_np
 * (align.c:4892-5230) returns D + |M - N| where D counts mismatches 1, steps away from the end diagonal 2 and steps
 * toward it 0, i.e. the unit-cost edit distance of the two strings.  Restated as the plain dynamic program.
 * ------------------------------------------------------------------------------------------ */
static int edit_distance(const int8_t *a, int m, const int8_t *b, int n, int *row)
{
    int i, j;
    for (j = 0; j <= n; j++) row[j] = j;
    for (i = 1; i <= m; i++) {
        int diag = row[0], v;
        row[0] = i;
        for (j = 1; j <= n; j++) {
            v = diag + (a[i - 1] != b[j - 1]);
            diag = row[j];
            if (row[j] + 1 < v) v = row[j] + 1;
            if (row[j - 1] + 1 < v) v = row[j - 1] + 1;
            row[j] = v;
        }
    }
    return row[n];
}

void ora_trace_point_diffs(const ora_block *ablk, const ora_block *bblk, const ora_block *cblk, const int32_t *recs,
                           const uint64_t *toff, const uint16_t *trace, int64_t novl, int tspace,
                           int32_t *rec_diffs, uint16_t *seg)
{
    int64_t r;
    int *row = (int *)malloc(sizeof(int) * ((size_t)bblk->maxlen + 2));
    for (r = 0; r < novl; r++) {
        const int32_t *o = recs + 10 * r;
        const int tlen = o[0], abpos = o[2], bbpos = o[3], aepos = o[4], bepos = o[5], comp = o[6] & 1;
        const int ar = o[7] - ablk->ufirst, br = o[8] - bblk->ufirst;
        const ora_block *bsrc = comp ? cblk : bblk;
        const int8_t *aseq = ablk->bases + ablk->boff[ar], *bseq = bsrc->bases + bsrc->boff[br];
        const uint16_t *t = trace + toff[r];
        const int nseg = tlen / 2;
        int i, ab = abpos, bb = bbpos, ae = (abpos / tspace) * tspace, be, total = 0;
        for (i = 0; i < nseg; i++) {
            if (i == nseg - 1) { ae = aepos; be = bepos; }
            else { ae += tspace; be = bb + t[2 * i + 1]; }
            {
                const int d = edit_distance(aseq + ab, ae - ab, bseq + bb, be - bb, row);
                if (seg) seg[toff[r] / 2 + i] = (uint16_t)d;
                total += d;
            }
            ab = ae; bb = be;
        }
        rec_diffs[r] = total;
    }
    free(row);
}
