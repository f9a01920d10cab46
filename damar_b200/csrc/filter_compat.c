/*
 * filter_compat.c -- drop-in replacement of dalign/filter.c for the UNMODIFIED reference host
 * (dalign/daligner.c): Set_Filter_Params / Sort_Kmers / Match_Filter on top of the C-ABI of
 * include/damar_b200.h.  Plain C; compiled into the reference's daligner link line in place of
 * filter.c.  All arithmetic happens in libdamar_b200.so on the GPU.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/damar_b200.h"
#include "../../include/damar_filter_compat.h"

_Static_assert(sizeof(DAMAR_HITS_READ) == 32, "HITS_READ layout (db/DB.h:319-325)");
_Static_assert(sizeof(DAMAR_HITS_DB) == 88, "HITS_DB layout (db/DB.h:361-389)");
_Static_assert(sizeof(DAMAR_Overlap) == 48, "Overlap layout (dalign/align.h:358-364)");

typedef struct {                 /* what Sort_Kmers hands back as the opaque `table` */
    damar_block *blk;
    damar_index *idx;
} shim_table;

static int g_nthreads = 4;
static int g_kmer = 14;

int Set_Filter_Params(int kmer, int binshift, int suppress, int hitmin, int nthreads)
{
    g_nthreads = nthreads;
    g_kmer = kmer;
    damar_set_biased(BIASED);                        /* daligner.c parses -b before this call (:729, :862) */
    return damar_set_filter_params(kmer, binshift, suppress, hitmin, nthreads);
}

/* filter.c:753.  The host has already applied complement_DB when it wants the complement index
 * (daligner.c:998-1003, 1043-1048), so the block is shipped as it is.                           */
void *Sort_Kmers(DAMAR_HITS_DB *block, int *len)
{
    damar_block_desc d;
    int64_t *boff, n = 0;
    int32_t *rlen;
    shim_table *t;
    int i;

    memset(&d, 0, sizeof(d));
    boff = (int64_t *)malloc(sizeof(int64_t) * (block->nreads + 1));
    rlen = (int32_t *)malloc(sizeof(int32_t) * (block->nreads + 1));
    if (boff == NULL || rlen == NULL) { fprintf(stderr, "Sort_Kmers: out of memory\n"); exit(1); }
    for (i = 0; i < block->nreads; i++) { boff[i] = block->reads[i].boff; rlen[i] = block->reads[i].rlen; }
    boff[block->nreads] = block->reads[block->nreads].boff;
    d.nreads = block->nreads; d.maxlen = block->maxlen; d.totlen = block->totlen;
    d.ufirst = block->ufirst; d.part = block->part;
    d.bases = (const uint8_t *)block->bases; d.boff = boff; d.rlen = rlen;
    if (block->tracks != NULL) {                     /* first track = merged mask (filter.c:474-480) */
        d.mask_anno = (const int64_t *)block->tracks->anno;
        d.mask_data = (const int32_t *)block->tracks->data;
    }
    memcpy(d.freq, block->freq, sizeof(d.freq));
    d.sizeof_db = sizeof_DB(block);
    d.flags = DAMAR_BLOCK_AS_IS;                     /* never complement on the device */
    t = (shim_table *)malloc(sizeof(shim_table));
    t->blk = damar_block_upload(&d);
    free(boff); free(rlen);
    if (t->blk == NULL) { fprintf(stderr, "%s\n", damar_last_error()); exit(1); }
    t->idx = damar_sort_kmers(t->blk, 0, &n);
    if (n < 0) { fprintf(stderr, "%s\n", damar_last_error()); exit(1); }      /* failed (NULL with n == 0 is an empty list) */
    *len = (int)n;
    if (t->idx == NULL) { damar_block_free(t->blk); free(t); return NULL; }   /* filter.c:990-993 */
    return t;
}

/* filter.c:2519 */
void Match_Filter(char *aname, DAMAR_HITS_DB *ablock, char *bname, DAMAR_HITS_DB *bblock,
                  void *atable, int alen, void *btable, int blen, int comp, DAMAR_Align_Spec *spec)
{
    shim_table *ta = (shim_table *)atable, *tb = (shim_table *)btable;
    damar_options o;
    damar_overlaps ov;
    DAMAR_Overlap_IO_Buffer *buf = OVL_IO_Buffer(spec);
    float *freq = Base_Frequencies(spec);
    int64_t i, *order;
    (void)alen; (void)blen;

    memset(&o, 0, sizeof(o));
    o.verbose = VERBOSE; o.minover = MINOVER; o.hgap_min = HGAP_MIN; o.symmetric = SYMMETRIC; o.identity = IDENTITY;
    o.trace_spacing = Trace_Spacing(spec); o.ave_corr = Average_Correlation(spec);
    memcpy(o.freq, freq, sizeof(o.freq));
    o.mem_limit = MEM_LIMIT;
    damar_set_options(&o);
    if (VERBOSE) { if (comp) printf("\nComparing %s to c(%s)\n", aname, bname); else printf("\nComparing %s to %s\n", aname, bname); }

    memset(&ov, 0, sizeof(ov));
    if (ta != NULL && tb != NULL)
        if (damar_match_filter(ta->blk, tb->blk, ta->idx, tb->idx, aname == bname, comp, &ov) != 0) {
            fprintf(stderr, "%s\n", damar_last_error()); exit(1);
        }

    /* hand the records to the host's Overlap_IO_Buffer in the order its report threads would have
       (emission rank), so that Write_Overlap_Buffer's address tie-break sees the same sequence   */
    order = (int64_t *)malloc(sizeof(int64_t) * (ov.novl + 1));
    damar_emission_order(&ov, order);
    for (i = 0; i < ov.novl; i++) {
        const damar_overlap *r = ov.ovl + order[i];
        DAMAR_Overlap v;
        v.path.trace = (void *)((const uint8_t *)ov.trace + ov.trace_off[order[i]] * (uint64_t)ov.trace_bytes);
        v.path.tlen = r->tlen; v.path.diffs = r->diffs;
        v.path.abpos = r->abpos; v.path.bbpos = r->bbpos; v.path.aepos = r->aepos; v.path.bepos = r->bepos;
        v.flags = r->flags; v.aread = r->aread; v.bread = r->bread;
        AddOverlapToBuffer(buf, &v, ov.trace_bytes);
    }
    free(order);
    if (VERBOSE) {
        /* filter.c:2910-2928 prints these three counts with their fraction of the alen x blen matrix */
        double mat = (double)ablock->totlen * (double)bblock->totlen;
        if (mat <= 0.) mat = 1.;
        printf("\n     %9lld %d-mers (%e of matrix)\n     %9lld seed hits (%e of matrix)\n     %9lld confirmed hits (%e of matrix)\n",
               (long long)ov.nhits, g_kmer, (double)ov.nhits / mat, (long long)ov.nfilt, (double)ov.nfilt / mat,
               (long long)ov.ncheck, (double)ov.ncheck / mat);
        fflush(stdout);
    }
    if (tb != NULL && tb != ta) {                    /* the reference consumes btable (filter.c:2880-2881) */
        damar_index_free(tb->idx); damar_block_free(tb->blk); free(tb);
    }
}
