/*
 * tandem_compat.c -- drop-in replacement of scrub/tandem.c for the UNMODIFIED reference host scrub/datander.c:
 * Set_Filter_Params / Match_Self (scrub/tandem.h:57-60) on top of the C-ABI of include/damar_b200.h.  Plain C; compiled
 * into the reference's datander link line in place of tandem.c.  All arithmetic happens in
 * libdamar_b200.so on the GPU.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/damar_b200.h"
#include "../../include/damar_tandem_compat.h"

int Set_Filter_Params(int kmer, int binshift, int hitmin, int nthreads)
{
    return damar_set_filter_params(kmer, binshift, 0, hitmin, nthreads);     /* tandem.c:137-168: no -t */
}

/* tandem.c:1182 */
void Match_Self(char *aname, DAMAR_HITS_DB *block, DAMAR_Align_Spec *spec)
{
    damar_block_desc d;
    damar_options o;
    damar_overlaps ov;
    damar_block *blk;
    DAMAR_Overlap_IO_Buffer *buf = OVL_IO_Buffer(spec);
    int64_t *boff, *order, i;
    int32_t *rlen;

    memset(&d, 0, sizeof(d));
    boff = (int64_t *)malloc(sizeof(int64_t) * (block->nreads + 1));
    rlen = (int32_t *)malloc(sizeof(int32_t) * (block->nreads + 1));
    if (boff == NULL || rlen == NULL) { fprintf(stderr, "Match_Self: out of memory\n"); exit(1); }
    for (i = 0; i < block->nreads; i++) { boff[i] = block->reads[i].boff; rlen[i] = block->reads[i].rlen; }
    boff[block->nreads] = block->reads[block->nreads].boff;
    d.nreads = block->nreads; d.maxlen = block->maxlen; d.totlen = block->totlen;
    d.ufirst = block->ufirst; d.part = block->part;
    d.bases = (const uint8_t *)block->bases; d.boff = boff; d.rlen = rlen;
    memcpy(d.freq, block->freq, sizeof(d.freq));
    d.sizeof_db = 0;
    blk = damar_block_upload(&d);
    free(boff); free(rlen);
    if (blk == NULL) { fprintf(stderr, "%s\n", damar_last_error()); exit(1); }

    memset(&o, 0, sizeof(o));
    o.verbose = VERBOSE; o.minover = MINOVER; o.symmetric = 1;
    o.trace_spacing = Trace_Spacing(spec); o.ave_corr = Average_Correlation(spec);
    memcpy(o.freq, Base_Frequencies(spec), sizeof(o.freq));
    damar_set_options(&o);
    if (VERBOSE) printf("\nIndexing %s\n\nComparing %s to itself\n", aname, aname);

    memset(&ov, 0, sizeof(ov));
    if (damar_match_self(blk, &ov) != 0) { fprintf(stderr, "%s\n", damar_last_error()); exit(1); }
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
    if (VERBOSE) {                                    /* tandem.c:1411-1425 */
        double mat = (double)block->totlen * (double)block->totlen;
        if (mat <= 0.) mat = 1.;
        printf("\n     %9lld seed hits (%e of matrix)\n     %9lld confirmed hits (%e of matrix)\n",
               (long long)ov.nfilt, (double)ov.nfilt / mat, (long long)ov.ncheck, (double)ov.ncheck / mat);
        fflush(stdout);
    }
    damar_block_free(blk);
}
