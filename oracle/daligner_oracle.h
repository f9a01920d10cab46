/*
 * daligner_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of DAmar's daligner overlap path
 * (reference: dalign/filter.c + dalign/align.c).  It exists so that the CUDA product
 * (damar_b200/csrc) can be checked stage by stage and end to end.  Nothing in the
 * product may include, link or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may.
 *
 * Parity pin: the reference ships no golden vectors, so this restatement
 * is pinned against the reference binary itself (oracle/_ref/daligner, built by
 * oracle/Makefile from /root/reference) -- see tests/test_oracle_vs_ref.py and the
 * committed fixtures under tests/golden/.
 */
#ifndef DALIGNER_ORACLE_H
#define DALIGNER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One DB block in core: what read_DB + Read_All_Sequences leave behind
 * (daligner.c:442-509, DB.c:1547-1608), reduced to plain arrays.               */
typedef struct {
    int32_t        nreads;
    int32_t        maxlen;
    int64_t        totlen;
    int32_t        ufirst;     /* DB-global index of read 0 of this block                 */
    const int8_t  *bases;      /* 1 byte/base 0..3; bases[-1]==4 and a 4 after every read */
    const int64_t *boff;       /* [nreads+1] offset of each read in bases                 */
    const int32_t *rlen;       /* [nreads]                                                */
    const int64_t *manno;      /* mask track: [nreads+1] offsets (in ints) or NULL        */
    const int32_t *mdata;      /* mask track: sorted [beg,end) pairs per read             */
} ora_block;

/* filter.c:121-126 / 128-134 (little-endian layouts) */
typedef struct { uint64_t code; int32_t rpos; int32_t read; } ora_kmer;
typedef struct { int32_t diag; int32_t apos; int32_t aread; int32_t bread; } ora_hit;

typedef struct {
    int32_t  kmer, binshift, suppress, hitmin, nthreads;  /* Set_Filter_Params, filter.c:171 */
    int32_t  identity, symmetric;                         /* -I, !-A                          */
    int32_t  minover;                                     /* already doubled (daligner.c:861) */
    int32_t  tspace;                                      /* -s                               */
    int32_t  no_trace;                                    /* -T                               */
    double   ave_corr;                                    /* -e                               */
    float    freq[4];                                     /* A block base frequencies         */
    uint64_t mem_limit;                                   /* MEM_LIMIT in bytes; 0 == -M0     */
    int64_t  sizeof_a, sizeof_b;                          /* sizeof_DB() of the two blocks    */
    int32_t  biased;                                      /* -b (BIASED, daligner.c:729)       */
} ora_params;

/* One .las record as it reaches Write_Overlap (align.c:3365): 36 meaningful bytes. */
typedef struct {
    int32_t  tlen, diffs, abpos, bbpos, aepos, bepos;
    uint32_t flags;
    int32_t  aread, bread;
    int64_t  toff;             /* offset of this record's trace in the trace pool (values) */
} ora_ovl;

typedef struct {
    int64_t   novl, omax;
    ora_ovl  *ovl;
    int64_t   ttop, tmax;
    uint16_t *trace;           /* uncompressed values; serialised as u8 when tspace<=125   */
    /* counters (what -v prints, filter.c:2910-2928) + diagnostics                         */
    int64_t   nhits, nfilt, ncheck;
    int64_t   multi_la_pairs;  /* pairs that would enter Handle_Redundancies (filter.c:2442) */
    int64_t   wave_cells, waves;
    int32_t   limit;
} ora_result;

/* ---- stage 1: Sort_Kmers (filter.c:753) ---- */
ora_kmer *ora_sort_kmers(const ora_block *blk, const ora_params *p, int64_t *len);

/* ---- stage 2: count + merge + seed-hit sort (filter.c:1039-1358, 2606-2781) ---- */
ora_hit *ora_seed_hits(const ora_block *ab, const ora_block *bb,
                       const ora_kmer *al, int64_t alen, const ora_kmer *bl, int64_t blen,
                       int self, int comp, const ora_params *p, int64_t *nhits, int32_t *limit);

/* ---- stages 3+4: report_thread + Local_Alignment (filter.c:2128, align.c:1904) ---- */
void ora_report(const ora_block *ab, const ora_block *bb, const ora_hit *hits, int64_t nhits,
                int self, int comp, const ora_params *p, ora_result *res);

/* whole Match_Filter for one orientation; appends to res */
void ora_match_filter(const ora_block *ab, const ora_block *bb,
                      const ora_kmer *al, int64_t alen, const ora_kmer *bl, int64_t blen,
                      int self, int comp, const ora_params *p, ora_result *res);

/* datander: Match_Self (scrub/tandem.c:1182-1428) up to its Handle_Redundancies call (tandem.c:1148) -- k-mer links inside a
 * read (Sort_Kmers :428, count_thread :556, the (read, rpos) re-sort :1281), report_thread's panel walk over the read's own
 * positions (:895-1140), Local_Alignment of the read against itself.  Appends one record per alignment, aread == bread,
 * per read in alignment order (unfused, like ora_report).  p->kmer/binshift/hitmin/minover/tspace/ave_corr as datander sets them. */
void ora_match_self(const ora_block *blk, const ora_params *p, ora_result *res);

/* complement_DB (daligner.c:529): returns a malloc'd mirror (bases+mask); free with ora_free_block */
ora_block *ora_complement_block(const ora_block *blk);

/* Compute_Trace_PTS in its cost form (align.c:5577-5692; iter_np returns the unit-cost edit distance, :5230): the exact
 * edit distance of every trace-point segment of every record, by the textbook O(M*N) recurrence.  `recs` are raw 40-byte
 * .las records (tlen, diffs, abpos, bbpos, aepos, bepos, flags, aread, bread, pad), `toff[r]` the index of record r's
 * first trace value in `trace` (uint16 values), cblk = ora_complement_block(bblk) for the comp records (or NULL).
 * rec_diffs[r] = sum over the segments (path->diffs after the call), seg (may be NULL) one value per segment at toff[r]/2 + i. */
void ora_trace_point_diffs(const ora_block *ablk, const ora_block *bblk, const ora_block *cblk, const int32_t *recs,
                           const uint64_t *toff, const uint16_t *trace, int64_t novl, int tspace,
                           int32_t *rec_diffs, uint16_t *seg);
void       ora_free_block(ora_block *blk);

/* one Local_Alignment call (align.c:1904) on raw sequences; for unit tests.
 * out[0..5] = abpos,bbpos,aepos,bepos,diffs,atlen ; bout likewise for the B path.
 * atrace/btrace need room for 4*(max(alen,blen)/tspace+2) values each.                   */
void ora_local_alignment(const int8_t *aseq, int alen, const int8_t *bseq, int blen,
                         int comp, int selfie, int diag, int anti, const ora_params *p,
                         int32_t *aout, uint16_t *atrace, int32_t *bout, uint16_t *btrace);

/* result handling */
ora_result *ora_new_result(void);
void        ora_free_result(ora_result *r);
void        ora_reset_result(ora_result *r);
/* Write_Overlap_Buffer (align.c:6166): sorts (SORT_OVL) and serialises. If split_read >= 0 two
 * files are produced (records with aread <= split_read go to the first).  Returns malloc'd
 * buffers via out1/out2 (out2 NULL when not split). Padding bytes 36..39 are written as 0. */
void ora_serialise_las(ora_result *r, const ora_params *p, int only_identity, int split_read,
                       uint8_t **out1, int64_t *len1, uint8_t **out2, int64_t *len2);
void ora_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
