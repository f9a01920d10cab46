/*
 * damar_b200.h -- C-ABI of the B200-native daligner overlap kernel.
 *
 * This is the drop-in boundary for the one hot path of MartinPippel/DAmar that this
 * repository accelerates: the three entry points of dalign/filter.h:64-70
 *
 *     int   Set_Filter_Params(int kmer, int binshift, int suppress, int hitmin, int nthreads);
 *     void *Sort_Kmers(HITS_DB *block, int *len);
 *     void  Match_Filter(char *aname, HITS_DB *ablock, char *bname, HITS_DB *bblock,
 *                        void *atable, int alen, void *btable, int blen,
 *                        int comp, Align_Spec *asettings);
 *
 * restated with plain pointers and sizes (no HITS_DB / Align_Spec / torch types), so that
 * any host (the reference's daligner.c through the shim in include/damar_filter_compat.h,
 * our own C++ driver, or Python via ctypes) can bind it.  Everything below runs on the GPU;
 * there is no CPU fallback: every call fails loudly (message on stderr + non-zero return /
 * NULL) when no CUDA device is usable.
 *
 * Reference citations are relative to the reference root (dalign/..., db/...).
 */
#ifndef DAMAR_B200_H
#define DAMAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAMAR_B200_ABI_VERSION 1

/* ---- host view of one DB block (what read_DB leaves in a HITS_DB, daligner.c:442-509) ---- */
typedef struct damar_block_desc {
    int32_t        nreads;      /* HITS_DB.nreads                                            */
    int32_t        maxlen;      /* HITS_DB.maxlen                                            */
    int64_t        totlen;      /* HITS_DB.totlen                                            */
    int32_t        ufirst;      /* HITS_DB.ufirst : DB-global id of read 0                   */
    int32_t        part;        /* HITS_DB.part   : block number (0 = whole DB)              */
    const uint8_t *bases;       /* HITS_DB.bases  : 1 byte/base (0..3), 4 after every read   */
    const int64_t *boff;        /* [nreads+1] in-memory offsets  (HITS_READ.boff)            */
    const int32_t *rlen;        /* [nreads]                       (HITS_READ.rlen)            */
    const int64_t *mask_anno;   /* merged mask track: [nreads+1] offsets in ints, or NULL    */
    const int32_t *mask_data;   /* sorted [beg,end) pairs (daligner.c:474-497)               */
    float          freq[4];     /* HITS_DB.freq                                              */
    int64_t        sizeof_db;   /* sizeof_DB(block), db/DB.c:726 (enters the -M rule)        */
    int32_t        flags;       /* DAMAR_BLOCK_* */
    int32_t        reserved;
} damar_block_desc;

/* The host already applied complement_DB (daligner.c:529) where it wanted a complemented block:
 * damar_match_filter(comp=1) then uses the block as shipped instead of complementing on the device. */
#define DAMAR_BLOCK_AS_IS 1

/* ---- the globals of daligner.c:131-140 that filter.c reads, plus the Align_Spec fields ---- */
typedef struct damar_options {
    int32_t  verbose;           /* VERBOSE                                                   */
    int32_t  minover;           /* MINOVER, already doubled (daligner.c:861)                 */
    int32_t  hgap_min;          /* HGAP_MIN (always 0 in DAmar, daligner.c:688)              */
    int32_t  symmetric;         /* SYMMETRIC (0 with -A or across DBs)                       */
    int32_t  identity;          /* IDENTITY (-I / -O)                                        */
    int32_t  trace_spacing;     /* Trace_Spacing(spec)          (-s)                         */
    int32_t  no_trace;          /* -T                                                        */
    int32_t  reserved;
    double   ave_corr;          /* Average_Correlation(spec)    (-e)                         */
    float    freq[4];           /* Base_Frequencies(spec) = A block freq (daligner.c:951)    */
    uint64_t mem_limit;         /* MEM_LIMIT in bytes, 0 == -M0                              */
} damar_options;

/* ---- one overlap as Write_Overlap serialises it (align.c:3345-3373): 40 bytes ---- */
typedef struct damar_overlap {
    int32_t  tlen, diffs, abpos, bbpos, aepos, bepos;
    uint32_t flags;
    int32_t  aread, bread;
    int32_t  pad;               /* always 0 (the reference leaves struct padding here)       */
} damar_overlap;

typedef struct damar_overlaps {
    int64_t              novl;
    const damar_overlap *ovl;        /* novl records, in no particular order                  */
    const uint64_t      *trace_off;  /* novl offsets (in values) into trace                   */
    const uint64_t      *seq;        /* novl emission ranks: the order in which the reference's
                                        report threads would have buffered the records (pair order,
                                        A records before B records); breaks full SORT_OVL ties      */
    const void          *trace;      /* trace values (d_i, b_i pairs): uint8 if trace_bytes == 1
                                        (trace_spacing <= 125, Compress_TraceTo8), else uint16      */
    int64_t              ntrace;
    int32_t              trace_bytes;
    int32_t              reserved;
    /* counters printed by -v (filter.c:2910-2928) */
    int64_t              nhits;      /* seed hits (k-mer matches)                             */
    int64_t              nfilt;      /* Local_Alignment calls                                 */
    int64_t              ncheck;     /* records                                               */
    int32_t              limit;      /* k-mer match cap chosen by the -M rule                 */
    int32_t              multi_la_pairs;
} damar_overlaps;

typedef struct damar_block damar_block;      /* device-resident block                         */
typedef struct damar_index damar_index;      /* device-resident sorted k-mer list             */

/* library / device life cycle.  damar_init returns 0 on success; every other entry point
 * calls it implicitly with device = $DAMAR_DEVICE or LOCAL_RANK or 0.                          */
int         damar_abi_version(void);
int         damar_init(int device);
void        damar_shutdown(void);
const char *damar_last_error(void);

/* Worker contexts: independent comparisons (block pair x orientation) may be in flight at once, one per
 * host thread.  A thread binds itself to context `index` (0 .. damar_worker_count()-1; 0 is the default
 * every thread starts with); each context has its own CUDA stream, device-memory cache and result slots.
 * Blocks and indices may be shared read-only between contexts once the call that created them returned. */
int damar_worker_bind(int index);
int damar_worker_count(void);

/* Set_Filter_Params (filter.c:171-201).  Returns 1 for an illegal combination. */
int  damar_set_filter_params(int kmer, int binshift, int suppress, int hitmin, int nthreads);
void damar_set_options(const damar_options *opt);
/* BIASED (-b, daligner.c:132/729): variable-length k-mers by base log-frequency (biased_tuple_thread,
 * filter.c:549-688).  The per-base weights are fixed from the freq of the block of the FIRST
 * damar_sort_kmers call that follows damar_set_filter_params / damar_set_biased (filter.c:774-789). */
void damar_set_biased(int biased);

/* block residency: ships the block to HBM and packs it to 2 bit/base.                          */
damar_block *damar_block_upload(const damar_block_desc *desc);
void         damar_block_free(damar_block *blk);

/* A resident block as device pointers, for device-to-device transfer between the GPUs of one box (the multi-GPU driver
 * broadcasts the PACKED block -- 2 bit/base -- over NCCL instead of shipping the unpacked bases through the host).
 * damar_block_view_get fills the view of a resident block (pointers stay owned by the block); damar_block_from_view
 * builds a new resident block from device buffers laid out the same way (device-to-device copies, no host bounce).     */
typedef struct damar_block_view {
    int32_t  nreads, maxlen, ufirst, part;
    int64_t  totlen, nbases, sizeof_db;
    float    freq[4];
    int64_t  npacked_words;   /* 32-bit words of packed bases, the 4 zero words of front padding NOT counted           */
    void    *packed;          /* device: uint32[4 + npacked_words], front padding included                             */
    void    *boff;            /* device: uint32[nreads + 1]                                                            */
    void    *rlen;            /* device: int32[nreads]                                                                 */
    int64_t  mask_len;        /* number of int32 in mask_data; mask_anno == NULL: no mask                               */
    void    *mask_anno;       /* device: int64[nreads + 1]                                                             */
    void    *mask_data;       /* device: int32[mask_len]                                                               */
    int32_t  flags, reserved;
} damar_block_view;
int          damar_block_view_get(damar_block *blk, damar_block_view *view);
damar_block *damar_block_from_view(const damar_block_view *view);

/* Sort_Kmers (filter.c:753-994).  comp != 0 builds the index of the reverse-complemented block
 * (what daligner.c:998-1003 obtains with complement_DB + Sort_Kmers) without touching host memory.
 * *len receives the k-mer count; NULL/0 when the block has no k-mers, NULL/-1 on failure (error set).
 *
 * Device memory: every call below that allocates runs alongside the calls of the other worker threads; one that runs out of
 * device memory is unwound, waits for the calls in flight, hands all cached memory back to the driver and is run again with
 * the device to itself.  Only a call that does not fit the device alone fails ("out of device memory").            */
damar_index *damar_sort_kmers(damar_block *blk, int comp, int64_t *len);
void         damar_index_free(damar_index *idx);

/* Match_Filter (filter.c:2519-2929) for one orientation.  `self` is the reference's
 * MG_self (aname == bname).  Results are owned by the library and stay valid until the next
 * damar_match_filter / damar_shutdown call.  Returns 0 on success.                             */
int damar_match_filter(damar_block *ablk, damar_block *bblk, damar_index *aidx, damar_index *bidx,
                       int self, int comp, damar_overlaps *out);

/* datander (scrub/datander.c:52-66): Match_Self (scrub/tandem.c:1182-1428), every read of the block against itself for
 * tandem repeats.  Parameters as datander sets them: damar_set_filter_params(kmer 12, binshift 4, suppress 0, hitmin 35,
 * nthreads) = tandem.c:137, and damar_options {minover = 2 * -l (500), trace_spacing, ave_corr, freq}.  One record per
 * alignment (aread == bread, flags 0); results are owned by the library like damar_match_filter's.  Returns 0 on success. */
int damar_match_self(damar_block *blk, damar_overlaps *out);

/* order[0..novl) = record indices sorted by emission rank (the order AddOverlapToBuffer would have
 * been called in by the reference's report threads)                                              */
void damar_emission_order(const damar_overlaps *o, int64_t *order);

/* Compute_Trace_PTS (align.c:5577-5692, called by the consumers of a .las file: utils/LAshow.c:261, LAcorrect, LAstitch)
 * in its cost form: the exact edit distance of every trace-point segment of every record, i.e. what iter_np returns
 * per segment (align.c:5230) and what Compute_Trace_PTS leaves in path->diffs per record -- not the edit script.
 * Records and traces are given as damar_match_filter / a .las file hold them (aread of ablk, bread of bblk, DB-global
 * ids; B coordinates of COMP records in the complemented read).  rec_diffs[novl] receives the per-record sums (-1:
 * "trace point out of bounds", align.c:5667, or a record without trace points); seg_diffs (may be NULL) one value per
 * segment, segment i of record r at trace_off[r] / 2 + i.  Returns 0, or 1 when any record was rejected.            */
int damar_trace_point_diffs(damar_block *ablk, damar_block *bblk, const damar_overlap *ovl, const uint64_t *trace_off,
                            const void *trace, int trace_bytes, int64_t novl, int64_t ntrace, int trace_spacing,
                            int32_t *rec_diffs, uint16_t *seg_diffs);

/* Compute_Trace_PTS proper: the cost form above PLUS the edit script the reference leaves in path->trace (align.c:5686-5688),
 * mode GREEDIEST (the mode LAshow and filter.c's bridging use): per record the concatenation over its segments of iter_np's
 * output (align.c:5238-5254) -- one value per inserted / deleted base, positive = B coordinate + 1 where an A base stands
 * alone, negative = -(A coordinate + 1) where a B base stands alone; substitutions write nothing.  script_off[novl + 1]
 * receives the records' offsets into *script (malloc'd, free with damar_host_free); rec_diffs as for damar_trace_point_diffs.
 * Returns 0, 1 when a record was rejected (its script is empty), 2 on failure.                                          */
int damar_trace_point_scripts(damar_block *ablk, damar_block *bblk, const damar_overlap *ovl, const uint64_t *trace_off,
                              const void *trace, int trace_bytes, int64_t novl, int64_t ntrace, int trace_spacing,
                              int32_t *rec_diffs, int64_t *script_off, int32_t **script, int64_t *script_len);
/* Compute_Trace_MID (align.c:5694-5830, mode GREEDIEST as corrector/LAcorrect.c:545 calls it): the same outputs, but iter_np
 * runs between the MIDDLES (middle_np, align.c:5263-5566) of successive trace-point segments instead of between the trace
 * points -- the nearer-optimal trace align.h:271-277 describes.  rec_diffs receives what the reference leaves in path->diffs,
 * including its double count of the piece that ends at the last middle (align.c:5814-5817).  A piece that needs more waves
 * than the largest d_i of its record is the reference's fatal "Bad alignment between trace points": returns 2.            */
int damar_trace_mid_scripts(damar_block *ablk, damar_block *bblk, const damar_overlap *ovl, const uint64_t *trace_off,
                            const void *trace, int trace_bytes, int64_t novl, int64_t ntrace, int trace_spacing,
                            int32_t *rec_diffs, int64_t *script_off, int32_t **script, int64_t *script_len);
void damar_host_free(void *p);

/* ---- stage-level read-backs (used by the parity tests; not needed by a host) ---- */
/* sorted k-mer list in the reference's KmerPos layout {u64 code; i32 rpos; i32 read}           */
int damar_index_download(const damar_index *idx, void *kmerpos_out, int64_t capacity);
/* sorted seed hits of the last damar_match_filter call in SeedPair layout {diag,apos,aread,bread} */
int64_t damar_last_hits_download(void *seedpair_out, int64_t capacity);
/* per-stage device timings (ms) of the last damar_sort_kmers / damar_match_filter call          */
typedef struct damar_timings {
    float upload_ms, extract_ms, ksort_ms;
    float count_ms, merge_ms, hsort_ms, pairs_ms, align_ms, download_ms;
    int64_t kmers, hits, pairs, la_calls, wave_cells, kernel_launches;
    /* accumulated since damar_reset_kernel_timers(): the radix-sort scatter kernel (dominant HBM kernel) */
    double  scatter_ms; int64_t scatter_elems; int64_t scatter_launches;
    double  align_kernel_ms; int64_t align_launches;
    int64_t redo_pairs;          /* pairs whose wave front outgrew one warp (strided second pass) */
    double  redo_kernel_ms;      /* accumulated like align_kernel_ms */
    int64_t scatter_bytes;       /* bytes the sort passes moved (read + write), accumulated like scatter_ms */
    int64_t wide_passes;         /* last call: wave passes whose front outgrew 32 diagonals (parked and resumed by a wide lane) */
    /* accumulated since damar_reset_kernel_timers(): whole sorts (histogram + scan + passes) and the contract bytes of
     * SURVEY section 8d for them (16-byte records, 8-bit digits: 16 n (1 + 2 ceil(bits / 8))); wave cells and passes */
    double  sort_ms; int64_t sort_contract_bytes; int64_t sort_elems;
    int64_t cells_total, wide_total, la_total;
} damar_timings;
/* The device radix sort on its own (stage 1 / stage 2 sorts, filter.c:230-435): n pseudo-random keys are sorted stably on
 * bits [lo_bit, hi_bit) (with a u32 payload when has_vals), `reps` times, and the result is checked ON THE DEVICE for order,
 * stability and being a permutation of the input.  *ms_out = mean device time of one sort.  Returns 0 when every check
 * passed.  Used by the GPU tests and tools/sort_bench.py; not needed by a host.                                           */
int  damar_sort_selftest(int64_t n, int lo_bit, int hi_bit, int has_vals, int reps, uint64_t seed, int skew, double *ms_out);
void damar_reset_kernel_timers(void);
void damar_dump_profile(void);   /* host wall-clock breakdown to stderr */
void damar_get_timings(damar_timings *t);

/* ---- output path: Overlap_IO_Buffer + Write_Overlap_Buffer (align.c:5969-6380), host only ---- */
typedef struct damar_las_sink damar_las_sink;
damar_las_sink *damar_sink_new(void);
void            damar_sink_free(damar_las_sink *s);
void            damar_sink_reset(damar_las_sink *s);                         /* Reset_Overlap_Buffer */
void            damar_sink_add(damar_las_sink *s, const damar_overlaps *o);  /* AddOverlapToBuffer x novl */
/* same, for calls that complete out of order: `ordinal` = position of the Match_Filter call in the job */
void            damar_sink_add_at(damar_las_sink *s, const damar_overlaps *o, int64_t ordinal);
int64_t         damar_sink_count(const damar_las_sink *s);
/* Sort by SORT_OVL (align.c:6104) and serialise to malloc'd .las images (header int64 novl,
 * int32 tspace; 40-byte records; u8 traces when trace_spacing <= 125).  split_read >= 0 writes
 * records with aread <= split_read to buf1 and the rest to buf2 (align.c:6330-6354).
 * Free the buffers with damar_free.  Returns 1 if a trace value does not fit 8 bits.           */
int  damar_sink_serialise(damar_las_sink *s, int trace_spacing, int no_trace, int only_identity, int split_read,
                          uint8_t **buf1, int64_t *len1, uint8_t **buf2, int64_t *len2);
void damar_free(void *p);
/* ---- the same on the device (section 8f-2): records never leave HBM before they are a finished .las image ----
 * damar_match_filter_sink = damar_match_filter whose records are appended to a device sink as call number `ordinal` of the
 * job (0 = B block in normal orientation, 1 = complemented); `out` receives the counters only (no record pointers).
 * damar_dsink_serialise = Write_Overlap_Buffer: SORT_OVL order by a device radix sort, image(s) built in HBM, ONE
 * device-to-host copy per file.  maxlen / maxread = largest read length and DB-wide read id among the two blocks (they
 * size the key fields).  Buffers as for damar_sink_serialise (free with damar_free).                                     */
typedef struct damar_device_sink damar_device_sink;
damar_device_sink *damar_dsink_new(void);
void               damar_dsink_free(damar_device_sink *s);
void               damar_dsink_reset(damar_device_sink *s);
int  damar_match_filter_sink(damar_block *ablk, damar_block *bblk, damar_index *aidx, damar_index *bidx, int self, int comp,
                             damar_device_sink *sink, int ordinal, damar_overlaps *out);
int  damar_dsink_serialise(damar_device_sink *s, int trace_spacing, int no_trace, int only_identity, int split_read,
                           int maxlen, int maxread, uint8_t **buf1, int64_t *len1, uint8_t **buf2, int64_t *len2);
/* LAmerge (utils/LAmerge.c; heap merge utils/LAmergeUtils.c:31-145): merges n sorted .las images (e.g. a block's pair files
 * R.i.R.1.las ... R.i.R.n.las) into one image ordered like SORT_OVL; records are copied verbatim.  Returns 0, 1 for a
 * malformed image, 2 for mixed trace spacings.  Free *out with damar_free.  Host only.                                   */
int  damar_las_merge(int n, const uint8_t *const *bufs, const int64_t *lens, uint8_t **out, int64_t *outlen);
/* The comparator of SURVEY section 8c: two .las images are equal when header, bytes 0..35 of every record and every trace
 * byte agree (bytes 36..39 are struct padding the reference leaves uninitialised, align.c:3345-3346).  Returns -1 when
 * equal, -2 for different lengths / a malformed image, else the offset of the first differing byte.  Host only.          */
int64_t damar_las_compare(const uint8_t *x, int64_t xlen, const uint8_t *y, int64_t ylen);

#ifdef __cplusplus
}
#endif
#endif /* DAMAR_B200_H */
