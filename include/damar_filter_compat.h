/*
 * damar_filter_compat.h -- the reference-facing face of libdamar_b200: the three entry points of
 * dalign/filter.h:64-70 with the reference's own signatures, for a host that is the UNMODIFIED
 * dalign/daligner.c.  Link damar_b200/csrc/filter_compat.c + libdamar_b200.so in place of
 * dalign/filter.c; nothing else in the reference tree changes.
 *
 * The struct declarations below restate the in-core layouts the reference hands across that
 * boundary (db/DB.h:319-325 HITS_READ, :327-347 HITS_TRACK, :361-389 HITS_DB; dalign/align.h:125-132
 * Path, :358-364 Overlap, :402-414 Overlap_IO_Buffer) so the shim compiles without the reference
 * headers; they are layout-compatible on LP64 and are checked with static asserts.
 */
#ifndef DAMAR_FILTER_COMPAT_H
#define DAMAR_FILTER_COMPAT_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int      rlen;          /* length of the read                                   */
    int64_t  boff;          /* in-memory offset of the read in `bases`              */
    int64_t  coff;
    int      flags;
} DAMAR_HITS_READ;          /* 32 bytes */

typedef struct damar_hits_track {
    struct damar_hits_track *next;
    char  *name;
    int    size;            /* 8: anno is int64[nreads+1]                           */
    void  *anno;
    void  *data;
} DAMAR_HITS_TRACK;

typedef struct {
    int      ureads;
    float    freq[4];
    int      maxlen;
    int64_t  totlen;
    int      nreads;
    int      part;
    int      ufirst;
    char    *path;
    int      loaded;
    void    *bases;
    DAMAR_HITS_READ  *reads;
    DAMAR_HITS_TRACK *tracks;
} DAMAR_HITS_DB;            /* 88 bytes */

typedef struct {
    void *trace;
    int   tlen;
    int   diffs;
    int   abpos, bbpos;
    int   aepos, bepos;
} DAMAR_Path;

typedef struct {
    DAMAR_Path path;
    uint32_t   flags;
    int        aread;
    int        bread;
} DAMAR_Overlap;            /* 48 bytes */

typedef struct DAMAR_Overlap_IO_Buffer DAMAR_Overlap_IO_Buffer;   /* opaque: only passed back to the host */
typedef void DAMAR_Align_Spec;

/* ---- what the library exports to the reference host (dalign/filter.h:64-70) ---- */
int   Set_Filter_Params(int kmer, int binshift, int suppress, int hitmin, int nthreads);
void *Sort_Kmers(DAMAR_HITS_DB *block, int *len);
void  Match_Filter(char *aname, DAMAR_HITS_DB *ablock, char *bname, DAMAR_HITS_DB *bblock,
                   void *atable, int alen, void *btable, int blen, int comp, DAMAR_Align_Spec *asettings);

/* ---- what the shim imports from the reference host ---- */
/* globals of dalign/daligner.c:131-140 */
extern int VERBOSE, BIASED, MINOVER, HGAP_MIN, SYMMETRIC, IDENTITY;
extern unsigned long long MEM_LIMIT, MEM_PHYSICAL;
/* accessors of dalign/align.c:318-345, 5935-5967 and the buffer append of :6020 */
int      Trace_Spacing(DAMAR_Align_Spec *spec);
double   Average_Correlation(DAMAR_Align_Spec *spec);
float   *Base_Frequencies(DAMAR_Align_Spec *spec);
DAMAR_Overlap_IO_Buffer *OVL_IO_Buffer(DAMAR_Align_Spec *spec);
int      AddOverlapToBuffer(DAMAR_Overlap_IO_Buffer *iobuf, DAMAR_Overlap *ovl, int tbytes);
/* db/DB.c:726 */
long long sizeof_DB(DAMAR_HITS_DB *db);

#ifdef __cplusplus
}
#endif
#endif
