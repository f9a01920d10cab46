/*
 * damar_tandem_compat.h -- the reference-facing face of libdamar_b200 for datander: the two entry points of
 * scrub/tandem.h:57-60 with the reference's own signatures, for a host that is the UNMODIFIED scrub/datander.c.
 * Link damar_b200/csrc/tandem_compat.c + libdamar_b200.so in place of scrub/tandem.c.
 * Struct layouts are those of damar_filter_compat.h (db/DB.h, dalign/align.h).
 */
#ifndef DAMAR_TANDEM_COMPAT_H
#define DAMAR_TANDEM_COMPAT_H

#define Set_Filter_Params damar_compat_unused_Set_Filter_Params     /* filter.h's five-argument form is not linked here */
#define VERBOSE  damar_compat_unused_VERBOSE
#define MINOVER  damar_compat_unused_MINOVER
#include "damar_filter_compat.h"
#undef Set_Filter_Params
#undef VERBOSE
#undef MINOVER

#ifdef __cplusplus
extern "C" {
#endif

/* ---- what the library exports to the reference host (scrub/tandem.h:57-60) ---- */
int  Set_Filter_Params(int kmer, int binshift, int hitmin, int nthreads);
void Match_Self(char *aname, DAMAR_HITS_DB *ablock, DAMAR_Align_Spec *aspec);

/* ---- what the shim imports from the reference host: globals of scrub/datander.c:68-69 ---- */
extern int VERBOSE, MINOVER;

#ifdef __cplusplus
}
#endif
#endif
