/*
 * ref_trace_pts.c -- TEST INFRASTRUCTURE (oracle/): a 60-line driver around the UNMODIFIED reference functions
 * Compute_Trace_PTS (dalign/align.c:5577-5692), Read_Overlap/Read_Trace, Load_Read and Complement_Seq, exactly the
 * call sequence of utils/LAshow.c:246-262.  For every record of a .las file it prints the record's own `diffs` and the
 * value Compute_Trace_PTS leaves in path->diffs (the sum of the exact edit distances of the trace-point segments).
 * Built by oracle/Makefile into oracle/_ref/trace_pts_ref; used by tests/golden/make_trace_pts.py to pin
 * ora_trace_point_diffs() and the CUDA kernel.  Nothing of the product links or runs this.
 *
 *      trace_pts_ref <db> <las>            one line per record: aread bread comp diffs-before diffs-after
 *      trace_pts_ref <db> <las> script     the same plus the length of the edit script Compute_Trace_PTS wrote (path->tlen
 *                                          ints) and the script itself (GREEDIEST mode, the one LAshow uses)
 *      trace_pts_ref <db> <las> mid        the same through Compute_Trace_MID (dalign/align.c:5694-5830, GREEDIEST as
 *                                          corrector/LAcorrect.c:545 calls it)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "DB.h"
#include "align.h"

int main(int argc, char *argv[])
{
    HITS_DB db;
    FILE *f;
    int64 novl, i;
    int tspace, tbytes, tmax = 0;
    Overlap ovl;
    Alignment aln;
    Work_Data *work;
    char *abuf, *bbuf;
    uint16 *trace = NULL;

    Prog_Name = "trace_pts_ref";
    if (argc != 3 && argc != 4) { fprintf(stderr, "usage: trace_pts_ref <db> <las> [script|mid]\n"); return 1; }
    if (Open_DB(argv[1], &db)) return 1;
    if ((f = fopen(argv[2], "r")) == NULL) return 1;
    if (fread(&novl, sizeof(int64), 1, f) != 1 || fread(&tspace, sizeof(int), 1, f) != 1) return 1;
    tbytes = (tspace <= TRACE_XOVR) ? sizeof(uint8) : sizeof(uint16);
    work = New_Work_Data();
    abuf = New_Read_Buffer(&db);
    bbuf = New_Read_Buffer(&db);
    for (i = 0; i < novl; i++) {
        Read_Overlap(f, &ovl);
        if (ovl.path.tlen > tmax) { tmax = 1.2 * ovl.path.tlen + 100; trace = (uint16 *)realloc(trace, sizeof(uint16) * tmax); }
        ovl.path.trace = (void *)trace;
        Read_Trace(f, &ovl, tbytes);
        if (tbytes == 1) Decompress_TraceTo16(&ovl);
        Load_Read(&db, ovl.aread, abuf, 0);
        Load_Read(&db, ovl.bread, bbuf, 0);
        aln.path = &(ovl.path); aln.flags = ovl.flags;
        aln.aseq = abuf; aln.bseq = bbuf;
        aln.alen = db.reads[ovl.aread].rlen; aln.blen = db.reads[ovl.bread].rlen;
        if (COMP(ovl.flags)) Complement_Seq(bbuf, aln.blen);
        {
            int before = ovl.path.diffs;
            if (argc == 4 && strcmp(argv[3], "mid") == 0) { if (Compute_Trace_MID(&aln, work, tspace, GREEDIEST)) return 1; }
            else if (Compute_Trace_PTS(&aln, work, tspace, GREEDIEST)) return 1;
            printf("%d %d %d %d %d", ovl.aread, ovl.bread, COMP(ovl.flags) ? 1 : 0, before, ovl.path.diffs);
            if (argc == 4) {
                int t, *scr = (int *)ovl.path.trace;
                printf(" %d", ovl.path.tlen);
                for (t = 0; t < ovl.path.tlen; t++) printf(" %d", scr[t]);
            }
            printf("\n");
        }
    }
    return 0;
}
