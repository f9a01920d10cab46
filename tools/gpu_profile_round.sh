#!/bin/bash
# Two gpurun calls (1 GPU each), outputs under gpurun_out/ (kept small: the .ncu-rep files are summarised on the box and removed).
#   tools/gpu_profile_round.sh bench   : default bench (reference plan + byte parity inside), the reference arm, the sort alone
#   tools/gpu_profile_round.sh ncu     : ncu launch list of a one-step bench, full captures of the dominant kernels on one job
set -x
mkdir -p gpurun_out
if [ "$1" = "bench" ]; then
  DAMAR_PROFILE=1 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err || { tail -20 gpurun_out/r2_bench_default.err; exit 1; }
  python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
  python tools/sort_bench.py > gpurun_out/r2_sortbench.log 2>&1
else
  CMD="python bench.py --steps 1 --warmup 2 --no-cpu-baseline"
  $CMD > gpurun_out/r2_plain_bench.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/r2_ncu_bench.log 2>&1
  python tools/launch_list.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench.txt && rm -f gpurun_out/r2_launches_bench.csv
  JOB="python tools/prof_job.py c2 2 1"
  DAMAR_WORKERS=1 $JOB > gpurun_out/r2_plain_job_c2.log 2>&1 || exit 1
  for spec in "os_pass 3" "os_hist 1" "merge_tiles 2" "wave_tasks 2" "filter_pairs 1" "finish_las 1" "extract_tuples 1" "align_pairs_strided 1"; do
    set -- $spec
    DAMAR_WORKERS=1 ncu --set full --clock-control none -k regex:"$1" -c $2 -o gpurun_out/tmp_$1 -f $JOB > gpurun_out/r2_ncu_$1.log 2>&1
    python tools/ncu_summary.py gpurun_out/tmp_$1.ncu-rep > gpurun_out/r2_ncu_$1.txt 2>&1
    rm -f gpurun_out/tmp_$1.ncu-rep
  done
fi
du -sh gpurun_out; ls -la gpurun_out | tail -25
