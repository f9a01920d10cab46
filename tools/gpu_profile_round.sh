#!/bin/bash
# One gpurun call (1 GPU): the default bench (with the reference plan + byte parity), the reference arm, the sort alone,
# then the ncu launch list of a one-step bench and full captures of the dominant kernels on one C1 job.  Outputs: gpurun_out/.
set -x
mkdir -p gpurun_out
DAMAR_PROFILE=1 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err || { tail -20 gpurun_out/r2_bench_default.err; exit 1; }
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python tools/sort_bench.py > gpurun_out/r2_sortbench.log 2>&1
CMD="python bench.py --steps 1 --warmup 2 --no-cpu-baseline"
$CMD > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/r2_ncu_bench.log 2>&1
DAMAR_WORKERS=1 python tools/prof_job.py c1 1 > gpurun_out/r2_plain_job.log 2>&1 &&
DAMAR_WORKERS=1 ncu --set full --clock-control none --import-source on -k regex:"wave_tasks|os_pass|merge_tiles|os_hist" -c 14 -o gpurun_out/prof_r02 -f python tools/prof_job.py c1 1 > gpurun_out/r2_ncu_full.log 2>&1
DAMAR_WORKERS=1 python tools/prof_job.py c2 2 1 > gpurun_out/r2_plain_job_c2.log 2>&1 &&
DAMAR_WORKERS=1 ncu --set full --clock-control none --import-source on -k regex:"merge_tiles|merge_bounds|extract_tuples|filter_pairs|finish_las|dsink" -c 12 -o gpurun_out/prof_r02_merge -f python tools/prof_job.py c2 2 1 > gpurun_out/r2_ncu_merge.log 2>&1
ls -la gpurun_out | tail -20
