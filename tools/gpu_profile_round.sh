#!/bin/bash
# One gpurun call: default bench (with host profile), the reference arm, then the ncu launch list of a one-step
# bench and full captures of the two dominant kernels (wave_tasks, rs_scatter) on one C1 job.  Outputs: gpurun_out/.
set -x
mkdir -p gpurun_out
DAMAR_PROFILE=1 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
# launches per pass vary a little with the number of lasta rounds: capture from the start of the timed step on
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
DAMAR_WORKERS=1 python tools/prof_job.py c1 1 > gpurun_out/plain_job.log 2>&1 &&
DAMAR_WORKERS=1 ncu --set full --clock-control none --import-source on -k regex:"wave_tasks|rs_scatter" -c 8 -o gpurun_out/prof_r01_final -f python tools/prof_job.py c1 1 > gpurun_out/ncu_final.log 2>&1
ls -la gpurun_out
