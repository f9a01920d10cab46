#!/bin/bash
# One gpurun call: default bench (with host profile), then the ncu launch list of a one-step bench and a
# full capture of the dominant HBM kernel (rs_scatter) on one C1 job.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
DAMAR_PROFILE=1 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || exit 1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4440 -c 1480 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
python tools/prof_job.py c1 1 > gpurun_out/plain_job.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rs_scatter -s 12 -c 3 -o gpurun_out/prof_scatter_r1 -f python tools/prof_job.py c1 1 > gpurun_out/ncu_scatter.log 2>&1
ls -la gpurun_out
