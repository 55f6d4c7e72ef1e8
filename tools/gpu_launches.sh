#!/bin/bash
# launch list (ncu duration per launch) of one short bench run of the default build; summary to gpurun_out/<TAG>_launch_summary_n256.txt
mkdir -p gpurun_out
TAG=${1:-r2b}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain256.log 2>&1 || { echo plain failed; tail gpurun_out/plain256.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_${TAG}_n256.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/launches_${TAG}_n256.csv 60 > gpurun_out/${TAG}_launch_summary_n256.txt 2>&1; head -45 gpurun_out/${TAG}_launch_summary_n256.txt
