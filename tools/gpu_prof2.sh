#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1e}
CMD="python bench.py --n 128 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain128.log 2>&1 || { echo plain failed; tail gpurun_out/plain128.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}_n128.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_sign<|k_fine_eval2|k_vol_rows" -c 4 -o gpurun_out/prof_${TAG}_misc -f $CMD > gpurun_out/ncu_a.log 2>&1; echo "ncu A rc=$?"
ls -la gpurun_out/*${TAG}*
