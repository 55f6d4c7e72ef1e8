#!/bin/bash
# DRAM traffic of the dominant kernel at the bench configuration (n = 256)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain256.log 2>&1 || { echo plain failed; tail gpurun_out/plain256.log; exit 1; }
tail -c 600 gpurun_out/plain256.log
ncu --set full --clock-control none --import-source on -k regex:"k_project_hex8" -s 1 -c 1 -o gpurun_out/prof_r1d_project_n256 -f $CMD > gpurun_out/ncu_t.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_r1d_project_n256.ncu-rep
