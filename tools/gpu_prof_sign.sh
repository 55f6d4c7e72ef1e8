#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --n 128 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"k_sign<|k_assemble<false, 8, true>|k_assemble" -s 1 -c 3 -o gpurun_out/prof_r1e_sign -f $CMD > gpurun_out/ncu_s.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_r1e_sign.ncu-rep
