#!/bin/bash
# parity tests + ncu launch list + full captures of the heavy kernels at n=128
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
CMD="python bench.py --n 128 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain128.log 2>&1 || { echo plain failed; tail gpurun_out/plain128.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1_n128.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_project_hex8|k_assemble|k_sign" -s 4 -c 4 -o gpurun_out/prof_dist_sign -f $CMD > gpurun_out/ncu_a.log 2>&1; echo "ncu A rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_fine_eval|k_stencil81|k_cg_update" -s 30 -c 6 -o gpurun_out/prof_rbf -f $CMD > gpurun_out/ncu_b.log 2>&1; echo "ncu B rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_vol_|k_cc_" -s 95 -c 8 -o gpurun_out/prof_vol_cc -f $CMD > gpurun_out/ncu_c.log 2>&1; echo "ncu C rc=$?"
ls -la gpurun_out/*.ncu-rep
