#!/bin/bash
# launch list + full captures of the heavy kernels at n=128 (current build)
mkdir -p gpurun_out
TAG=${1:-r1d}
CMD="python bench.py --n 128 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain128.log 2>&1 || { echo plain failed; tail gpurun_out/plain128.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}_n128.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_project_hex8|k_faces_crossing|k_assemble|k_sign<|k_cc_merge|k_cc_flatten" -c 8 -o gpurun_out/prof_${TAG}_dist_sign -f $CMD > gpurun_out/ncu_a.log 2>&1; echo "ncu A rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_stencil81_march2|k_cg_update|k_fine_eval2" -s 20 -c 5 -o gpurun_out/prof_${TAG}_rbf -f $CMD > gpurun_out/ncu_b.log 2>&1; echo "ncu B rc=$?"
ls -la gpurun_out/*${TAG}*.ncu-rep
