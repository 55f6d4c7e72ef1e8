#!/bin/bash
# A/B of the projection kernel's occupancy target (R2S_PL_MINB, r2s_dist.cu): libr2s_m<k>.so are extra builds linked by hand
# (nvcc -DR2S_PL_MINB=k -c r2s_dist.cu; link with the other objects).  Runs the parity suite on the default build first.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
run() { timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench $1 rc=$?"
  python -c "import json; d=json.loads(open('gpurun_out/bench_$1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages_ms'], d['report'].get('ms_solve'))"; tail -2 gpurun_out/bench_$1.err; }
run m5
for k in 4 3; do
  [ -f rho2sdf.jl_b200/libr2s_m$k.so ] || continue
  cp rho2sdf.jl_b200/libr2s.so /tmp/libr2s_keep.so; cp rho2sdf.jl_b200/libr2s_m$k.so rho2sdf.jl_b200/libr2s.so
  run m$k
  cp /tmp/libr2s_keep.so rho2sdf.jl_b200/libr2s.so
done
