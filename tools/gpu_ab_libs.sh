#!/bin/bash
# A/B of compile-time variants: every rho2sdf.jl_b200/libr2s_<tag>.so (extra builds linked by hand, e.g.
#   nvcc ... -DR2S_PL_MINB=5 -c r2s_dist.cu -o /tmp/d.o; nvcc -shared -o ../libr2s_m5.so <other objects> /tmp/d.o -lcudart -ldl)
# is benchmarked after the default build.  Usage: gpu_ab_libs.sh [pytest-args]   (runs the parity suite on the default build first)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf -x $* > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
run() { timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench $1 rc=$?"
  python -c "import json; d=json.loads(open('gpurun_out/bench_$1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages_ms'], d['report'].get('ms_solve'))"; tail -2 gpurun_out/bench_$1.err; }
run default
cp rho2sdf.jl_b200/libr2s.so /tmp/libr2s_keep.so
for f in rho2sdf.jl_b200/libr2s_*.so; do
  [ -f "$f" ] || continue
  tag=${f##*libr2s_}; tag=${tag%.so}
  cp $f rho2sdf.jl_b200/libr2s.so; run $tag
done
cp /tmp/libr2s_keep.so rho2sdf.jl_b200/libr2s.so
