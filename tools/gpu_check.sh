#!/bin/bash
# GPU regression call: parity suite, the A/B knobs side by side on the bench workload, one short bench run
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_variants.py --steps 2 > gpurun_out/ab_variants.jsonl 2> gpurun_out/ab_variants.err; echo "ab rc=$?"; cut -c1-420 gpurun_out/ab_variants.jsonl; tail -2 gpurun_out/ab_variants.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/bench_check.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages_ms'])"; tail -3 gpurun_out/bench_check.err
