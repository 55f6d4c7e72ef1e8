#!/bin/bash
# First GPU call of the next round: the opt-in projection variants (FAST restoration, phase-1 table) that were written after the
# round-1 GPU budget ended.  Parity tests with the opt-in tests enabled, then the variants side by side on the bench workload.
mkdir -p gpurun_out
R2S_TEST_OPTIN=1 timeout 400 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_gpu_optin.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_optin.log
timeout 200 python tools/ab_project.py --steps 3 > gpurun_out/ab_project_r2.jsonl 2> gpurun_out/ab_project_r2.err; echo "ab rc=$?"; cut -c1-330 gpurun_out/ab_project_r2.jsonl; tail -2 gpurun_out/ab_project_r2.err
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pipelined-e2e > gpurun_out/bench_r2_pipelined.json 2> gpurun_out/bench_r2_pipelined.err; echo "bench rc=$?"; python -c "import json; d=json.loads(open('gpurun_out/bench_r2_pipelined.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_pipelined'))"
