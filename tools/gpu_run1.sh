#!/bin/bash
# first GPU pass: parity tests, bench at the full config, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --n 128 --steps 3 --warmup 3 > gpurun_out/bench128.json 2> gpurun_out/bench128.err; echo "bench128 rc=$?"
tail -c 3000 gpurun_out/bench128.json; tail -5 gpurun_out/bench128.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; echo "bench256 rc=$?"
tail -c 3000 gpurun_out/bench256.json; tail -5 gpurun_out/bench256.err
