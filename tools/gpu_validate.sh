#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_val.json 2> gpurun_out/bench_val.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_val.json").read().strip().splitlines()[-1])
print("ms/step %.2f value %.3e e2e %.1f ms" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])); print({k: round(v, 2) for k, v in d["stages_ms"].items()})
PY
tail -2 gpurun_out/bench_val.err
