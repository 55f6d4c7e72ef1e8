#!/bin/bash
# one short GPU call: parity tests, the projection variants side by side, then the bench line of the default build
mkdir -p gpurun_out
timeout 45 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_box.json 2> gpurun_out/bench_box.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_box.json").read().strip().splitlines()[-1])
print("ms/step %.2f value %.3e e2e %.1f ms" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])); print({k: round(v, 2) for k, v in d["stages_ms"].items()}); print(d["roofline"])
PY
tail -2 gpurun_out/bench_box.err
timeout 60 python tools/ab_project.py --steps 2 > gpurun_out/ab_project.jsonl 2> gpurun_out/ab_project.err; echo "ab rc=$?"; cut -c1-400 gpurun_out/ab_project.jsonl; tail -2 gpurun_out/ab_project.err
