#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms/step %.2f" % d["ms_per_step"], "e2e %.1f ms" % d["e2e"]["ms_per_step"], {k: round(v, 2) for k, v in d["stages_ms"].items()})
except Exception as e:
    print(f, "ERR", e)
PY
}
for sv in 1 2; do
R2S_STENCIL=$sv timeout 600 python bench.py --n 128 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench128_s$sv.json 2> gpurun_out/bench128_s$sv.err; echo "stencil=$sv rc=$?"
show gpurun_out/bench128_s$sv.json; tail -1 gpurun_out/bench128_s$sv.err
done
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench256.json 2> gpurun_out/bench256.err; echo "bench256 rc=$?"
show gpurun_out/bench256.json; tail -3 gpurun_out/bench256.err
