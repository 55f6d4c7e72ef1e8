#!/bin/bash
mkdir -p gpurun_out
R2S_PROJ=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sphere or simp or mat_fixtures or block" > gpurun_out/pytest_refill.log 2>&1; echo "pytest(refill) rc=$?"
tail -5 gpurun_out/pytest_refill.log
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_project","ms_assemble","ms_total")}, "iters", d["report"]["newton_iters"])
except Exception as e:
    print(f, "ERR", e)
PY
}
for mb in 2 3 4; do
R2S_PROJ=1 R2S_PROJ_MINB=$mb timeout 600 python bench.py --n 128 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench128_rf$mb.json 2> gpurun_out/bench128_rf$mb.err; echo "bench128 refill minb=$mb rc=$?"
show gpurun_out/bench128_rf$mb.json; tail -2 gpurun_out/bench128_rf$mb.err
done
R2S_PROJ=1 R2S_PROJ_MINB=3 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench256_rf.json 2> gpurun_out/bench256_rf.err; echo "bench256 refill rc=$?"
show gpurun_out/bench256_rf.json; tail -3 gpurun_out/bench256_rf.err
