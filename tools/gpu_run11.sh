#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms/step %.2f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_project","ms_assemble","ms_sign","ms_cg","ms_cc","ms_total")})
except Exception as e:
    print(f, "ERR", e)
PY
}
for cfg in "4 0" "5 0" "6 0" "4 1" "5 1" "6 1"; do
set -- $cfg
R2S_PROJ_MINB=$1 R2S_PROJ_SMEMA=$2 timeout 600 python bench.py --n 128 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench128_v$1$2.json 2> gpurun_out/bench128_v$1$2.err; echo "minb=$1 smemA=$2 rc=$?"
show gpurun_out/bench128_v$1$2.json; tail -1 gpurun_out/bench128_v$1$2.err
done
