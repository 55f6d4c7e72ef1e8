#!/bin/bash
mkdir -p gpurun_out
NG=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29741 tests/slab_parity_ranks.py 48 > gpurun_out/slab_parity_$NG.log 2>&1; echo "slab parity ($NG ranks) rc=$?"
grep -E "FAIL|OK|flipped|Error" gpurun_out/slab_parity_$NG.log | head -16
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print("   ", {k: round(v, 2) for k, v in d["stages_ms"].items()})
except Exception as e:
    print(f, "ERR", e)
PY
}
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "N=1 rc=$?"; show gpurun_out/scale_1.json; tail -2 gpurun_out/scale_1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29754 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/scale_$NG.json 2> gpurun_out/scale_$NG.err; echo "N=$NG rc=$?"; show gpurun_out/scale_$NG.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_$NG.err | tail -3
