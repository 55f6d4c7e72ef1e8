#!/bin/bash
mkdir -p gpurun_out
NG=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29741 tests/slab_parity_ranks.py 48 > gpurun_out/slab_parity.log 2>&1; echo "slab parity rc=$?"
grep -E "FAIL|OK|flipped|Error|error" gpurun_out/slab_parity.log | head -20; tail -5 gpurun_out/slab_parity.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print({k: round(v, 2) for k, v in d["stages_ms"].items()})
    print(d["report"]); print(d.get("clocks"))
except Exception as e:
    print(f, "ERR", e)
PY
}
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench256_1.json 2> gpurun_out/bench256_1.err; echo "bench256 N=1 rc=$?"
show gpurun_out/bench256_1.json; tail -3 gpurun_out/bench256_1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench256_$NG.json 2> gpurun_out/bench256_$NG.err; echo "bench256 N=$NG rc=$?"
show gpurun_out/bench256_$NG.json; tail -5 gpurun_out/bench256_$NG.err
