#!/bin/bash
mkdir -p gpurun_out
NG=${1:-4}
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "planes", d["config"].get("slab_planes"))
    for r in d.get("per_rank", []): print("    ", r)
except Exception as e:
    print(f, "ERR", e)
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29754 bench.py --gpus $NG --steps 3 --warmup 3 --no-balance > gpurun_out/scale_${NG}_nobal.json 2> gpurun_out/scale_${NG}_nobal.err; echo "N=$NG no-balance rc=$?"; show gpurun_out/scale_${NG}_nobal.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_${NG}_nobal.err | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29755 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/scale_$NG.json 2> gpurun_out/scale_$NG.err; echo "N=$NG balanced rc=$?"; show gpurun_out/scale_$NG.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_$NG.err | tail -3
