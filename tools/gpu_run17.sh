#!/bin/bash
mkdir -p gpurun_out
NG=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29758 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/scale_${NG}_final.json 2> gpurun_out/scale_${NG}_final.err; echo "N=$NG rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/scale_8_final.json").read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v, 2) for k, v in d["stages_ms"].items()}); print(d["config"]["slab_planes"], d["cg_iteration_ms"])
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_${NG}_final.err | tail -3
