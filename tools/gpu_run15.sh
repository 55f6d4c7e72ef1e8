#!/bin/bash
mkdir -p gpurun_out
NG=${1:-2}
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "cg", round(d["stages_ms"]["ms_cg"], 2), "cg_iteration_ms", d.get("cg_iteration_ms"))
    for r in d.get("per_rank", []): print("    ", {k: r[k] for k in ("ms_cg", "cg_matvec", "cg_xchg1", "cg_update", "cg_xchg2", "planes")})
except Exception as e:
    print(f, "ERR", e)
PY
}
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/probe_1.json 2> gpurun_out/probe_1.err; echo "N=1 rc=$?"; show gpurun_out/probe_1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29755 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/probe_$NG.json 2> gpurun_out/probe_$NG.err; echo "N=$NG p2p rc=$?"; show gpurun_out/probe_$NG.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/probe_$NG.err | tail -3
R2S_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29756 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/probe_${NG}_nccl.json 2> gpurun_out/probe_${NG}_nccl.err; echo "N=$NG nccl rc=$?"; show gpurun_out/probe_${NG}_nccl.json
