#!/bin/bash
mkdir -p gpurun_out
NG=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29741 tests/slab_parity_ranks.py 48 > gpurun_out/slab_parity_$NG.log 2>&1; echo "slab parity ($NG ranks, p2p) rc=$?"
grep -E "FAIL|OK|flipped|Error|error" gpurun_out/slab_parity_$NG.log | head -12
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.1f ms" % d["e2e"]["ms_per_step"], "planes", d["config"].get("slab_planes"))
    print("    ", {k: round(v, 2) for k, v in d["stages_ms"].items()})
except Exception as e:
    print(f, "ERR", e)
PY
}
R2S_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29754 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/scale_${NG}_nccl.json 2> gpurun_out/scale_${NG}_nccl.err; echo "N=$NG NCCL only rc=$?"; show gpurun_out/scale_${NG}_nccl.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_${NG}_nccl.err | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29755 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/scale_$NG.json 2> gpurun_out/scale_$NG.err; echo "N=$NG p2p rc=$?"; show gpurun_out/scale_$NG.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_$NG.err | tail -3
