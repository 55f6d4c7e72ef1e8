#!/bin/bash
# 8-GPU box: slab parity on 8 ranks (peer-memory mailbox), the in-process multi test on 8 devices, bench lines at N = 8 and 4
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 tests/slab_parity_ranks.py 48 > gpurun_out/slab_parity_8.log 2>&1; echo "slab parity (8 ranks) rc=$?"
grep -E "FAIL|OK|Error" gpurun_out/slab_parity_8.log | head -8
for n in ${SCALE_NS:-8 4}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2975$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  echo "N=$n rc=$?"
  python - gpurun_out/scale_$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "e2e %.1f ms" % d["e2e"]["ms_per_step"], {k: round(v, 2) for k, v in d["stages_ms"].items()})
    print("   cg iteration", d.get("cg_iteration_ms"), "planes", d["config"].get("slab_planes"))
    for r in d.get("per_rank", [])[:8]: print("      ", {k: r[k] for k in ("ms_bin", "ms_project", "ms_assemble", "ms_sign", "ms_cc", "ms_cg", "ms_threshold", "planes")})
except Exception as e:
    print("ERR", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_$n.err | tail -3
done
timeout 200 python -m pytest tests/test_multi.py -m gpu -q -k "simp24" 2>&1 | tail -3
