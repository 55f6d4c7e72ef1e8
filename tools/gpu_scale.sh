#!/bin/bash
# strong-scaling sweep on one box: N = 1, 2, 4, 8 (as the driver does), plus the slab parity check on NG ranks
mkdir -p gpurun_out
NG=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29741 tests/slab_parity_ranks.py 48 > gpurun_out/slab_parity_$NG.log 2>&1; echo "slab parity ($NG ranks) rc=$?"
grep -E "FAIL|OK|flipped|Error" gpurun_out/slab_parity_$NG.log | head -16
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print("   ", {k: round(v, 2) for k, v in d["stages_ms"].items()})
    print("    planes", d["config"].get("slab_planes"), "cg iteration", d.get("cg_iteration_ms"))
    for r in d.get("per_rank", [])[:8]: print("      ", {k: r[k] for k in ("ms_project", "ms_assemble", "ms_sign", "ms_cc", "ms_cg", "cg_matvec", "cg_xchg1", "cg_update", "cg_xchg2", "planes")})
except Exception as e:
    print(f, "ERR", e)
PY
}
for n in 1 2 4 8; do
  if [ $n -gt $NG ]; then break; fi
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2975$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "N=$n rc=$?"; show gpurun_out/scale_$n.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_$n.err | tail -3
done
