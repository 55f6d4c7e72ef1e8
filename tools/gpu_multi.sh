#!/bin/bash
# multi-GPU call (gpurun --gpus N): NCCL slab parity (both exchange paths), the in-process r2s_multi tests on distinct devices, scaling lines
mkdir -p gpurun_out
NG=${1:-2}
timeout 900 python -m pytest tests/test_slabs.py tests/test_multi.py -m gpu -q -rf > gpurun_out/pytest_multi_$NG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_multi_$NG.log
for n in 1 2 4 8; do
  if [ $n -gt $NG ]; then break; fi
  if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2975$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  echo "N=$n rc=$?"
  python - gpurun_out/scale_$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "e2e %.1f ms" % d["e2e"]["ms_per_step"], {k: round(v, 2) for k, v in d["stages_ms"].items()})
    print("   cg iteration", d.get("cg_iteration_ms"), "planes", d["config"].get("slab_planes"))
except Exception as e:
    print("ERR", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/scale_$n.err | tail -3
done
