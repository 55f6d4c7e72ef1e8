#!/bin/bash
mkdir -p gpurun_out
NG=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29741 tests/slab_parity_ranks.py 48 > gpurun_out/slab_parity_$NG.log 2>&1; echo "slab parity ($NG ranks) rc=$?"
grep -E "FAIL|OK|flipped|Error|error" gpurun_out/slab_parity_$NG.log | head -12
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
