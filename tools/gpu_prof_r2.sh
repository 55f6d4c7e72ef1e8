#!/bin/bash
# ncu evidence for the default build of the next round at the target size: launch list of one bench step and a --set full capture of the
# dominant kernel (HexBox projection), then profiles/project_hex8_traffic.json (bench.py quotes it as roofline.traffic).
mkdir -p gpurun_out
TAG=${1:-r2a}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain256.log 2>&1 || { echo plain failed; tail gpurun_out/plain256.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}_n256.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_project_hex8" -s 1 -c 1 -o gpurun_out/prof_${TAG}_project_n256 -f $CMD > gpurun_out/ncu_a.log 2>&1; echo "ncu project rc=$?"
python tools/extract_traffic.py gpurun_out/prof_${TAG}_project_n256.ncu-rep k_project_hex8 256 && cp profiles/project_hex8_traffic.json gpurun_out/
python tools/launch_summary.py gpurun_out/launches_${TAG}_n256.csv > gpurun_out/${TAG}_launch_summary_n256.txt 2>&1; head -20 gpurun_out/${TAG}_launch_summary_n256.txt
