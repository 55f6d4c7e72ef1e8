#!/bin/bash
# ncu evidence for the default build at the target size (BASELINE configs[4], one B200): the launch list of one bench run and --set full
# captures of the heavy kernels (second pipeline pass, so that buffers are allocated and caches warm like in the timed region).
# usage: tools/gpu_prof_r2.sh TAG [kernel-group ...]   groups: proj scan asm sign cg upd vol fine   (default: all)
mkdir -p gpurun_out
TAG=${1:-r2a}; shift
GROUPS_=${@:-proj asm sign cg upd vol}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain256.log 2>&1 || { echo plain failed; tail gpurun_out/plain256.log; exit 1; }
if [ -z "$NOLIST" ]; then      # NOLIST=1: captures only
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}_n256.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/launches_${TAG}_n256.csv > gpurun_out/${TAG}_launch_summary_n256.txt 2>&1; head -40 gpurun_out/${TAG}_launch_summary_n256.txt
fi
cap() {   # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/prof_${TAG}_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
for g in $GROUPS_; do
  case $g in
    proj) cap proj "k_project_list" 2 2 ;;
    scan) cap scan "k_pair_scan|k_box_records" 3 3 ;;
    asm)  cap asm "k_assemble|k_faces_crossing" 2 2 ;;
    sign) cap sign "k_sign_lattice|k_lat_info" 3 3 ;;
    cg)   cap cg "k_stencil81_tma" 24 1 ;;
    upd)  cap upd "k_cg_update" 23 1 ;;
    vol)  cap vol "k_vol_cut|k_vol_rows|k_vl_step|k_vl_eval" 90 4 ;;
    fine) cap fine "k_fine_eval2" 1 1 ;;
  esac
done
ls -la gpurun_out/*${TAG}*.ncu-rep
