"""Key metrics per profiled kernel from an .ncu-rep (ncu -i ... --page raw --csv). Usage: ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("sass__inst_executed_register_spilling", "spill_inst"), ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%")]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("== %s  [%s]" % (r[col["Kernel Name"]][:90], r[col["ID"]]))
    out = []
    for k, n in want:
        if k in col:
            out.append("%s=%s%s" % (n, r[col[k]], rows[1][col[k]] if n in ("time", "dram_rd", "dram_wr") else ""))
    print("   " + "  ".join(out))
    st = sorted(((float(r[col[h]] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls), reverse=True)[:6]
    print("   stalls/issue: " + "  ".join("%s=%.2f" % (n, v) for v, n in st))
