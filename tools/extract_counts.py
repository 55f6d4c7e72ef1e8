"""Executed FP64 work and DRAM traffic of the projection kernel from an `ncu --set full` report -> profiles/project_list_counts.json.

bench.py quotes `roofline.achieved` for k_project_list from these MEASURED counts (thread-level dfma / dadd / dmul instructions with the
predicate on; a dfma counts 2 flop) instead of an assumed flop count per pair, and `roofline.traffic` from the measured DRAM bytes.
Usage: extract_counts.py report.ncu-rep [pairs_solved_per_step]"""
import csv, json, os, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]; col = {h: i for i, h in enumerate(hdr)}
def num(r, k): return float(r[col[k]].replace(",", ""))
def to_bytes(r, k):
    return num(r, k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[units[col[k]].lower()]
def to_ms(r, k):
    return num(r, k) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[col[k]]]
out = {"kernel": "k_project_list", "launches": [], "source": os.path.basename(rep)}
for r in rows[2:]:
    if len(r) < len(hdr) or "k_project_list" not in r[col["Kernel Name"]]:
        continue
    cyc = num(r, "smsp__cycles_elapsed.avg")
    ops = {o: num(r, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % o) * cyc for o in ("dfma", "dadd", "dmul")}
    out["launches"].append({"ms_under_ncu": to_ms(r, "gpu__time_duration.sum"), "thread_dfma": ops["dfma"], "thread_dadd": ops["dadd"], "thread_dmul": ops["dmul"],
                            "fp64_flop": 2 * ops["dfma"] + ops["dadd"] + ops["dmul"], "dram_bytes": to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum"),
                            "lanes_per_warp_inst": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                            "fp64_pipe_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"), "warp_inst": num(r, "smsp__inst_executed.sum"),
                            "registers": num(r, "launch__registers_per_thread")})
if not out["launches"]:
    sys.exit("k_project_list not found in the report")
out["fp64_flop_per_step"] = sum(l["fp64_flop"] for l in out["launches"])
out["dram_bytes_per_step"] = sum(l["dram_bytes"] for l in out["launches"])
out["ms_per_step_under_ncu"] = sum(l["ms_under_ncu"] for l in out["launches"])
if len(sys.argv) > 2:
    out["pairs_solved_per_step"] = int(sys.argv[2])
    out["fp64_flop_per_solved_pair"] = out["fp64_flop_per_step"] / out["pairs_solved_per_step"]
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "project_list_counts.json")
json.dump(out, open(p, "w"), indent=1); print(json.dumps(out, indent=1))
