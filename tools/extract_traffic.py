"""Write profiles/project_hex8_traffic.json (DRAM bytes per launch of the dominant kernel) from an ncu --set full report.
Usage: extract_traffic.py report.ncu-rep kernel_regex n_elements_per_axis"""
import csv, json, os, re, subprocess, sys
rep, pat, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; units = rows[1]; col = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v.replace(",", "")); u = u.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
out = None
for r in rows[2:]:
    if len(r) < len(hdr) or not re.search(pat, r[col["Kernel Name"]]): continue
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]); wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    kname = r[col["Kernel Name"]].split("(")[0]
    targs = re.search(r"k_project_hex8<([^>]*)>", kname)       # <XP, MINB, SMEM_A, BOX, MODE, P1>: the 4th argument tells the element variant
    variant = "box" if targs and len(targs.group(1).split(",")) >= 4 and targs.group(1).split(",")[3].strip() in ("1", "true") else "general"
    out = {"kernel": kname, "variant": variant, "n": n, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "duration_ms_under_ncu": float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[col["gpu__time_duration.sum"]]],
           "source": os.path.basename(rep)}
    break
if out is None: sys.exit("kernel not found")
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "project_hex8_traffic.json")
json.dump(out, open(p, "w"), indent=1); print(out)
