"""Per-source-line warp instructions, lane utilisation and stall samples of one kernel (oneka_device.cuh only).
    ncu -i prof.ncu-rep --page source --csv > sass.csv ; cuobjdump -xelf all lib.so ; nvdisasm -g x.cubin > dis.txt
    python tools/ncu_lines.py sass.csv dis.txt <mangled kernel prefix> [min share %]"""
import csv, importlib.util, os, sys
from collections import defaultdict
here = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("nbl", os.path.join(here, "ncu_by_line.py"))
nbl = importlib.util.module_from_spec(spec); spec.loader.exec_module(nbl)
sass_csv, dis, kernel = sys.argv[1:4]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
amap = nbl.line_map(dis, kernel)
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); ix = {n: i for i, n in enumerate(rows[h])}
src = open(os.path.join(here, "..", "onekapy_b200", "csrc", "oneka_device.cuh")).read().split("\n")
stalls = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_math", "stall_lg", "stall_no_inst"]
per = defaultdict(lambda: defaultdict(int)); base = None
for r in rows[h + 1:]:
    try: addr = int(r[0], 16)
    except ValueError: continue
    base = addr if base is None else base
    key = amap.get(addr - base)
    ln = key[1] if key and key[0] == "oneka_device.cuh" else 0
    per[ln]["inst"] += int(r[ix["Instructions Executed"]] or 0); per[ln]["thr"] += int(r[ix["Thread Instructions Executed"]] or 0)
    per[ln]["samp"] += int(r[ix["# Samples"]] or 0); per[ln]["n"] += 1
    for s in stalls: per[ln][s] += int(r[ix[s]] or 0)
ti = sum(v["inst"] for v in per.values()); ts = sum(v["samp"] for v in per.values())
print("total warp inst %.4g, avg lanes %.1f, samples %d" % (ti, sum(v["thr"] for v in per.values()) / ti, ts))
print(" line sass inst%  lanes samp% | " + " ".join(s.replace("stall_", "")[:7].rjust(7) for s in stalls))
for ln in sorted(per):
    v = per[ln]
    if v["inst"] >= thr / 100 * ti or v["samp"] >= thr / 100 * ts:
        print("%5d %4d %5.2f %5.1f %5.2f | " % (ln, v["n"], 100 * v["inst"] / ti, v["thr"] / max(v["inst"], 1), 100 * v["samp"] / ts)
              + " ".join("%7.2f" % (100 * v[s] / ts) for s in stalls) + " | " + (src[ln - 1].strip()[:100] if ln else "(other files)"))
