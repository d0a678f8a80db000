"""Stall reasons (warp samples) of a kernel folded onto regions of oneka_device.cuh.
    ncu -i prof.ncu-rep --page source --csv > sass.csv ; cuobjdump -xelf all lib.so ; nvdisasm -g x.cubin > dis.txt
    python tools/ncu_stalls_by_region.py sass.csv dis.txt <mangled kernel prefix>"""
import csv, importlib.util, os, sys
here = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("nbl", os.path.join(here, "ncu_by_line.py"))
nbl = importlib.util.module_from_spec(spec); spec.loader.exec_module(nbl)
sass_csv, dis, kernel = sys.argv[1:4]
amap = nbl.line_map(dis, kernel)
src = open(os.path.join(here, "..", "onekapy_b200", "csrc", "oneka_device.cuh")).read().split("\n")
def find(s):
    return next(i + 1 for i, l in enumerate(src) if s in l)
b = {k: find(v) for k, v in dict(poly="void ff_poly_eval", loc="bool ff_locate", ff="int field_feval_ff(", unc="// ---- unconfined flow through the far field",
                                 raster="bool raster_seg(", dkey="unsigned long long dkey", dopri="void dopri_track", stage="void stage_realization",
                                 feval="int field_feval(").items()}
reg = {"scaled_term + rcp": (30, b["feval"] - 1), "direct well loops": (b["feval"], b["poly"] - 40), "horner": (b["poly"], b["loc"] - 1), "tile lookup": (b["loc"], b["ff"] - 10),
       "regional + near wells": (b["ff"] - 9, b["unc"] - 1), "unconfined far field": (b["unc"], b["raster"] - 60), "rasteriser": (b["raster"] - 59, b["dkey"] - 1),
       "dopri (RK algebra, controller)": (b["dopri"] - 3, b["stage"] - 5)}
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {n: i for i, n in enumerate(hdr)}
stalls = ["stall_wait", "stall_not_selected", "stall_selected", "stall_math", "stall_short_sb", "stall_branch_resolving", "stall_dispatch", "stall_no_inst", "stall_long_sb", "stall_lg", "stall_mio"]
agg = {k: dict.fromkeys(stalls + ["inst"], 0) for k in list(reg) + ["other"]}
base = None
for r in rows[h + 1:]:
    try:
        addr = int(r[0], 16) if not r[0].isdigit() else int(r[0])
    except ValueError:
        continue
    base = addr if base is None else base
    key = amap.get(addr - base)
    k = "other"
    if key and key[0] == "oneka_device.cuh":
        k = next((n for n, (lo, hi) in reg.items() if lo <= key[1] <= hi), "other")
    agg[k]["inst"] += int(r[ix["Instructions Executed"]] or 0)
    for s in stalls:
        agg[k][s] += int(r[ix[s]] or 0)
tot = sum(sum(v[s] for s in stalls) for v in agg.values())
ti = sum(v["inst"] for v in agg.values())
print("%-32s %6s %7s | " % ("region", "inst%", "samp%") + " ".join("%8s" % s.replace("stall_", "")[:8] for s in stalls))
for k, v in agg.items():
    if v["inst"]:
        print("%-32s %5.1f%% %6.1f%% | " % (k, 100 * v["inst"] / ti, 100 * sum(v[s] for s in stalls) / tot) + " ".join("%7.1f%%" % (100 * v[s] / tot) for s in stalls))
