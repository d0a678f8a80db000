"""Aggregate an ncu source-page CSV (SASS view) by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<k> [--launch-skip n --launch-count 1] > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g x.cubin > dis.txt
    python tools/ncu_by_line.py sass.csv dis.txt <mangled kernel name> [top]

Maps each SASS address to the `//## File "...", line N` marker preceding it in the nvdisasm
listing of the same kernel and sums warp-level instruction counts and stall samples per line.
"""
import csv
import re
import sys
from collections import defaultdict


def line_map(dis_path, kernel):
    amap = {}
    cur = None
    inside = False
    for ln in open(dis_path, errors="replace"):
        if ln.startswith(".text.") or ln.strip().startswith(".section"):
            inside = kernel in ln and ".text." in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            inl = " (inlined)" if "inlined at" in ln else ""
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m and cur:
            amap[int(m.group(1), 16)] = cur
    return amap


def main():
    sass_csv, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    amap = line_map(dis, kernel)
    rows = list(csv.reader(open(sass_csv)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    agg = defaultdict(lambda: [0, 0, 0])
    base = None
    for r in rows[h + 1:]:
        try:
            addr = int(r[0], 16) if not r[0].isdigit() else int(r[0])
        except ValueError:
            continue
        if base is None:
            base = addr
        key = amap.get(addr - base, ("?", 0))
        a = agg[key]
        a[0] += int(r[ci] or 0)
        a[1] += int(r[ti] or 0)
        a[2] += int(r[si] or 0)
    tot = sum(a[0] for a in agg.values())
    tots = sum(a[2] for a in agg.values())
    print("total warp instructions %d, samples %d, mapped lines %d" % (tot, tots, len(agg)))
    src = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        f, n = key
        if f not in src:
            try:
                src[f] = open("/root/repo/onekapy_b200/csrc/" + f).read().split("\n")
            except OSError:
                src[f] = []
        text = src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ""
        print("%5.2f%% inst  %5.2f%% samp  thr/inst %4.1f  %s:%d  %s" % (100.0 * a[0] / tot, 100.0 * a[2] / max(1, tots), a[1] / max(1, a[0]), f, n, text))


if __name__ == "__main__":
    main()
