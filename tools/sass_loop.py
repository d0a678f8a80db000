#!/usr/bin/env python
"""Dump the SASS of one kernel of a built library and count the instructions of its hottest loops
(backward branches), by class.  usage: sass_loop.py lib.so 'track_kernelILb1ELi0E' [--dump]"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = next(f for f in funcs if pat in f.split("\n", 1)[0])
ins = []
for l in body.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((int(m[1], 16), m[2].strip()))
if "--dump" in sys.argv:
    for a, t in ins: print("%05x  %s" % (a, t))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m[1], 16) <= a and int(m[1], 16) in addr:
        loops.append((addr[int(m[1], 16)], i))
def cls(t):
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    if op.startswith(("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")): return "fp64"
    if op.startswith("MUFU"): return "mufu"
    if op.startswith("LDCU"): return "ctl/uniform"
    if op.startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL", "RED", "ATOM", "LDC")): return "mem"
    if op.startswith(("U", "S2UR", "R2UR", "BRA", "BSSY", "BSYNC")): return "ctl/uniform"
    if op.startswith(("MOV", "IMAD.MOV")): return "mov"
    return "other"
print("kernel has %d instructions, %d MUFU.RCP64H" % (len(ins), sum("RCP64H" in t for _, t in ins)))
for s, e in loops:
    n = e - s + 1
    c = {}
    for _, t in ins[s:e + 1]: c[cls(t)] = c.get(cls(t), 0) + 1
    nr = sum("RCP64H" in t for _, t in ins[s:e + 1])
    if nr >= 2: print("loop %05x..%05x: %3d instr, %d wells -> %.2f instr/well  %s" % (ins[s][0], ins[e][0], n, nr, n / nr, c))
