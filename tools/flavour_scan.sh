#!/bin/bash
# Where does the heavy rasteriser flavour start to pay?  C3 (far-field kernel) and C5's field (direct kernel) at umbra / spacing
# ratios from 2 to 5, both flavours forced (ONEKA_RASTER_MODE).  Output: gpurun_out/flavour_scan.txt
set -u
mkdir -p gpurun_out
out=gpurun_out/flavour_scan.txt; : > $out
one() {
  label=$1; mode=$2; shift 2
  r=$(ONEKA_RASTER_MODE=$mode timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --legs none "$@" 2>>gpurun_out/flavour_scan_err.log | tail -1 |
      python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d.get('parity') or {}; print('%.2f ms/step  %.4g attempts/s | parity cells %s' % (d['ms_per_step'], d['value'], p.get('differing_cells')))" 2>&1)
  echo "$label | $mode | $r" >> $out
}
for u in 8 10 12 14 16 20; do
  for m in plain heavy; do
    one "c3 umbra $u (rows $((2*u/4+1)))" $m --workload c3 --realizations 2000 --umbra $u
  done
done
for u in 8 12 16 20; do
  for m in plain heavy; do
    one "c5 umbra $u (rows $((2*u/4+1)))" $m --workload c5 --realizations 512 --umbra $u
  done
done
cat $out
