#!/bin/bash
# Where do the rasteriser flavours pay?  tools/flavour_scan.sh <out name> "<modes>" <workload> <realizations> <umbra> [<umbra> ...]
# (spacing 4: window rows = 2 umbra / 4 + 1; ONEKA_RASTER_MODE forces the flavour).  Output: gpurun_out/<out name>.txt
set -u
mkdir -p gpurun_out
name=$1; modes=$2; wl=$3; R=$4; shift 4
out=gpurun_out/$name.txt
for u in "$@"; do
  for m in $modes; do
    r=$(ONEKA_RASTER_MODE=$m timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --legs none --workload $wl --realizations $R --umbra $u 2>>gpurun_out/${name}_err.log | tail -1 |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d.get('parity') or {}; print('%.2f ms/step  %.4g attempts/s | parity cells %s' % (d['ms_per_step'], d['value'], p.get('differing_cells')))" 2>&1)
    echo "$wl umbra $u (rows $((2*u/4+1))) | $m | $r" >> $out
  done
done
