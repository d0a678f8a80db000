#!/bin/bash
# A/B of library variants on the unconfined legs (C3 and C4 fields, confined=False):  tools/ab_unc.sh <out name> <variant> [<variant> ...]
set -u
mkdir -p gpurun_out
name=$1; shift
out=gpurun_out/$name.txt; : > $out
one() {
  label=$1; lib=$2; shift 2
  r=$(ONEKA_B200_LIB=$lib timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --legs none --unconfined "$@" 2>>gpurun_out/${name}_err.log | tail -1 |
      python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d.get('parity') or {}; print('%.2f ms/step  %.4g attempts/s | parity cells %s steps_equal %s err %.1e' % (d['ms_per_step'], d['value'], p.get('differing_cells'), p.get('step_counts_equal'), p.get('endpoint_max_rel_err', -1)))" 2>&1)
  echo "$label | $* | $r" >> $out
}
for v in "$@" "$1"; do
  lib=$PWD/build/lib_$v.so
  one $v $lib --workload c3 --realizations 4000
  one $v $lib --workload c4 --realizations 1024
done
cat $out
