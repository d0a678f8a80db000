#!/bin/bash
# HISTORICAL: how profiles/r02_knob_scan*.txt were produced.  The -D macros of the variants (ONEKA_TRACK_THREADS, FUSED_MIN_CTAS,
# ONEKA_FF_ORDER_FIXED, ONEKA_FF_PREFETCH, ONEKA_FF_NEAR_TAIL ...) existed only while the scan ran; the winners are now the code.
# round 2, sixth scan: odd near well on its own (no dummy padding); CTA shape of the unconfined far-field kernel
set -u
mkdir -p gpurun_out
out=gpurun_out/knob_scan6.txt; : > $out
line() { # label, workloads..., env assignments after --
  label=$1; shift
  wl=()
  while [ "$1" != "--" ]; do wl+=("$1"); shift; done; shift
  for w in "${wl[@]}"; do
    r=$(env "$@" timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --legs none $w 2>>gpurun_out/knob_scan6_err.log | tail -1 |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['config'].get('farfield') or {}; p=d.get('parity') or {}; print('%.2f ms/step  %.4g attempts/s  tiles %sx%s order %s near %.2f | parity cells %s steps_equal %s err %.1e' % (d['ms_per_step'], d['value'], f.get('ntx'), f.get('nty'), f.get('order'), f.get('mean_near', 0), p.get('differing_cells'), p.get('step_counts_equal'), p.get('endpoint_max_rel_err', -1)))" 2>&1)
    echo "$label | $w | $r" >> $out
  done
}
B=$PWD/build
C="--workload c3 --realizations 4000"; D="--workload c4 --realizations 1024"
U="--workload c3 --realizations 4000 --unconfined"; V="--workload c4 --realizations 1024 --unconfined"
line base "$C" "$D" "$U" "$V" -- ONEKA_B200_LIB=$B/lib_base.so
line neartail "$C" "$D" -- ONEKA_B200_LIB=$B/lib_neartail.so
line unc128 "$U" "$V" -- ONEKA_B200_LIB=$B/lib_unc128.so
line unc128_t100 "$U" "$V" -- ONEKA_B200_LIB=$B/lib_unc128.so ONEKA_FARFIELD_TILES=90
line unc_off "$U" "$V" -- ONEKA_B200_LIB=$B/lib_base.so ONEKA_FARFIELD_UNCONFINED=0
line base_again "$C" "$D" -- ONEKA_B200_LIB=$B/lib_base.so
cat $out
