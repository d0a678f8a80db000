#!/bin/bash
# HISTORICAL: how profiles/r02_knob_scan*.txt were produced.  The -D macros of the variants (ONEKA_TRACK_THREADS, FUSED_MIN_CTAS,
# ONEKA_FF_ORDER_FIXED, ONEKA_FF_PREFETCH, ONEKA_FF_NEAR_TAIL ...) existed only while the scan ran; the winners are now the code.
# round 2, fifth scan: CTA shape / register budget on the kernels WITHOUT the far field (direct sums, unconfined, raster-heavy C5, tiny C1)
set -u
mkdir -p gpurun_out
out=gpurun_out/knob_scan5.txt; : > $out
line() { # label, env assignments...
  label=$1; shift
  for w in "--workload c5" "--workload c1" "--workload c3 --realizations 4000 --farfield off" "--workload c3 --realizations 4000 --unconfined" "--workload c4 --realizations 512 --farfield off"; do
    r=$(env "$@" timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --legs none $w 2>>gpurun_out/knob_scan5_err.log | tail -1 |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['config'].get('farfield') or {}; p=d.get('parity') or {}; print('%.2f ms/step  %.4g attempts/s  ff %s | parity cells %s steps_equal %s err %.1e' % (d['ms_per_step'], d['value'], f.get('order'), p.get('differing_cells'), p.get('step_counts_equal'), p.get('endpoint_max_rel_err', -1)))" 2>&1)
    echo "$label | $w | $r" >> $out
  done
}
B=$PWD/build
line default X=1
line t256c2 ONEKA_B200_LIB=$B/lib_t256c2.so
line t256 ONEKA_B200_LIB=$B/lib_t256.so
line c5regs ONEKA_B200_LIB=$B/lib_c5.so
line c4regs ONEKA_B200_LIB=$B/lib_c4.so
cat $out
