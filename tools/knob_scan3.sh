#!/bin/bash
# HISTORICAL: how profiles/r02_knob_scan*.txt were produced.  The -D macros of the variants (ONEKA_TRACK_THREADS, FUSED_MIN_CTAS,
# ONEKA_FF_ORDER_FIXED, ONEKA_FF_PREFETCH, ONEKA_FF_NEAR_TAIL ...) existed only while the scan ran; the winners are now the code.
# round 2, third scan: (eta, order) pairs at equal truncation with as many tiles as 3 CTAs x 256 threads leave room for; register budgets
set -u
mkdir -p gpurun_out
out=gpurun_out/knob_scan3.txt; : > $out
line() { # label, env assignments...
  label=$1; shift
  for w in "--workload c3 --realizations 4000" "--workload c4 --realizations 1024"; do
    r=$(env "$@" timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --legs none $w 2>>gpurun_out/knob_scan3_err.log | tail -1 |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['config'].get('farfield') or {}; p=d.get('parity') or {}; print('%.2f ms/step  %.4g attempts/s  tiles %sx%s order %s eta %s near %.2f | parity cells %s steps_equal %s err %.1e' % (d['ms_per_step'], d['value'], f.get('ntx'), f.get('nty'), f.get('order'), f.get('eta'), f.get('mean_near', 0), p.get('differing_cells'), p.get('step_counts_equal'), p.get('endpoint_max_rel_err', -1)))" 2>&1)
    echo "$label | $w | $r" >> $out
  done
}
B=$PWD/build
T=$B/lib_t256.so
line t256_e15_o16_t246 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=246 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line t256_e20_o20_t202 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=202 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
line t256_e20_o18_t220 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=220 ONEKA_FARFIELD_ORDER=18 ONEKA_FARFIELD_ETA=0.2
line t256_e25_o22_t185 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=185 ONEKA_FARFIELD_ORDER=22 ONEKA_FARFIELD_ETA=0.25
line t256_e25_o20_t202 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=202 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.25
line t256_e30_o26_t159 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=159 ONEKA_FARFIELD_ORDER=26 ONEKA_FARFIELD_ETA=0.3
line t256_e30_o24_t170 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=170 ONEKA_FARFIELD_ORDER=24 ONEKA_FARFIELD_ETA=0.3
line t256_e175_o18_t220 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=220 ONEKA_FARFIELD_ORDER=18 ONEKA_FARFIELD_ETA=0.175
line c5_e15_o16_t140 ONEKA_B200_LIB=$B/lib_c5.so ONEKA_FARFIELD_TILES=140 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line c5_e20_o20_t115 ONEKA_B200_LIB=$B/lib_c5.so ONEKA_FARFIELD_TILES=115 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
line c5pf_e15_o16_t140 ONEKA_B200_LIB=$B/lib_c5pf.so ONEKA_FARFIELD_TILES=140 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line c4_e15_o16_t180 ONEKA_B200_LIB=$B/lib_c4.so ONEKA_FARFIELD_TILES=180 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line t256c2_e15_o16_t380 ONEKA_B200_LIB=$B/lib_t256c2.so ONEKA_FARFIELD_TILES=380 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line t256c2_e20_o20_t300 ONEKA_B200_LIB=$B/lib_t256c2.so ONEKA_FARFIELD_TILES=300 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
line t256c2pf_e15_o16_t380 ONEKA_B200_LIB=$B/lib_t256c2pf.so ONEKA_FARFIELD_TILES=380 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line t256c2pf_e20_o20_t300 ONEKA_B200_LIB=$B/lib_t256c2pf.so ONEKA_FARFIELD_TILES=300 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
cat $out
