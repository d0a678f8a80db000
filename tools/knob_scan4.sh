#!/bin/bash
# HISTORICAL: how profiles/r02_knob_scan*.txt were produced.  The -D macros of the variants (ONEKA_TRACK_THREADS, FUSED_MIN_CTAS,
# ONEKA_FF_ORDER_FIXED, ONEKA_FF_PREFETCH, ONEKA_FF_NEAR_TAIL ...) existed only while the scan ran; the winners are now the code.
# round 2, fourth scan: 256 threads x 2 CTAs per SM (128 registers, 113 KB of shared memory per CTA): tiles / order / eta; + one ncu capture
set -u
mkdir -p gpurun_out
out=gpurun_out/knob_scan4.txt; : > $out
line() { # label, env assignments...
  label=$1; shift
  for w in "--workload c3 --realizations 4000" "--workload c4 --realizations 1024"; do
    r=$(env "$@" timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --legs none $w 2>>gpurun_out/knob_scan4_err.log | tail -1 |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['config'].get('farfield') or {}; p=d.get('parity') or {}; print('%.2f ms/step  %.4g attempts/s  tiles %sx%s order %s eta %s near %.2f | parity cells %s steps_equal %s err %.1e' % (d['ms_per_step'], d['value'], f.get('ntx'), f.get('nty'), f.get('order'), f.get('eta'), f.get('mean_near', 0), p.get('differing_cells'), p.get('step_counts_equal'), p.get('endpoint_max_rel_err', -1)))" 2>&1)
    echo "$label | $w | $r" >> $out
  done
}
B=$PWD/build
T=$B/lib_t256c2.so
line c2_e115_o14_t430 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=430 ONEKA_FARFIELD_ORDER=14 ONEKA_FARFIELD_ETA=0.115
line c2_e15_o16_t330 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=330 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line c2_e15_o16_t400 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=400 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line c2_e25_o22_t290 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=290 ONEKA_FARFIELD_ORDER=22 ONEKA_FARFIELD_ETA=0.25
line c2_e20_o20_t320 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=320 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
line c2_e17_o18_t350 ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=350 ONEKA_FARFIELD_ORDER=18 ONEKA_FARFIELD_ETA=0.17
line c2o16u_e15_t380 ONEKA_B200_LIB=$B/lib_t256c2_o16.so ONEKA_FARFIELD_TILES=380 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line c2lds6_e15_o16_t380 ONEKA_B200_LIB=$B/lib_t256c2_lds6.so ONEKA_FARFIELD_TILES=380 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line t512c1_e15_o16_t640 ONEKA_B200_LIB=$B/lib_t512c1.so ONEKA_FARFIELD_TILES=640 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line c2_direct ONEKA_B200_LIB=$T ONEKA_FARFIELD=off
cat $out
ONEKA_B200_LIB=$T ONEKA_FARFIELD_TILES=380 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15 timeout 600 ncu --set full --import-source on --clock-control none \
  --kernel-name regex:track_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_r02_t256c2 python bench.py --realizations 1000 --steps 1 --warmup 3 --no-cpu --no-e2e --legs none > gpurun_out/prof_r02_t256c2.log 2>&1
ls -la gpurun_out/prof_r02_t256c2.ncu-rep
