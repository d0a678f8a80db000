#!/bin/bash
# HISTORICAL: how profiles/r02_knob_scan*.txt were produced.  The -D macros of the variants (ONEKA_TRACK_THREADS, FUSED_MIN_CTAS,
# ONEKA_FF_ORDER_FIXED, ONEKA_FF_PREFETCH, ONEKA_FF_NEAR_TAIL ...) existed only while the scan ran; the winners are now the code.
# round 2, second scan: CTA size x tile count x order x eta of the far field (and the Horner loop unrolled at a fixed order).
# The variants are built here by tools/build_variant.sh; every line is a short bench run (C3 R=4000, C4 R=1024).
set -u
mkdir -p gpurun_out
out=gpurun_out/knob_scan2.txt; : > $out
line() { # label, env assignments...
  label=$1; shift
  for w in "--workload c3 --realizations 4000" "--workload c4 --realizations 1024"; do
    r=$(env "$@" timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --legs none $w 2>>gpurun_out/knob_scan2_err.log | tail -1 |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['config'].get('farfield') or {}; p=d.get('parity') or {}; print('%.2f ms/step  %.4g attempts/s  tiles %sx%s order %s eta %s near %.2f | parity cells %s steps_equal %s err %.1e' % (d['ms_per_step'], d['value'], f.get('ntx'), f.get('nty'), f.get('order'), f.get('eta'), f.get('mean_near', 0), p.get('differing_cells'), p.get('step_counts_equal'), p.get('endpoint_max_rel_err', -1)))" 2>&1)
    echo "$label | $w | $r" >> $out
  done
}
B=$PWD/build
line default X=1
line c5regs ONEKA_B200_LIB=$B/lib_c5.so
line o28_unrolled ONEKA_B200_LIB=$B/lib_o28.so
line o20_unrolled_t96_e20 ONEKA_B200_LIB=$B/lib_o20.so ONEKA_FARFIELD_TILES=96 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
line t256 ONEKA_B200_LIB=$B/lib_t256.so
line t256_t128_o24_e25 ONEKA_B200_LIB=$B/lib_t256.so ONEKA_FARFIELD_TILES=128 ONEKA_FARFIELD_ORDER=24 ONEKA_FARFIELD_ETA=0.25
line t256_t150_o24_e25 ONEKA_B200_LIB=$B/lib_t256.so ONEKA_FARFIELD_TILES=150 ONEKA_FARFIELD_ORDER=24 ONEKA_FARFIELD_ETA=0.25
line t256_t160_o22_e20 ONEKA_B200_LIB=$B/lib_t256.so ONEKA_FARFIELD_TILES=160 ONEKA_FARFIELD_ORDER=22 ONEKA_FARFIELD_ETA=0.2
line t256_t190_o20_e20 ONEKA_B200_LIB=$B/lib_t256.so ONEKA_FARFIELD_TILES=190 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
line t256_t200_o18_e15 ONEKA_B200_LIB=$B/lib_t256.so ONEKA_FARFIELD_TILES=200 ONEKA_FARFIELD_ORDER=18 ONEKA_FARFIELD_ETA=0.15
line t256_t230_o16_e15 ONEKA_B200_LIB=$B/lib_t256.so ONEKA_FARFIELD_TILES=230 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line t256_o20u_t190_e20 ONEKA_B200_LIB=$B/lib_t256_o20.so ONEKA_FARFIELD_TILES=190 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
line t256_o16u_t230_e15 ONEKA_B200_LIB=$B/lib_t256_o16.so ONEKA_FARFIELD_TILES=230 ONEKA_FARFIELD_ORDER=16 ONEKA_FARFIELD_ETA=0.15
line t256_o20u_t120_e20 ONEKA_B200_LIB=$B/lib_t256_o20.so ONEKA_FARFIELD_TILES=120 ONEKA_FARFIELD_ORDER=20 ONEKA_FARFIELD_ETA=0.2
cat $out
