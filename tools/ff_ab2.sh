#!/bin/bash
# second A/B: FP32 tail build (default lib) with 64 / 80 tiles, FP64 orders, and the noinline-evaluation build
out=gpurun_out/ff_ab2.txt; : > $out
run() {
  label=$1; shift
  line=$(env "$@" timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e $BARGS 2>>gpurun_out/ff_ab2_err.log | tail -1)
  echo "$label $BARGS :: $(echo "$line" | python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['config'].get('farfield') or {}; print('ms/step %.2f attempts/s %.4g real/s %.0f frac %.3f ff=%s' % (d['ms_per_step'], d['value'], d['realizations_per_s'], d['roofline']['frac'], {k: f.get(k) for k in ('ntx','nty','order','eta','mean_near')} if f else None))" 2>&1)" >> $out
}
for w in "--workload c3 --realizations 4000" "--workload c4 --realizations 1024"; do
  BARGS="$w"
  run tail_64tiles ONEKA_FARFIELD=auto
  run tail_80tiles ONEKA_FARFIELD=auto ONEKA_FARFIELD_TILES=80
  run tail_fp64_10 ONEKA_FARFIELD=auto ONEKA_FARFIELD_FP64=10
  run noinline_64 ONEKA_FARFIELD=auto ONEKA_B200_LIB=$PWD/build/lib_ff_noinline.so
done
cat $out
