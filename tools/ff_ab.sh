#!/bin/bash
# A/B of the far-field compression (direct sums vs tiled expansions) on short bench runs; output in gpurun_out/ff_ab.txt
out=gpurun_out/ff_ab.txt; : > $out
run() { # label, env..., -- bench args
  label=$1; shift
  line=$(env "$@" timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e $BARGS 2>>gpurun_out/ff_ab_err.log | tail -1)
  echo "$label $BARGS :: $(echo "$line" | python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['config'].get('farfield') or {}; print('ms/step %.2f attempts/s %.4g real/s %.0f frac %.3f ff=%s' % (d['ms_per_step'], d['value'], d['realizations_per_s'], d['roofline']['frac'], {k: f.get(k) for k in ('ntx','nty','order','eta','mean_near')} if f else None))" 2>&1)" >> $out
}
for w in "--workload c3 --realizations 4000" "--workload c4 --realizations 1024"; do
  BARGS="$w"
  run direct ONEKA_FARFIELD=off
  run ff_28_0.30 ONEKA_FARFIELD=auto
  run ff_24_0.25 ONEKA_FARFIELD=auto ONEKA_FARFIELD_ORDER=24 ONEKA_FARFIELD_ETA=0.25
done
cat $out
