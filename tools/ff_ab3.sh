#!/bin/bash
# third A/B: tile lookup (1.5*2^52 trick vs F2I/I2F), software-pipelined Horner; all-FP64 polynomial
out=gpurun_out/ff_ab3.txt; : > $out
run() {
  label=$1; shift
  line=$(env "$@" timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e $BARGS 2>>gpurun_out/ff_ab3_err.log | tail -1)
  echo "$label $BARGS :: $(echo "$line" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.2f attempts/s %.4g real/s %.0f' % (d['ms_per_step'], d['value'], d['realizations_per_s']))" 2>&1)" >> $out
}
for w in "--workload c3 --realizations 4000" "--workload c4 --realizations 1024"; do
  BARGS="$w"
  run magic X=1
  run cvt ONEKA_B200_LIB=$PWD/build/lib_ff_cvt.so
  run cvt_pf ONEKA_B200_LIB=$PWD/build/lib_ff_cvt_pf.so
  run magic_pf ONEKA_B200_LIB=$PWD/build/lib_ff_pf.so
done
cat $out
