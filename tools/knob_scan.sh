#!/bin/bash
# A/B scan of the prepared build knobs and far-field settings (run on the GPU box, e.g. under gpurun; ~15 s per bench line).
#   tools/knob_scan.sh            builds the variants here (nvcc cross-compiles without a GPU) if build/lib_*.so are missing
# Every variant first has to pass the far-field parity tests; then C3 / C4 are timed in short runs.  Output: gpurun_out/knob_scan.txt
set -u
mkdir -p build gpurun_out
B="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -I include onekapy_b200/csrc/oneka_api.cu"
declare -A VAR=( [rkloop]="-DONEKA_RK_LOOP=1" [global]="-DONEKA_FF_COEF_GLOBAL=1" [rkloop_global]="-DONEKA_RK_LOOP=1 -DONEKA_FF_COEF_GLOBAL=1" )
for v in "${!VAR[@]}"; do [ -f build/lib_$v.so ] || $B ${VAR[$v]} -o build/lib_$v.so; done
out=gpurun_out/knob_scan.txt; : > $out
line() { # label, env assignments...
  label=$1; shift
  for w in "--workload c3 --realizations 4000" "--workload c4 --realizations 1024"; do
    r=$(env "$@" timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e $w 2>>gpurun_out/knob_scan_err.log | tail -1 |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['config'].get('farfield') or {}; print('%.2f ms/step  %.4g attempts/s  tiles %sx%s order %s eta %s near %.2f' % (d['ms_per_step'], d['value'], f.get('ntx'), f.get('nty'), f.get('order'), f.get('eta'), f.get('mean_near', 0)))" 2>&1)
    echo "$label | $w | $r" >> $out
  done
}
parity() { env "$@" timeout 200 python -m pytest tests/test_gpu_farfield.py tests/test_gpu_parity.py -x -q -k "farfield or fused or traces_vs" 2>&1 | tail -1; }
line default X=1
line default_o24 ONEKA_FARFIELD_ORDER=24 ONEKA_FARFIELD_ETA=0.25
for v in rkloop global rkloop_global; do
  L=$PWD/build/lib_$v.so
  echo "$v parity: $(parity ONEKA_B200_LIB=$L)" >> $out
  line $v ONEKA_B200_LIB=$L
  line ${v}_o24 ONEKA_B200_LIB=$L ONEKA_FARFIELD_ORDER=24 ONEKA_FARFIELD_ETA=0.25
done
for t in 128 256; do    # more tiles only make sense when the coefficients do not live in shared memory
  for v in global rkloop_global; do
    line ${v}_t${t}_o24 ONEKA_B200_LIB=$PWD/build/lib_$v.so ONEKA_FARFIELD_TILES=$t ONEKA_FARFIELD_ORDER=24 ONEKA_FARFIELD_ETA=0.25
  done
done
cat $out
