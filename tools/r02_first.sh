#!/bin/bash
# round 2, first GPU call: the knob scan prepared in profiles/r02_plan.md + the unconfined far field on hardware
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
bash tools/knob_scan.sh > /dev/null 2>&1
ONEKA_TEST_UNCONFINED_FF=1 timeout 300 python -m pytest tests/test_gpu_farfield.py -x -q -k unconfined -s 2>&1 | tail -15 > gpurun_out/r02_unc_ff_test.txt
for ff in 0 1; do
  for w in "c3 4000" "c4 1024"; do set -- $w
    ONEKA_FARFIELD_UNCONFINED=$ff timeout 200 python bench.py --workload $1 --realizations $2 --unconfined --steps 3 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/r02_unc_err.log | tail -1 |
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unconfined ff=$ff $1: %.2f ms/step %.4g attempts/s farfield=%s' % (d['ms_per_step'], d['value'], d['config'].get('farfield')))" >> gpurun_out/r02_unc_ff_bench.txt 2>&1
  done
done
cat gpurun_out/knob_scan.txt gpurun_out/r02_unc_ff_test.txt gpurun_out/r02_unc_ff_bench.txt
