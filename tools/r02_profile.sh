#!/bin/bash
# round 2 profiles (one GPU): launch list of the bench command (kernel shares), ncu --set full of the fused far-field kernel (C3) and of
# the direct kernel on the raster-heavy C5, with source-level pages.  Numbers printed by a run under ncu are never bench values.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --realizations 2000 --no-cpu --no-e2e --legs none > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:track_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_r02_c3 \
    python bench.py --realizations 1000 --steps 1 --warmup 3 --no-cpu --no-e2e --legs none > gpurun_out/prof_r02_c3.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:track_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_r02_c4 \
    python bench.py --workload c4 --realizations 512 --steps 1 --warmup 3 --no-cpu --no-e2e --legs none > gpurun_out/prof_r02_c4.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:track_kernel --launch-skip 2 -c 1 -f -o gpurun_out/prof_r02_c5 \
    python bench.py --workload c5 --realizations 512 --steps 1 --warmup 3 --no-cpu --no-e2e --legs none > gpurun_out/prof_r02_c5.log 2>&1
ncu --set full --clock-control none --kernel-name regex:"farfield_coef_kernel|flush_kernel" --launch-skip 4 -c 2 -f -o gpurun_out/prof_r02_small \
    python bench.py --realizations 2000 --steps 1 --warmup 3 --no-cpu --no-e2e --legs none > gpurun_out/prof_r02_small.log 2>&1
ls -la gpurun_out/*.ncu-rep
