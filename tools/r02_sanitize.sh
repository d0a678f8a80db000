#!/bin/bash
# compute-sanitizer over the new kernels of round 2 (256 x 2 far-field shape, coefficient GEMM, fast-row rasteriser, NCCL-free paths) + a fresh
# ncu capture of the final C3 kernel.  One GPU.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/r02_sanitizer_smoke.txt
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_farfield.py tests/test_gpu_round2.py tests/test_gpu_postprocess.py -m gpu -q -x \
    -k "not atomic_probes and not own_communicator" 2>&1 | tail -6 > gpurun_out/r02_sanitizer_tests.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "run_exact or fused_capture or raster_scanline or insert" 2>&1 | tail -6 >> gpurun_out/r02_sanitizer_tests.txt
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/r02_racecheck_smoke.txt
ncu --set full --import-source on --clock-control none --kernel-name regex:track_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_r02_c3_final \
    python bench.py --realizations 1000 --steps 1 --warmup 3 --no-cpu --no-e2e --legs none > gpurun_out/prof_r02_c3_final.log 2>&1
cat gpurun_out/r02_sanitizer_smoke.txt gpurun_out/r02_sanitizer_tests.txt gpurun_out/r02_racecheck_smoke.txt
