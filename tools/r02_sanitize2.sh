#!/bin/bash
# compute-sanitizer memcheck over the rasteriser flavours (the parity module runs every test with the flavour by lattice, heavy forced and
# plain forced) and over the smoke test.  One GPU.
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "run_exact or fused_capture or raster_scanline or insert or raster_flat or raster_random or batches" 2>&1 | tail -6 > gpurun_out/r02_sanitizer_flavours.txt
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 >> gpurun_out/r02_sanitizer_flavours.txt
timeout 600 env ONEKA_RASTER_MODE=heavy compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 >> gpurun_out/r02_sanitizer_flavours.txt
cat gpurun_out/r02_sanitizer_flavours.txt
