set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r01c.txt
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/sanitizer_smoke_r01c.txt
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_c3_ref_r01c.json
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_c3_r01c.json
python bench.py --workload c4 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_c4_r01c.json
python bench.py --workload c5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_c5_r01c.json
python bench.py --workload c1 2>/dev/null | tail -1 > gpurun_out/bench_c1_r01c.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 1 --realizations 2000 --no-cpu > gpurun_out/launches_r01c_bench.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:track_kernel -c 2 -o gpurun_out/prof_track_r01c python bench.py --realizations 1000 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
head -c 600 gpurun_out/bench_c3_r01c.json
ncu --set full --clock-control none --kernel-name regex:flush_kernel -c 1 -o gpurun_out/prof_flush_r01c python bench.py --workload c5 --realizations 512 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
