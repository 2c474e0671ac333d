cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02h_gputest.txt; cat gpurun_out/r02h_gputest.txt
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3)
export TB_ROWS=2000000 TB_COLS=1024 TB_LEAVES=127 TB_TREES=4
for v in "LGBMB200_DEBUG=0" "LGBMB200_DEBUG=512"; do echo "== $v"; env $v timeout 300 python tools/tree_bench.py 2>&1 | tail -1; done
TB_PROFILE=1 timeout 300 python tools/tree_bench.py 2>&1 | tail -2
timeout 900 python bench.py --workload C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_c4_1gpu.json 2> gpurun_out/r02h_bench_c4.err; tail -3 gpurun_out/r02h_bench_c4.err; cat gpurun_out/r02h_bench_c4_1gpu.json | cut -c1-1800
timeout 900 python bench.py --workload C5 --steps 10 --warmup 10 --no-cpu-baseline > gpurun_out/r02h_bench_c5_1gpu.json 2> gpurun_out/r02h_bench_c5.err; tail -3 gpurun_out/r02h_bench_c5.err; cat gpurun_out/r02h_bench_c5_1gpu.json | cut -c1-1800
