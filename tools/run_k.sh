cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests/test_gpu_dropin.py -x -q 2>&1 | tail -8) > gpurun_out/r02k_dropin_tests.txt; cat gpurun_out/r02k_dropin_tests.txt
for impl in dropin dropin_device; do
  timeout 600 python bench.py --workload C2 --impl $impl --steps 20 --warmup 5 > gpurun_out/r02k_bench_c2_$impl.json 2> gpurun_out/r02k_$impl.err; tail -2 gpurun_out/r02k_$impl.err; cat gpurun_out/r02k_bench_c2_$impl.json | cut -c1-900
done
for impl in dropin dropin_device; do
  timeout 1200 python bench.py --impl $impl --steps 10 --warmup 3 > gpurun_out/r02k_bench_c3_$impl.json 2> gpurun_out/r02k_c3_$impl.err; tail -2 gpurun_out/r02k_c3_$impl.err; cat gpurun_out/r02k_bench_c3_$impl.json | cut -c1-900
done
