cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests/test_gpu_scale.py -x -q -s 2>&1 | grep -E "iter |splits compared|passed|failed|Error|error" | tail -30) > gpurun_out/r02c_scaletest.txt
cat gpurun_out/r02c_scaletest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_c3_1gpu.json 2> gpurun_out/r02c_bench_c3.err; tail -3 gpurun_out/r02c_bench_c3.err; cat gpurun_out/r02c_bench_c3_1gpu.json
timeout 600 python bench.py --workload C2 --steps 20 --warmup 5 > gpurun_out/r02c_bench_c2_1gpu.json 2> gpurun_out/r02c_bench_c2.err; tail -3 gpurun_out/r02c_bench_c2.err; cat gpurun_out/r02c_bench_c2_1gpu.json
timeout 600 python bench.py --workload C2 --impl reference --steps 20 --warmup 5 > gpurun_out/r02c_bench_c2_ref.json 2> gpurun_out/r02c_bench_c2_ref.err; tail -3 gpurun_out/r02c_bench_c2_ref.err; cat gpurun_out/r02c_bench_c2_ref.json
timeout 600 python bench.py --workload C2 --impl reference_cuda --steps 20 --warmup 5 > gpurun_out/r02c_bench_c2_refcuda.json 2> gpurun_out/r02c_bench_c2_refcuda.err; tail -5 gpurun_out/r02c_bench_c2_refcuda.err; cat gpurun_out/r02c_bench_c2_refcuda.json
