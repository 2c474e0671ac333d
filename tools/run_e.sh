cd $GRAFT_REPO_ROOT
bash tools/run_d.sh > gpurun_out/r02d_sweep.txt 2>&1; cat gpurun_out/r02d_sweep.txt
timeout 1500 python bench.py --impl reference_cuda --steps 10 --warmup 3 > gpurun_out/r02e_bench_c3_refcuda.json 2> gpurun_out/r02e_bench_c3_refcuda.err; tail -3 gpurun_out/r02e_bench_c3_refcuda.err; cat gpurun_out/r02e_bench_c3_refcuda.json
