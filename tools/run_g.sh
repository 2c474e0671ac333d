cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -4
(timeout 1200 python -m pytest tests/test_distributed.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r02g_multigpu_tests.txt; cat gpurun_out/r02g_multigpu_tests.txt
(timeout 900 python -m pytest tests/test_gpu_goss.py tests/test_gpu_golden.py -x -q 2>&1 | tail -8) > gpurun_out/r02g_goss_golden_tests.txt; cat gpurun_out/r02g_goss_golden_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench_c3_2gpu.json 2> gpurun_out/r02g_bench_c3_2gpu.err; tail -3 gpurun_out/r02g_bench_c3_2gpu.err; cat gpurun_out/r02g_bench_c3_2gpu.json | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --workload C5 --steps 10 --warmup 10 --no-cpu-baseline > gpurun_out/r02g_bench_c5_2gpu.json 2> gpurun_out/r02g_bench_c5_2gpu.err; tail -5 gpurun_out/r02g_bench_c5_2gpu.err; cat gpurun_out/r02g_bench_c5_2gpu.json | cut -c1-1500
