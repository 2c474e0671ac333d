cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NG:-2}
if [ "$N" = "2" ]; then
  (timeout 1500 python -m pytest tests/test_distributed.py tests/test_gpu_dropin.py -q -m gpu 2>&1 | tail -5) | tee gpurun_out/r02t_pytest_gpu_2gpu.txt
  (timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee gpurun_out/r02t_smoke_2gpu.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_bench_c3_${N}gpu.json 2> gpurun_out/r02t_bench_c3_${N}gpu.err
tail -2 gpurun_out/r02t_bench_c3_${N}gpu.err | cut -c1-300; cut -c1-2600 gpurun_out/r02t_bench_c3_${N}gpu.json
if [ "$N" = "2" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload C5 --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/r02t_bench_c5_${N}gpu.json 2> gpurun_out/r02t_bench_c5_${N}gpu.err
tail -2 gpurun_out/r02t_bench_c5_${N}gpu.err | cut -c1-300; cut -c1-1500 gpurun_out/r02t_bench_c5_${N}gpu.json
fi
