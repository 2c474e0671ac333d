cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NG:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_c3_${N}gpu.json 2> gpurun_out/r02o_bench_c3_${N}gpu.err
tail -3 gpurun_out/r02o_bench_c3_${N}gpu.err | cut -c1-300; cat gpurun_out/r02o_bench_c3_${N}gpu.json | cut -c1-2500
