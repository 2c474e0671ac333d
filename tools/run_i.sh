cd $GRAFT_REPO_ROOT
S=$(date +%s)
timeout 1700 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02i_bench_c3_reference_full.json 2> gpurun_out/r02i_ref.err
echo "rc=$? wall=$(( $(date +%s) - S ))s"; tail -3 gpurun_out/r02i_ref.err; cat gpurun_out/r02i_bench_c3_reference_full.json | cut -c1-1500
