cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# trees 4..6 of `bench.py --steps 1 --warmup 3` are the ones its live roofline block profiles: skip 2 set-up launches + 4 trees of 761
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ --launch-skip 3046 -c 2290 --csv \
  --log-file gpurun_out/r02n_launches_c3_1gpu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02n_ncu_bench.log 2>&1
tail -1 gpurun_out/r02n_ncu_bench.log | cut -c1-300
python tools/launch_list_summary.py gpurun_out/r02n_launches_c3_1gpu.csv gpurun_out/hist_traffic.json C3 1
cat gpurun_out/hist_traffic.json
echo "== hist fixed cost sweep, 128 columns"
HB_ROWS=10000000 HB_COLS=128 HB_SWEEP=1 HB_REPS=3 timeout 300 python tools/hist_bench.py 2>&1 | tail -12
echo "== hist fixed cost sweep, 1024 columns"
HB_ROWS=2000000 HB_COLS=1024 HB_SWEEP=1 HB_REPS=3 timeout 300 python tools/hist_bench.py 2>&1 | tail -12
