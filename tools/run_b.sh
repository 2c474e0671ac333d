cd $GRAFT_REPO_ROOT
HB_REPS=2 HB_CONST=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hist_a -c 4 -o gpurun_out/r02b_hist_ch python tools/hist_bench.py > gpurun_out/r02b_ncu.log 2>&1
tail -5 gpurun_out/r02b_ncu.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
