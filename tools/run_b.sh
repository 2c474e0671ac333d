cd $GRAFT_REPO_ROOT
HB_ROWS=2000000 HB_COLS=1024 HB_REPS=1 HB_CONST=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hist_a -c 3 -o gpurun_out/r02f_hist_ch python tools/hist_bench.py > gpurun_out/r02f_ncu.log 2>&1
tail -3 gpurun_out/r02f_ncu.log
TB_ROWS=2000000 TB_COLS=1024 TB_LEAVES=63 TB_TREES=1 TB_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_select|k_part_flags|k_part_scatter|k_hist_reduce" -s 50 -c 10 -o gpurun_out/r02f_chain python tools/tree_bench.py > gpurun_out/r02f_ncu2.log 2>&1
tail -3 gpurun_out/r02f_ncu2.log
