cd $GRAFT_REPO_ROOT
FB_ROWS=500000 FB_COLS=256 FB_TREES=100 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_predict_f -s 3 -c 1 -o gpurun_out/r02z_predf python tools/f34_bench.py > gpurun_out/r02z_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02z_predf.ncu-rep 24 > gpurun_out/r02z_predict_f_500Kx256x100.txt 2>&1; cat gpurun_out/r02z_predict_f_500Kx256x100.txt | cut -c1-220
