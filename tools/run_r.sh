cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_binning.py tests/test_model.py -x -q -m gpu 2>&1 | tail -4)
BB_ROWS=2000000 BB_COLS=1024 timeout 600 python tools/binning_bench.py 2>&1 | tail -3
LGBMB200_BIN_SIMPLE=1 BB_ROWS=2000000 BB_COLS=1024 timeout 600 python tools/binning_bench.py 2>&1 | tail -2 | head -1
BB_ROWS=4000000 BB_COLS=128 timeout 600 python tools/binning_bench.py 2>&1 | tail -2 | head -1
FB_ROWS=2000000 FB_COLS=256 FB_TREES=100 timeout 1200 python tools/f34_bench.py 2>gpurun_out/f34.err | tail -1 > gpurun_out/r02r_f34_bench_2Mx256.json; tail -3 gpurun_out/f34.err; cut -c1-1800 gpurun_out/r02r_f34_bench_2Mx256.json
BB_ROWS=1000000 BB_COLS=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_value_to_bin_tile -s 2 -c 1 -o gpurun_out/r02r_vtb python tools/binning_bench.py > gpurun_out/r02r_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/r02r_vtb.ncu-rep 18 > gpurun_out/r02r_value_to_bin_tile_1Mx1024.txt 2>&1; head -40 gpurun_out/r02r_value_to_bin_tile_1Mx1024.txt
FB_ROWS=500000 FB_COLS=256 FB_TREES=100 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_predict -s 3 -c 1 -o gpurun_out/r02r_pred python tools/f34_bench.py > gpurun_out/r02r_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02r_pred.ncu-rep 18 > gpurun_out/r02r_predict_500Kx256x100.txt 2>&1; head -40 gpurun_out/r02r_predict_500Kx256x100.txt
