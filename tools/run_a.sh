cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -8) > gpurun_out/r02a_gputest.txt
(HB_CONST=1 timeout 300 python tools/hist_bench.py; HB_CONST=0 timeout 300 python tools/hist_bench.py; HB_ROWS=4000000 HB_COLS=1024 HB_CONST=1 timeout 300 python tools/hist_bench.py; HB_ROWS=4000000 HB_COLS=1024 HB_CONST=0 timeout 300 python tools/hist_bench.py) > gpurun_out/r02a_histbench.txt 2>&1
(TB_ROWS=2000000 TB_COLS=1024 TB_LEAVES=127 TB_PROFILE=1 timeout 300 python tools/tree_bench.py) > gpurun_out/r02a_treebench.txt 2>&1
cat gpurun_out/r02a_gputest.txt gpurun_out/r02a_histbench.txt gpurun_out/r02a_treebench.txt
(timeout 1500 python -m pytest tests/test_gpu_scale.py -x -q -s 2>&1 | tail -30) > gpurun_out/r02a_scaletest.txt
cat gpurun_out/r02a_scaletest.txt
