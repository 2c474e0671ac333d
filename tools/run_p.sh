cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_binning.py -x -q -m gpu -s 2>&1 | tail -8)
BB_ROWS=2000000 BB_COLS=1024 timeout 600 python tools/binning_bench.py 2>&1 | tail -5
BB_ROWS=4000000 BB_COLS=128 timeout 600 python tools/binning_bench.py 2>&1 | tail -5
