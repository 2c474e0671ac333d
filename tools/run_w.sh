cd $GRAFT_REPO_ROOT
for K in 0.044 0.1 0.25 0.6; do
  echo "== split_k $K"
  LGBMB200_SPLIT_K=$K TB_ROWS=10000000 TB_COLS=128 TB_LEAVES=127 TB_TREES=5 timeout 300 python tools/tree_bench.py 2>&1 | tail -1 | cut -c1-150
  LGBMB200_SPLIT_K=$K TB_ROWS=2000000 TB_COLS=1024 TB_LEAVES=127 TB_TREES=5 timeout 300 python tools/tree_bench.py 2>&1 | tail -1 | cut -c1-150
  LGBMB200_SPLIT_K=$K TB_ROWS=1000000 TB_COLS=256 TB_LEAVES=63 TB_TREES=8 timeout 300 python tools/tree_bench.py 2>&1 | tail -1 | cut -c1-150
done
(timeout 600 python -m pytest tests/test_binning.py -x -q -m gpu 2>&1 | tail -2)
