cd $GRAFT_REPO_ROOT
for k in 0.011 0.044 0.176 0.7; do
  echo "== SPLIT_K=$k  10M x 128 (the per-GPU shard of C3 at 8 GPUs), 127 leaves"
  LGBMB200_SPLIT_K=$k TB_ROWS=10000000 TB_COLS=128 TB_LEAVES=127 TB_TREES=4 timeout 300 python tools/tree_bench.py 2>&1 | tail -1
  echo "== SPLIT_K=$k  2M x 1024"
  LGBMB200_SPLIT_K=$k TB_ROWS=2000000 TB_COLS=1024 TB_LEAVES=127 TB_TREES=4 timeout 300 python tools/tree_bench.py 2>&1 | tail -1
done
LGBMB200_SPLIT_K=0.044 TB_ROWS=10000000 TB_COLS=128 TB_LEAVES=127 TB_TREES=3 TB_PROFILE=1 timeout 300 python tools/tree_bench.py 2>&1 | tail -2
