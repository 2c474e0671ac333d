cd $GRAFT_REPO_ROOT
for g4 in 0 1; do
  echo "== LGBMB200_G4=$g4"
  (LGBMB200_G4=$g4 HB_ROWS=4000000 HB_COLS=1024 HB_CONST=1 timeout 300 python tools/hist_bench.py; LGBMB200_G4=$g4 HB_ROWS=4000000 HB_COLS=1024 HB_CONST=0 timeout 300 python tools/hist_bench.py) 2>&1 | grep -E "gather|constant"
  LGBMB200_G4=$g4 TB_ROWS=2000000 TB_COLS=1024 TB_LEAVES=127 TB_TREES=4 timeout 300 python tools/tree_bench.py 2>&1 | tail -1
done
(LGBMB200_G4=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edges.py -x -q 2>&1 | tail -4)
