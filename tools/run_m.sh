cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edges.py tests/test_gpu_quantized.py -x -q 2>&1 | tail -3)
echo "== 10M x 128 (the per-GPU shard of C3 at 8 GPUs), 127 leaves"
TB_ROWS=10000000 TB_COLS=128 TB_LEAVES=127 TB_TREES=4 TB_PROFILE=1 timeout 300 python tools/tree_bench.py 2>&1 | tail -2
echo "== 2M x 1024"
TB_ROWS=2000000 TB_COLS=1024 TB_LEAVES=127 TB_TREES=4 TB_PROFILE=1 timeout 300 python tools/tree_bench.py 2>&1 | tail -2
echo "== 5M x 256"
TB_ROWS=5000000 TB_COLS=256 TB_LEAVES=127 TB_TREES=4 TB_PROFILE=1 timeout 300 python tools/tree_bench.py 2>&1 | tail -2
