cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_binning.py tests/test_model.py -x -q -m gpu 2>&1 | tail -4)
BB_ROWS=2000000 BB_COLS=1024 timeout 600 python tools/binning_bench.py 2>&1 | tail -3
BB_ROWS=4000000 BB_COLS=128 timeout 600 python tools/binning_bench.py 2>&1 | tail -2 | head -1
FB_ROWS=2000000 FB_COLS=256 FB_TREES=100 timeout 1200 python tools/f34_bench.py 2>gpurun_out/f34.err | tail -1 > gpurun_out/r02s_f34_bench_2Mx256.json; tail -3 gpurun_out/f34.err
python -c "
import json
d=json.load(open('gpurun_out/r02s_f34_bench_2Mx256.json'))
print(d['dataset_construction']); print(d['predict'])"
LGBMB200_PRED_PASS=16 FB_ROWS=2000000 FB_COLS=256 FB_TREES=100 timeout 1200 python tools/f34_bench.py 2>gpurun_out/f34.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('pass16', d['predict']['b200'])"
