cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_model.py -x -q -m gpu 2>&1 | tail -3)
for P in 16 32; do
LGBMB200_PRED_PASS=$P FB_ROWS=2000000 FB_COLS=256 FB_TREES=100 timeout 1200 python tools/f34_bench.py 2>gpurun_out/f34.err | tail -1 > gpurun_out/r02v_f34_bench_2Mx256_p$P.json; tail -2 gpurun_out/f34.err
python -c "
import json
d=json.load(open('gpurun_out/r02v_f34_bench_2Mx256_p$P.json'))
print('pass $P', d['predict'])"
done
