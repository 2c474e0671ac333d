cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4) | tee gpurun_out/r02x_pytest_gpu_1gpu.txt
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee gpurun_out/r02x_smoke.txt
FB_ROWS=2000000 FB_COLS=256 FB_TREES=100 timeout 1200 python tools/f34_bench.py 2>gpurun_out/f34.err | tail -1 > gpurun_out/r02x_f34_bench_2Mx256.json; tail -2 gpurun_out/f34.err
python -c "
import json
d=json.load(open('gpurun_out/r02x_f34_bench_2Mx256.json'))
print(d['dataset_construction']['b200']); print(d['predict'])"
timeout 900 python bench.py > gpurun_out/r02x_bench_c3_1gpu_default.json 2> gpurun_out/r02x_bench.err; tail -2 gpurun_out/r02x_bench.err | cut -c1-200; cut -c1-600 gpurun_out/r02x_bench_c3_1gpu_default.json
timeout 900 python bench.py --impl dropin_device --steps 10 --warmup 2 > gpurun_out/r02x_bench_c3_dropin_device.json 2> gpurun_out/r02x_dd.err; tail -2 gpurun_out/r02x_dd.err | cut -c1-200; cut -c1-400 gpurun_out/r02x_bench_c3_dropin_device.json
timeout 900 python bench.py --impl dropin --steps 10 --warmup 2 > gpurun_out/r02x_bench_c3_dropin.json 2> gpurun_out/r02x_d.err; tail -2 gpurun_out/r02x_d.err | cut -c1-200; cut -c1-400 gpurun_out/r02x_bench_c3_dropin.json
