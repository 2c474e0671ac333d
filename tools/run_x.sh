cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4) | tee gpurun_out/r02x_pytest_gpu_1gpu.txt
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee gpurun_out/r02x_smoke.txt
timeout 900 python bench.py > gpurun_out/r02x_bench_c3_1gpu_default.json 2> gpurun_out/r02x_bench.err; tail -2 gpurun_out/r02x_bench.err | cut -c1-200; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r02x_bench_c3_1gpu_default.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_over_alg'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])"
