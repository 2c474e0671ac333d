cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=$(date +%s)
(timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -6) | tee gpurun_out/r02t_pytest_gpu_1gpu.txt
echo "pytest wall $(( $(date +%s) - S )) s"
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee gpurun_out/r02t_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02t_bench_c3_1gpu.json 2> gpurun_out/r02t_bench_c3_1gpu.err; tail -2 gpurun_out/r02t_bench_c3_1gpu.err | cut -c1-300; cut -c1-2600 gpurun_out/r02t_bench_c3_1gpu.json
for W in C2 C4 C5; do
  timeout 900 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_bench_${W}_1gpu.json 2> gpurun_out/r02t_bench_${W}.err; tail -1 gpurun_out/r02t_bench_${W}.err | cut -c1-300
  python -c "
import json,sys
for l in open('gpurun_out/r02t_bench_${W}_1gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$W', d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['roofline']['frac'])"
done
