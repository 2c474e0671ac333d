cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_binning.py tests/test_model.py -x -q -m gpu 2>&1 | tail -3)
BB_ROWS=2000000 BB_COLS=1024 timeout 600 python tools/binning_bench.py 2>&1 | tail -3
BB_ROWS=4000000 BB_COLS=128 timeout 600 python tools/binning_bench.py 2>&1 | tail -2 | head -1
FB_ROWS=2000000 FB_COLS=256 FB_TREES=100 timeout 1200 python tools/f34_bench.py 2>gpurun_out/f34.err | tail -1 > gpurun_out/r02u_f34_bench_2Mx256.json; tail -3 gpurun_out/f34.err
python -c "
import json
d=json.load(open('gpurun_out/r02u_f34_bench_2Mx256.json'))
print(d['dataset_construction']); print(d['predict'])"
BB_ROWS=1000000 BB_COLS=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_value_to_bin_tile|k_bundle_columns" -s 8 -c 2 -o gpurun_out/r02u_vtb python tools/binning_bench.py > gpurun_out/r02u_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/r02u_vtb.ncu-rep 10 > gpurun_out/r02u_value_to_bin_tile_1Mx1024.txt 2>&1; cat gpurun_out/r02u_value_to_bin_tile_1Mx1024.txt | cut -c1-200
ncu -i gpurun_out/r02u_vtb.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    d=dict(zip(h,r)); print(d['Kernel Name'][:40], d['gpu__time_duration.sum'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], d['smsp__inst_executed.sum'])"
