set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_golden.py tests/test_gpu_dynamic.py -x -q) > gpurun_out/tests_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_m.log
timeout 300 python - <<'PY' > gpurun_out/refgat.log 2>&1
import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scripts')
import bench_reference_gpu as B
B.run_gat()
json.dump(B.out, open('gpurun_out/reference_gpu_gat.json', 'w'), indent=1)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vm_kernel" -s 20 -c 12 --csv --log-file gpurun_out/vm_launches.csv python scripts/prof_gat_stock.py > gpurun_out/vm_ncu.log 2>&1
tail -3 gpurun_out/tests_m.log; tail -2 gpurun_out/refgat.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/vm_launches.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d['ID'], d['Kernel Name'][:50], d['Grid Size'], d['Block Size'], d['Metric Value'])
PY
