set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_l.log
timeout 300 python - <<'PY' > gpurun_out/refgat.log 2>&1
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scripts')
import bench_reference_gpu as B
B.run_gat()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vm_kernel" -s 20 -c 12 --csv --log-file gpurun_out/vm_launches.csv python scripts/prof_gat_stock.py > gpurun_out/vm_ncu.log 2>&1
timeout 600 python scripts/bench_configs.py 2 4 > gpurun_out/configs24.log 2>&1; cp gpurun_out/configs.json gpurun_out/configs24.json
tail -5 gpurun_out/tests_l.log; tail -3 gpurun_out/refgat.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/vm_launches.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d['ID'], d['Kernel Name'][:50], d['Grid Size'], d['Block Size'], d['Metric Value'])
PY
cat gpurun_out/configs24.json
