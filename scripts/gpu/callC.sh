set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_golden.py -x -q) > gpurun_out/tests_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_c.log
for thr in 1024 256 128 64; do
STG_GAT_HUB_THRESHOLD=$thr timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gat_ -s 12 -c 6 --csv --log-file gpurun_out/gat_launches_$thr.csv python scripts/prof_gat.py 3 > gpurun_out/gat_ncu_$thr.log 2>&1
done
tail -3 gpurun_out/tests_c.log
python - <<'PY'
import csv
for thr in (1024, 256, 128, 64):
    rows=[r for r in csv.reader(open(f'gpurun_out/gat_launches_{thr}.csv')) if len(r)>10]
    h=rows[0]; tot=0
    for r in rows[1:]:
        d=dict(zip(h,r)); tot+=float(d['Metric Value'])
        print(thr, d['Kernel Name'][15:60], d['Metric Value'])
    print(thr, 'total us', tot/1e3)
PY
