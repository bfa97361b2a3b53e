set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum --clock-control none -k regex:gat_ -s 6 -c 6 --csv --log-file gpurun_out/gat_launches_f.csv python scripts/prof_gat.py 2 > gpurun_out/gat_ncu_f.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/gat_launches_f.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r))
    print(d['ID'], d['Kernel Name'][15:52], d['Metric Name'], d['Metric Value'])
PY
