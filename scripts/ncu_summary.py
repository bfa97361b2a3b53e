"""Print the metrics quoted in profiles/*.md from an .ncu-rep (needs `ncu` on PATH): python scripts/ncu_summary.py rep [...]."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'lts__t_sectors.sum',
        'lts__t_sectors_srcunit_tex.sum', 'lts__t_requests_srcunit_tex.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'lts__cycles_elapsed.avg.per_second', 'lts__t_sectors.sum.per_second',
        'l1tex__m_xbar2l1tex_read_bytes.sum.per_second']
for rep in sys.argv[1:]:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('==', rep, vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f'  {h:62s} {vals[i]:>18s} {units[i]}')
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    if float(vals[i]) > 0.5:
                        print(f'  {h:62s} {float(vals[i]):18.2f}')
                except ValueError:
                    pass
