set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
for c in 8 4 16 32; do
  STG_AGG_CHUNK=$c timeout 300 python scripts/r2_agg_sweep.py gpurun_out/r2b_sweep_c$c.json > gpurun_out/r2b_sweep_c$c.log 2>&1
  tail -1 gpurun_out/r2b_sweep_c$c.log | cut -c1-1500
done
STG_SLICE_QUEUE=1 timeout 300 python scripts/slice_timing.py 8 > gpurun_out/r2b_slice_q.log 2>&1; tail -3 gpurun_out/r2b_slice_q.log
STG_SLICE_QUEUE=0 timeout 300 python scripts/slice_timing.py 8 > gpurun_out/r2b_slice_s.log 2>&1; tail -3 gpurun_out/r2b_slice_s.log
timeout 300 python scripts/r2_reuse_table.py gpurun_out/r2b_reuse.json > gpurun_out/r2b_reuse.log 2>&1; tail -8 gpurun_out/r2b_reuse.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__t_sector_hit_rate.pct
STG_AGG_CHUNK=8 timeout 300 ncu --metrics $M --clock-control none -k regex:"agg_" -s 2 -c 4 --csv --log-file gpurun_out/r2b_ncu_q8.csv python scripts/prof_agg.py > gpurun_out/r2b_ncu_q8.log 2>&1
STG_AGG_CHUNK=0 timeout 300 ncu --metrics $M --clock-control none -k regex:"agg_" -s 2 -c 4 --csv --log-file gpurun_out/r2b_ncu_q0.csv python scripts/prof_agg.py > gpurun_out/r2b_ncu_q0.log 2>&1
tail -5 gpurun_out/r2b_ncu_q8.csv | cut -c1-400
