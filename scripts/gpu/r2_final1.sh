set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2z_tests.log 2>&1; tail -6 gpurun_out/r2z_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
(time timeout 900 python bench.py) > gpurun_out/r2z_bench.log 2> gpurun_out/r2z_bench.err; tail -1 gpurun_out/r2z_bench.log | cut -c1-1500; tail -3 gpurun_out/r2z_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:"agg_|pack_" -c 40 --csv --log-file gpurun_out/r2z_bench_launches.csv python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/r2z_bench_ncu.log 2>&1
tail -4 gpurun_out/r2z_bench_launches.csv | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_rows_pipe -s 2 -c 1 -o gpurun_out/r2z_agg_full python scripts/prof_agg.py > gpurun_out/r2z_agg_full.log 2>&1; ls -la gpurun_out/r2z_agg_full.ncu-rep
