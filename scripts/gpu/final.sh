set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"agg_|pack_" -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/bench_ncu.log 2>&1
tail -3 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.log
