set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
(time timeout 600 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"agg_|pack_" -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/bench_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_rows_pipe -s 2 -c 1 -o gpurun_out/agg_pair_packed_f100 -f python scripts/prof_agg.py 0.9 100 > gpurun_out/ncu_full.log 2>&1
timeout 120 python scripts/bench_gat.py > gpurun_out/bench_gat.log 2>&1
timeout 600 python scripts/bench_configs.py 1 2 3 4 > gpurun_out/configs.log 2>&1
timeout 300 python scripts/bench_reference_gpu.py > gpurun_out/refgpu.log 2>&1
tail -4 gpurun_out/tests.log; cat gpurun_out/bench.log gpurun_out/bench_ref.log gpurun_out/bench_gat.log; tail -5 gpurun_out/configs.log; tail -12 gpurun_out/refgpu.log
