set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_golden.py -x -q) > gpurun_out/tests_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_d.log
for thr in 1024 256 128 64; do STG_GAT_HUB_THRESHOLD=$thr timeout 120 python scripts/bench_gat.py; done > gpurun_out/bench_gat.log 2>&1
tail -3 gpurun_out/tests_d.log; cat gpurun_out/bench_gat.log
