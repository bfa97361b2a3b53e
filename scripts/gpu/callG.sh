set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_agg.py tests/test_gpu_packed.py -x -q) > gpurun_out/tests_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_g.log
for r in 32 16 8 4 0; do echo "rows_per_warp=$r"; STG_AGG_ROWS_PER_WARP=$r timeout 200 python scripts/slice_timing.py 8; done > gpurun_out/slice.log 2>&1
tail -3 gpurun_out/tests_g.log; cat gpurun_out/slice.log
