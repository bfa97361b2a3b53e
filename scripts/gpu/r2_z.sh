set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_dynamic.py -m gpu -q -x -k "tgcn" > gpurun_out/r2z2_tests.log 2>&1; tail -4 gpurun_out/r2z2_tests.log
STG_CONFIGS_OUT=gpurun_out/r2z2_c2.json timeout 400 python scripts/bench_configs.py 2 > gpurun_out/r2z2_c2.log 2>&1; tail -48 gpurun_out/r2z2_c2.log
STG_CONFIGS_OUT=gpurun_out/r2z2_c4.json timeout 400 python scripts/bench_configs.py 4 > gpurun_out/r2z2_c4.log 2>&1; grep -E "epoch_ms|error" gpurun_out/r2z2_c4.json
