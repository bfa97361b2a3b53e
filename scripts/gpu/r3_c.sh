set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_dynamic.py tests/test_gpu_gemm_tn.py -m gpu -q -x -k "tgcn or gru or clamp or gemm" > gpurun_out/r3c_tests.log 2>&1; tail -5 gpurun_out/r3c_tests.log
STG_CONFIGS_OUT=gpurun_out/r3c_c4.json timeout 400 python scripts/bench_configs.py 4 > gpurun_out/r3c_c4.log 2>&1; grep -E "epoch_ms|error" gpurun_out/r3c_c4.json
timeout 400 python scripts/r2_config4_prof.py gpurun_out/r3c_c4prof.json > gpurun_out/r3c_c4prof.log 2>&1
