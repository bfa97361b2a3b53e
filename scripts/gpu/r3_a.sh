set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tn.py tests/test_abi.py -m gpu -q -x > gpurun_out/r3a_tests.log 2>&1; tail -8 gpurun_out/r3a_tests.log
timeout 200 python scripts/bench_gemm_tn.py > gpurun_out/r3a_gemm.log 2>&1; tail -8 gpurun_out/r3a_gemm.log | cut -c1-250
STG_CONFIGS_OUT=gpurun_out/r3a_c4.json timeout 400 python scripts/bench_configs.py 4 > gpurun_out/r3a_c4.log 2>&1; grep -E "epoch_ms|error" gpurun_out/r3a_c4.json
