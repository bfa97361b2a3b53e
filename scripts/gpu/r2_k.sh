set -x
mkdir -p gpurun_out
STG_CONFIGS_OUT=gpurun_out/r2k_c3_peep.json timeout 200 python scripts/bench_configs.py 3 > gpurun_out/r2k_c3_peep.log 2>&1; grep -E "stock|fused_" gpurun_out/r2k_c3_peep.json
STG_PEEPHOLE=0 STG_CONFIGS_OUT=gpurun_out/r2k_c3_nopeep.json timeout 200 python scripts/bench_configs.py 3 > gpurun_out/r2k_c3_nopeep.log 2>&1; grep -E "stock" gpurun_out/r2k_c3_nopeep.json
STG_PEEPHOLE=0 STG_SHARE_EDGE=0 STG_CONFIGS_OUT=gpurun_out/r2k_c3_r1.json timeout 200 python scripts/bench_configs.py 3 > gpurun_out/r2k_c3_r1.log 2>&1; grep -E "stock" gpurun_out/r2k_c3_r1.json
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q -k "gat or clamp or stack or Gat or GAT or softmax" > gpurun_out/r2k_tests.log 2>&1; tail -5 gpurun_out/r2k_tests.log
STG_PEEPHOLE=0 timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_golden.py -m gpu -q -k "gat or Gat or GAT" > gpurun_out/r2k_tests_nopeep.log 2>&1; tail -3 gpurun_out/r2k_tests_nopeep.log
