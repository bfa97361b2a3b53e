set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_layers.py tests/test_gpu_dynamic.py tests/test_gpu_gemm_tn.py tests/test_gpu_packed.py -m gpu -q -x -k "tgcn or gcn or packed" > gpurun_out/r3n_tests.log 2>&1; tail -3 gpurun_out/r3n_tests.log
