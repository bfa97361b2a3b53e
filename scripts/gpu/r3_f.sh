set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tn.py tests/test_gpu_layers.py -m gpu -q -x -k "gcn or gemm" > gpurun_out/r3f_tests.log 2>&1; tail -4 gpurun_out/r3f_tests.log
timeout 300 python scripts/r3_gcn_epoch_prof.py gpurun_out/r3f_gcn_prof.json > gpurun_out/r3f_gcn_prof.log 2>&1; tail -3 gpurun_out/r3f_gcn_prof.log
