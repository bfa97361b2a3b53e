set -x
mkdir -p gpurun_out
STG_PEEPHOLE=0 STG_SHARE_EDGE=0 timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_golden.py -m gpu -q -k "gat or Gat or GAT" > gpurun_out/r2l_a.log 2>&1; tail -3 gpurun_out/r2l_a.log
STG_PEEPHOLE=1 STG_SHARE_EDGE=0 timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_golden.py -m gpu -q -k "gat or Gat or GAT" > gpurun_out/r2l_b.log 2>&1; tail -3 gpurun_out/r2l_b.log
