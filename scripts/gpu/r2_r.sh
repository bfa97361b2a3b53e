set -x
mkdir -p gpurun_out
for h in 1 0; do
STG_GAT_HALFWARP=$h STG_CONFIGS_OUT=gpurun_out/r2r_c3_half$h.json timeout 200 python scripts/bench_configs.py 3 > gpurun_out/r2r_c3_half$h.log 2>&1; grep -E "fused_" gpurun_out/r2r_c3_half$h.json
STG_GAT_HALFWARP=$h timeout 100 python scripts/bench_gat.py > gpurun_out/r2r_gat_raw$h.log 2>&1; tail -3 gpurun_out/r2r_gat_raw$h.log
done
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_fullsize.py -m gpu -q -k "softmax or fused" > gpurun_out/r2r_tests.log 2>&1; tail -4 gpurun_out/r2r_tests.log
