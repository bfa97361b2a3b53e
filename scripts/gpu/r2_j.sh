set -x
mkdir -p gpurun_out
timeout 300 python scripts/r2_config4_prof.py gpurun_out/r2j_c4prof.json > gpurun_out/r2j_c4prof.log 2>&1; tail -60 gpurun_out/r2j_c4prof.log | cut -c1-200
for c in 0 4 2; do
STG_GAT_CHUNK=$c STG_CONFIGS_OUT=gpurun_out/r2j_c3_gat$c.json timeout 200 python scripts/bench_configs.py 3 > gpurun_out/r2j_c3_gat$c.log 2>&1; grep -E "fwd|bwd" gpurun_out/r2j_c3_gat$c.json
done
STG_PEEPHOLE=0 STG_CONFIGS_OUT=gpurun_out/r2j_c3_nopeep.json timeout 200 python scripts/bench_configs.py 3 > gpurun_out/r2j_c3_nopeep.log 2>&1; grep -E "stock" gpurun_out/r2j_c3_nopeep.json
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q -k "gat or clamp or stack or Gat or GAT or softmax" > gpurun_out/r2j_tests.log 2>&1; tail -5 gpurun_out/r2j_tests.log
