set -x
mkdir -p gpurun_out
STG_SWEEP_F=48,64,47,36 timeout 300 python scripts/r2_agg_sweep.py gpurun_out/r3g_narrow.json > gpurun_out/r3g_narrow.log 2>&1; grep -E "^(48|64|47|36) " gpurun_out/r3g_narrow.log | cut -c1-260
