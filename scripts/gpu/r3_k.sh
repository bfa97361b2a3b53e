set -x
mkdir -p gpurun_out
STG_CONFIGS_OUT=gpurun_out/r3k_ref.json timeout 200 python scripts/bench_configs.py ref > gpurun_out/r3k_ref.log 2>&1; grep -E "config[12]_gcn|our|error" -A3 gpurun_out/r3k_ref.json | head -40
