set -x
mkdir -p gpurun_out
timeout 400 python scripts/r2_config4_prof.py gpurun_out/r3b_c4prof.json > gpurun_out/r3b_c4prof.log 2>&1; tail -3 gpurun_out/r3b_c4prof.log | cut -c1-300
