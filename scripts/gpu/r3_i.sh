set -x
mkdir -p gpurun_out
for u in 8 16; do
STG_AGG_NARROW_UNROLL=$u STG_SWEEP_F=48,47,64 timeout 300 python scripts/r2_agg_sweep.py gpurun_out/r3i_u$u.json > gpurun_out/r3i_u$u.log 2>&1; echo "unroll $u"; grep -E "^(48|64|47|36) " gpurun_out/r3i_u$u.log | cut -c1-200
done
