set -x
mkdir -p gpurun_out
export STG_SWEEP_F=100
for m in 2 1; do
  STG_HUB_GRID=$m timeout 200 python scripts/r2_agg_sweep.py gpurun_out/r2q_hubgrid$m.json > gpurun_out/r2q_hubgrid$m.log 2>&1
  echo "hubgrid $m: $(tail -1 gpurun_out/r2q_hubgrid$m.log | cut -c1-330)"
done
STG_HUB_GRID=1 STG_HUB_THRESHOLD=2048 timeout 200 python scripts/r2_agg_sweep.py gpurun_out/r2q_hubgrid1_t2048.json > gpurun_out/r2q_hubgrid1_t2048.log 2>&1; echo "hubgrid 1 thr 2048: $(tail -1 gpurun_out/r2q_hubgrid1_t2048.log | cut -c1-330)"
