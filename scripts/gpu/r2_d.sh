set -x
mkdir -p gpurun_out
export STG_SWEEP_F=100
for cfg in "1 0" "2 0" "3 0" "4 0" "6 0" "4 16384" "4 40000" "2 16384"; do
  set -- $cfg
  STG_AGG_CHUNK=$1 STG_AGG_FAR=$2 timeout 200 python scripts/r2_agg_sweep.py gpurun_out/r2d_sweep_c$1_f$2.json > gpurun_out/r2d_sweep_c$1_f$2.log 2>&1
  echo "chunk $1 far $2: $(tail -1 gpurun_out/r2d_sweep_c$1_f$2.log | cut -c1-420)"
done
timeout 200 python scripts/slice_timing.py 8 > gpurun_out/r2d_slice8.log 2>&1; tail -2 gpurun_out/r2d_slice8.log
timeout 200 python scripts/slice_timing.py 2 > gpurun_out/r2d_slice2.log 2>&1; tail -2 gpurun_out/r2d_slice2.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2d_tests.log 2>&1; tail -15 gpurun_out/r2d_tests.log
