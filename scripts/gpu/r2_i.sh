set -x
mkdir -p gpurun_out
export STG_SWEEP_F=100
for t in 512 1024 2048; do
  STG_HUB_THRESHOLD=$t timeout 200 python scripts/r2_agg_sweep.py gpurun_out/r2i_sweep_hub$t.json > gpurun_out/r2i_sweep_hub$t.log 2>&1
  echo "hub $t: $(tail -1 gpurun_out/r2i_sweep_hub$t.log | cut -c1-330)"
done
(time timeout 400 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/r2i_ref.log 2> gpurun_out/r2i_ref.err; tail -1 gpurun_out/r2i_ref.log | cut -c1-900; tail -4 gpurun_out/r2i_ref.err
(time timeout 900 python bench.py) > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err; tail -1 gpurun_out/r2i_bench.log | cut -c1-3000; tail -5 gpurun_out/r2i_bench.err
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2i_tests.log 2>&1; tail -8 gpurun_out/r2i_tests.log
