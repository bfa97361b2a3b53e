set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err
tail -4 gpurun_out/r2f_tests.log; tail -2 gpurun_out/r2f_smoke.log; tail -c 600 gpurun_out/r2f_bench.log
