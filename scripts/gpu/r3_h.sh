set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r3h_tests.log 2>&1; tail -5 gpurun_out/r3h_tests.log
timeout 300 python scripts/r3_gcn_epoch_prof.py gpurun_out/r3h_gcn_prof.json > gpurun_out/r3h_gcn_prof.log 2>&1; head -2 gpurun_out/r3h_gcn_prof.json
