set -x
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_layers.py -k "hub_rows" -x -q) > gpurun_out/tests_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_q.log
(timeout 170 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_layers.py -k "test_fused_edge_softmax_hub_rows and 8-16" -x -q) > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck.log
tail -3 gpurun_out/tests_q.log; grep -E "RACECHECK SUMMARY|hazard|passed|failed|rc=" gpurun_out/racecheck.log | head -10
