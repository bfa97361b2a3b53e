set -x
mkdir -p gpurun_out
(timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_packed.py -k "hub_rows or pair_form or padded or nonfinite" -x -q) > gpurun_out/memcheck_agg.log 2>&1; echo "rc=$?" >> gpurun_out/memcheck_agg.log
(timeout 120 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_packed.py tests/test_gpu_layers.py -k "test_packed_hub_rows or test_stock_gat_vm_kernel_hub_rows" -x -q) > gpurun_out/racecheck_agg.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck_agg.log
(timeout 120 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_layers.py -k "hub_rows or gru" -x -q) > gpurun_out/memcheck_layers.log 2>&1; echo "rc=$?" >> gpurun_out/memcheck_layers.log
for f in memcheck_agg racecheck_agg memcheck_layers; do echo == $f; grep -E "SUMMARY|passed|failed|^rc=|Invalid|hazard" gpurun_out/$f.log | head -6; done
