set -x
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_gpu_native_mirrors.py -x -q) > gpurun_out/tests_s.log 2>&1; echo "mirror rc=$?" >> gpurun_out/tests_s.log
(timeout 300 python -m pytest tests -m gpu -q) > gpurun_out/tests_all.log 2>&1; echo "all rc=$?" >> gpurun_out/tests_all.log
grep -v "^$" gpurun_out/tests_s.log | tail -30; tail -4 gpurun_out/tests_all.log
