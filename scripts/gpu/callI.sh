set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log


tail -4 gpurun_out/tests.log
