set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_packed.py tests/test_gpu_agg.py tests/test_gpu_golden.py -x -q) > gpurun_out/tests_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_b.log
STG_AGG_PAIR=0 FEATS=100,128,96 timeout 240 python scripts/packed_ab.py > gpurun_out/ab_pair0.log 2>&1
STG_AGG_PAIR=1 FEATS=100,128,96 timeout 240 python scripts/packed_ab.py > gpurun_out/ab_pair1.log 2>&1
tail -4 gpurun_out/tests_b.log; cat gpurun_out/ab_pair0.log gpurun_out/ab_pair1.log
