set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q) > gpurun_out/tests_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_n.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/tests_n.log; cat gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err
