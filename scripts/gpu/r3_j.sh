set -x
mkdir -p gpurun_out
NP=${NP:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NP --steps 20 --warmup 5 > gpurun_out/r3j_bench_n$NP.log 2> gpurun_out/r3j_bench_n$NP.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r3j_bench_n$NP.log
