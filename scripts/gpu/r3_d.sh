set -x
mkdir -p gpurun_out
NP=${NP:-2}
STG_HALO_MODE=ce timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 tests/dist_worker.py > gpurun_out/r3d_worker.log 2>&1; echo "worker rc=$?"; tail -3 gpurun_out/r3d_worker.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NP --steps 20 --warmup 5 > gpurun_out/r3d_bench_n$NP.log 2> gpurun_out/r3d_bench_n$NP.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r3d_bench_n$NP.log
