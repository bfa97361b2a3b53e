set -x
mkdir -p gpurun_out
NP=${NP:-2}
STG_HALO_MODE=ce timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 tests/dist_worker.py > gpurun_out/r3l_worker.log 2>&1; echo "worker rc=$?"; tail -12 gpurun_out/r3l_worker.log
