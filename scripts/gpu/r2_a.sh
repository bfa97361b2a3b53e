set -x
mkdir -p gpurun_out
export STG_DIST_PROFILE=1
for pb in 32 64; do
STG_PUSH_BLOCKS=$pb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2a_n8_pb$pb.log 2> gpurun_out/r2a_n8_pb$pb.err
grep -E "segments|^\{" gpurun_out/r2a_n8_pb$pb.err gpurun_out/r2a_n8_pb$pb.log | cut -c1-600
done
