set -x
mkdir -p gpurun_out
for m in ce sm; do
STG_HALO_MODE=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_worker.py > gpurun_out/r2c_worker_$m.log 2>&1; echo "worker $m rc=$?"; tail -5 gpurun_out/r2c_worker_$m.log
done
for m in ce sm; do
STG_HALO_MODE=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c_n2_$m.log 2> gpurun_out/r2c_n2_$m.err; echo "bench $m rc=$?"; tail -3 gpurun_out/r2c_n2_$m.err
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2c_n2_$m.log').read().strip().splitlines()[-1])
    print('$m', l['ms_per_step'], l['value'], l['roofline']['kernel_ms'], l.get('e2e',{}).get('ms_per_step'))
    for s in l['segments']: print(s)
except Exception as ex: print('parse fail', ex)
PY
done
