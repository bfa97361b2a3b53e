set -x
mkdir -p gpurun_out
STG_HALO_MODE=ce timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tests/dist_worker.py > gpurun_out/r2g_worker.log 2>&1; echo "worker rc=$?"; tail -3 gpurun_out/r2g_worker.log
run() { # name nproc env...
  name=$1; np=$2; shift 2
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $np --steps 10 --warmup 3 --no-extras > gpurun_out/r2g_$name.log 2> gpurun_out/r2g_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/r2g_$name.err | cut -c1-300
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2g_$name.log').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(l['ms_per_step'],4), 'value', round(l['value'],1), 'fwd_ms', round(l['roofline']['kernel_ms'],4))
    for s in l['segments'][:4]: print(' ', s['rank'], s['fwd'])
except Exception as ex: print('parse fail', ex)
PY
}
run n8_a 8 STG_COPY_STREAMS=3
run n8_b 8 STG_COPY_STREAMS=1
run n8_c 8 STG_COPY_STREAMS=3 STG_GATHER_FIRST=0
run n4_a 4 STG_COPY_STREAMS=3
run n4_b 4 STG_COPY_STREAMS=3 STG_GATHER_FIRST=1
run n2_a 2 STG_COPY_STREAMS=3
