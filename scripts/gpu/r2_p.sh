set -x
mkdir -p gpurun_out
NP=${NP:-2}
STG_HALO_MODE=ce timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 tests/dist_worker.py > gpurun_out/r2p_worker.log 2>&1; echo "worker rc=$?"; tail -3 gpurun_out/r2p_worker.log
run() { # name nproc env...
  name=$1; np=$2; shift 2
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $np --steps 20 --warmup 5 --no-extras > gpurun_out/r2p_$name.log 2> gpurun_out/r2p_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/r2p_$name.err | cut -c1-300
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2p_$name.log').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(l['ms_per_step'],4), 'value', round(l['value'],1), 'fwd_ms', round(l['roofline']['kernel_ms'],4), l['config']['halo']['visit_schedule_fwd'])
    for s in l['segments'][:2]: print(' ', s['rank'], s['fwd'])
except Exception as ex: print('parse fail', ex)
PY
}
run auto $NP A=1
run twopass $NP STG_FIRST_PHASE_UNITS=1e18
run half $NP STG_FIRST_PHASE_UNITS=8e6
