set -x
mkdir -p gpurun_out
run() { # name nproc env...
  name=$1; np=$2; shift 2
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $np --steps 20 --warmup 5 --no-extras > gpurun_out/r2o_$name.log 2> gpurun_out/r2o_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/r2o_$name.err | cut -c1-300
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2o_$name.log').read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(l['ms_per_step'],4), 'value', round(l['value'],1), 'fwd_ms', round(l['roofline']['kernel_ms'],4))
    for s in l['segments'][:3]: print(' ', s['rank'], s['fwd'])
except Exception as ex: print('parse fail', ex)
PY
}
run n8_g2 8 STG_HALO_GROUPS=2
run n8_g1 8 STG_HALO_GROUPS=1
run n8_g3 8 STG_HALO_GROUPS=3
