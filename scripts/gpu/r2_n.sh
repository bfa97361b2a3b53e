set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n_n2.log 2> gpurun_out/r2n_n2.err; echo "rc=$?"; tail -2 gpurun_out/r2n_n2.err | cut -c1-300
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2n_n2.log').read().strip().splitlines()[-1])
    print('n2', 'ms/step', round(l['ms_per_step'],4), 'value', round(l['value'],1), 'fwd_ms', round(l['roofline']['kernel_ms'],4), 'frac', round(l['roofline']['frac'],4))
    print('  e2e', l.get('e2e'))
    print('  gcn', l.get('gcn_2layer_epoch'))
    for s in l['segments'][:3]: print(' ', s['rank'], s['fwd'])
except Exception as ex: print('parse fail', ex)
PY
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist.py -m gpu -q > gpurun_out/r2n_tests.log 2>&1; tail -4 gpurun_out/r2n_tests.log
