set -x
mkdir -p gpurun_out
for thr in 512 1024 2048 4096; do
STG_HUB_THRESHOLD=$thr timeout 200 python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hub_threshold', $thr, 'ms_per_step', round(d['ms_per_step'],3), 'fwd_kernel_ms', round(d['roofline']['kernel_ms'],3))"
done > gpurun_out/hub_sweep.log 2>&1
cat gpurun_out/hub_sweep.log
