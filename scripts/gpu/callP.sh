set -x
mkdir -p gpurun_out
for v in "" _hint1 _hint2 _hint3 ""; do
STG_B200_LIB=$PWD/stgraph_b200/lib/libstgraph_b200$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('variant', '$v', 'ms_per_step', round(d['ms_per_step'],3), 'fwd_kernel_ms', round(d['roofline']['kernel_ms'],3))"
done > gpurun_out/hint_sweep.log 2>&1
cat gpurun_out/hint_sweep.log
