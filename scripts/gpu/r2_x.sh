set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 100 python scripts/bench_gat.py > gpurun_out/r2x_$name.log 2>&1; echo "$name $(tail -1 gpurun_out/r2x_$name.log | cut -c1-110)"; }
run g4 STG_GAT_HUB_GRID=4
run g3 STG_GAT_HUB_GRID=3
run g2 STG_GAT_HUB_GRID=2
run g1 STG_GAT_HUB_GRID=1
run g6 STG_GAT_HUB_GRID=6
run g4b STG_GAT_HUB_GRID=4
run g2b STG_GAT_HUB_GRID=2
run g2_h64 STG_GAT_HUB_GRID=2 STG_GAT_HUB_THRESHOLD=64
run g2_c1 STG_GAT_HUB_GRID=2 STG_GAT_CHUNK=1
run g2_c4 STG_GAT_HUB_GRID=2 STG_GAT_CHUNK=4
