set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gat_rows_pipe_kernel|gat_hub_kernel" -s 6 -c 6 -o gpurun_out/gat_full -f python scripts/prof_gat.py 2 > gpurun_out/ncu_gat_full.log 2>&1
tail -2 gpurun_out/ncu_gat_full.log
