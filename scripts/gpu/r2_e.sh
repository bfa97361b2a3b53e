set -x
mkdir -p gpurun_out
timeout 200 python scripts/r2_halo_prof.py 8 > gpurun_out/r2e_halo.log 2>&1; tail -3 gpurun_out/r2e_halo.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_rows_pipe -s 4 -c 1 -o gpurun_out/r2e_halo python scripts/r2_halo_prof.py 8 > gpurun_out/r2e_ncu.log 2>&1; tail -3 gpurun_out/r2e_ncu.log
ls -la gpurun_out/r2e_halo.ncu-rep
