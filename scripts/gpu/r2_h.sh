set -x
mkdir -p gpurun_out
for c in 4 2 8; do
STG_AGG_CHUNK=$c timeout 200 python scripts/r2_halo_prof.py 8 > gpurun_out/r2h_p8_c$c.log 2>&1; grep -E "pass|rows" gpurun_out/r2h_p8_c$c.log
done
STG_AGG_CHUNK=4 timeout 200 python scripts/r2_halo_prof.py 2 > gpurun_out/r2h_p2_c4.log 2>&1; grep -E "pass|rows" gpurun_out/r2h_p2_c4.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2h_tests.log 2>&1; tail -8 gpurun_out/r2h_tests.log
