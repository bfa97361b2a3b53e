set -x
mkdir -p gpurun_out
timeout 300 python scripts/r3_gcn_epoch_prof.py gpurun_out/r3e_gcn_prof.json > gpurun_out/r3e_gcn_prof.log 2>&1; tail -3 gpurun_out/r3e_gcn_prof.log
