set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r3m_tests.log 2>&1; tail -4 gpurun_out/r3m_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3m_smoke.log 2>&1; echo "smoke rc=$?"
