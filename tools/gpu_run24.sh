mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest24.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest24.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6
