mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest32.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest32.log
