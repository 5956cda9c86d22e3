mkdir -p gpurun_out
timeout 600 python tools/layer_report.py --precision f16x3 --json gpurun_out/r02_layers9.json > gpurun_out/r02_layers9.txt 2>&1; echo "layers rc=$?"; head -44 gpurun_out/r02_layers9.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest9.log
