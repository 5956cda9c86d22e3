mkdir -p gpurun_out
timeout 600 python tools/layer_report.py --precision f16x3 --json gpurun_out/r02_layers16.json > gpurun_out/r02_layers16.txt 2>&1; echo "layers rc=$?"; head -30 gpurun_out/r02_layers16.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest16.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest16.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench16.json 2> gpurun_out/r02_bench16.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench16.err; cut -c1-700 gpurun_out/r02_bench16.json
