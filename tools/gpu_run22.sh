mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "pair or benched or golden or ragged or graph" > gpurun_out/r02_pytest22.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest22.log
timeout 300 python tools/layer_report.py --precision f16x3 --pair 1 --json gpurun_out/r02_layers22_pair.json > gpurun_out/r02_layers22_pair.txt 2>&1; echo "layers rc=$?"; head -1 gpurun_out/r02_layers22_pair.txt; grep "^P" gpurun_out/r02_layers22_pair.txt
timeout 300 python tools/ab_bench.py --batch 32 --secs 10 --rounds 3 --calls 3 --settings "pair=0;pair=1" 2>&1 | tail -3
