mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_run1_gpu.txt
timeout 60 ./build_probe/tc_f16_test > gpurun_out/r02_probe_f16.txt 2>&1; echo "probe rc=$?"
cat gpurun_out/r02_probe_f16.txt
timeout 300 python tools/tc_check.py > gpurun_out/r02_tc_check.txt 2>&1; echo "tc_check rc=$?"
tail -45 gpurun_out/r02_tc_check.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest_parity1.log 2>&1; echo "parity rc=$?"
tail -30 gpurun_out/r02_pytest_parity1.log
timeout 900 python -m pytest tests/test_gpu_tts.py -m gpu -q --timeout 300 -s > gpurun_out/r02_pytest_tts1.log 2>&1; echo "tts rc=$?"
tail -30 gpurun_out/r02_pytest_tts1.log
for wv in 0 1 2; do timeout 300 python tools/layer_report.py --precision f16x3 --wide-variant $wv --json gpurun_out/r02_layers_wv$wv.json > gpurun_out/r02_layers_wv$wv.txt 2>&1; echo "layers wv$wv rc=$?"; head -3 gpurun_out/r02_layers_wv$wv.txt; done
timeout 300 python tools/layer_report.py --precision f16 --json gpurun_out/r02_layers_f16.json > gpurun_out/r02_layers_f16.txt 2>&1
head -2 gpurun_out/r02_layers_f16.txt
