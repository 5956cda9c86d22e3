mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 300 > gpurun_out/r02_tc_check3.txt 2>&1; echo "tc_check rc=$?"
cut -c1-250 gpurun_out/r02_tc_check3.txt
WV=$(python - <<'PY'
import json
try:
    s = json.load(open("gpurun_out/tc_check_summary.json"))
    ok = [k.split(":")[1] for k, v in s.items() if k.startswith("f16x3") and v < 1e-4]
    print(ok[0] if ok else "none")
except Exception as e:
    print("none")
PY
)
echo "first passing wide variant: $WV"
if [ "$WV" != "none" ]; then
  export OVC_WIDE_VARIANT=$WV
  timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest_parity3.log 2>&1; echo "parity rc=$?"
  tail -15 gpurun_out/r02_pytest_parity3.log
  timeout 900 python -m pytest tests/test_gpu_tts.py -m gpu -q --timeout 300 -s > gpurun_out/r02_pytest_tts3.log 2>&1; echo "tts rc=$?"
  tail -15 gpurun_out/r02_pytest_tts3.log
  grep -h "long text\|tts_b\|benched" gpurun_out/r02_pytest_tts3.log gpurun_out/r02_pytest_parity3.log | head -20
fi
for wv in 0 1 2; do timeout 300 python tools/layer_report.py --precision f16x3 --wide-variant $wv --json gpurun_out/r02_layers3_wv$wv.json > gpurun_out/r02_layers3_wv$wv.txt 2>&1; echo "layers wv$wv rc=$?"; head -2 gpurun_out/r02_layers3_wv$wv.txt; done
timeout 300 python tools/layer_report.py --precision f16 --json gpurun_out/r02_layers3_f16.json > gpurun_out/r02_layers3_f16.txt 2>&1; head -1 gpurun_out/r02_layers3_f16.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench3.err; cat gpurun_out/r02_bench3.json
