mkdir -p gpurun_out
timeout 900 python tools/tc_check.py 300 > gpurun_out/r02_tc_check4.txt 2>&1; echo "tc_check rc=$?"
grep -v "per 128-step" gpurun_out/r02_tc_check4.txt | cut -c1-220 | grep "wide_variant\|ragged\|done\|rror"
SEL=$(python - <<'PY'
import json
try:
    s = json.load(open("gpurun_out/tc_check_summary.json"))
    ok = [k for k, v in s.items() if k.startswith("f16x3") and v < 1e-4]
    print(ok[0].split(":")[1] + " " + ok[0].split(":")[2] if ok else "none")
except Exception as e:
    print("none")
PY
)
echo "first passing (wide variant, act_tma): $SEL"
if [ "$SEL" != "none" ]; then
  export OVC_WIDE_VARIANT=$(echo $SEL | cut -d' ' -f1)
  export OVC_ACT_TMA=$(echo $SEL | cut -d' ' -f2)
  timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tts.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?"
  tail -8 gpurun_out/r02_pytest4.log
  unset OVC_WIDE_VARIANT OVC_ACT_TMA
fi
for cfg in "0 1" "0 0" "2 1" "2 0" "1 1"; do set -- $cfg; timeout 300 python tools/layer_report.py --precision f16x3 --wide-variant $1 --act-tma $2 --json gpurun_out/r02_layers4_wv$1_tma$2.json > gpurun_out/r02_layers4_wv$1_tma$2.txt 2>&1; echo "layers wv$1 tma$2 rc=$?"; head -1 gpurun_out/r02_layers4_wv$1_tma$2.txt; done
timeout 300 python tools/layer_report.py --precision f16 --json gpurun_out/r02_layers4_f16.json > gpurun_out/r02_layers4_f16.txt 2>&1; head -1 gpurun_out/r02_layers4_f16.txt
