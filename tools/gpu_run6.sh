mkdir -p gpurun_out
timeout 60 ./build_probe/tmem_ld_probe > gpurun_out/r02_tmem_probe.txt 2>&1; echo "probe rc=$?"; tail -40 gpurun_out/r02_tmem_probe.txt
timeout 900 python tools/tc_check.py 300 > gpurun_out/r02_tc_check6.txt 2>&1; echo "tc_check rc=$?"
grep -v "per 128-step" gpurun_out/r02_tc_check6.txt | cut -c1-200 | grep "wide_variant\|ragged\|done\|rror" | head -40
for cfg in "0 1" "2 1"; do set -- $cfg; timeout 300 python tools/layer_report.py --precision f16x3 --wide-variant $1 --act-tma $2 --json gpurun_out/r02_layers6_wv$1_tma$2.json > gpurun_out/r02_layers6_wv$1_tma$2.txt 2>&1; echo "layers wv$1 tma$2 rc=$?"; head -42 gpurun_out/r02_layers6_wv$1_tma$2.txt; done
timeout 300 python tools/layer_report.py --precision f16 --json gpurun_out/r02_layers6_f16.json > gpurun_out/r02_layers6_f16.txt 2>&1; head -3 gpurun_out/r02_layers6_f16.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest6.log
