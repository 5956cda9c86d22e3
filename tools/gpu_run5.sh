mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python tools/tc_check.py 300 > gpurun_out/r02_tc_check5.txt 2>&1; echo "tc_check rc=$?"
grep -v "per 128-step" gpurun_out/r02_tc_check5.txt | cut -c1-200 | grep "wide_variant\|ragged\|done\|rror" | head -40
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_pytest5.log
for cfg in "0 1" "0 0" "2 1" "1 1"; do set -- $cfg; timeout 300 python tools/layer_report.py --precision f16x3 --wide-variant $1 --act-tma $2 --json gpurun_out/r02_layers5_wv$1_tma$2.json > gpurun_out/r02_layers5_wv$1_tma$2.txt 2>&1; echo "layers wv$1 tma$2 rc=$?"; head -40 gpurun_out/r02_layers5_wv$1_tma$2.txt; done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench5.err; cat gpurun_out/r02_bench5.json
