mkdir -p gpurun_out
timeout 900 python tools/tc_check.py 300 > gpurun_out/r02_tc_check7.txt 2>&1; echo "tc_check rc=$?"
grep -v "per 128-step" gpurun_out/r02_tc_check7.txt | cut -c1-200 | grep "wide_variant\|ragged\|done\|rror" | head -40
for wv in 0 2 3; do timeout 300 python tools/layer_report.py --precision f16x3 --wide-variant $wv --json gpurun_out/r02_layers7_wv$wv.json > gpurun_out/r02_layers7_wv$wv.txt 2>&1; echo "layers wv$wv rc=$?"; head -42 gpurun_out/r02_layers7_wv$wv.txt; done
timeout 300 python tools/layer_report.py --precision f16 --json gpurun_out/r02_layers7_f16.json > gpurun_out/r02_layers7_f16.txt 2>&1; head -3 gpurun_out/r02_layers7_f16.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench7.err; cat gpurun_out/r02_bench7.json
