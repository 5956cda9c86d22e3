mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 300 > gpurun_out/r02_tc_check19.txt 2>&1; echo "tc_check rc=$?"
grep -v "per 128-step" gpurun_out/r02_tc_check19.txt | cut -c1-170 | grep "wide_variant\|ragged\|dec.stage\|done\|rror\|rap" | head -60
timeout 300 python tools/layer_report.py --precision f16x3 --pair 1 --json gpurun_out/r02_layers19_pair.json > gpurun_out/r02_layers19_pair.txt 2>&1; echo "layers rc=$?"; head -32 gpurun_out/r02_layers19_pair.txt
timeout 300 python tools/ab_bench.py --batch 32 --secs 10 --rounds 3 --calls 3 --settings "pair=0;pair=1" 2>&1 | tail -3
