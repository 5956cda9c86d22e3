mkdir -p gpurun_out
timeout 600 python tools/tc_check.py 300 > gpurun_out/r02_tc_check20.txt 2>&1; echo "tc_check rc=$?"
grep "tc_check done\|rror\|rap" gpurun_out/r02_tc_check20.txt | cut -c1-400
timeout 300 python tools/layer_report.py --precision f16x3 --pair 1 --json gpurun_out/r02_layers20_pair.json > gpurun_out/r02_layers20_pair.txt 2>&1; echo "layers rc=$?"; head -3 gpurun_out/r02_layers20_pair.txt; grep "^T32\|^T64" gpurun_out/r02_layers20_pair.txt | head -20
timeout 300 python tools/layer_report.py --precision f16x3 --pair 0 --json gpurun_out/r02_layers20_nopair.json > gpurun_out/r02_layers20_nopair.txt 2>&1; head -1 gpurun_out/r02_layers20_nopair.txt; grep "^T32\|^T64" gpurun_out/r02_layers20_nopair.txt | head -20
timeout 300 python tools/ab_bench.py --batch 32 --secs 10 --rounds 3 --calls 3 --settings "pair=0;pair=1" 2>&1 | tail -3
