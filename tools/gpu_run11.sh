mkdir -p gpurun_out
timeout 600 python tools/ab_bench.py --batch 32 --secs 10 --rounds 3 --calls 3 --settings "tune=2,pdl=0,wide_variant=0;tune=6,pdl=0,wide_variant=0" --json gpurun_out/r02_ab11.json 2>&1 | tail -4
for t in 2 6; do timeout 300 python tools/layer_report.py --precision f16x3 --wide-variant 0 --tune $t --json gpurun_out/r02_layers11_t$t.json > gpurun_out/r02_layers11_t$t.txt 2>&1; head -1 gpurun_out/r02_layers11_t$t.txt; done
