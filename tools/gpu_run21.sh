mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest21.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_pytest21.log
timeout 300 python tools/ab_bench.py --batch 32 --secs 10 --rounds 3 --calls 3 --settings "pair=0;pair=1" 2>&1 | tail -3
timeout 300 python tools/ab_bench.py --batch 1 --secs 3 --rounds 4 --calls 10 --settings "pair=0,graph=1;pair=1,graph=1" 2>&1 | tail -3
