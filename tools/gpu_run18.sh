mkdir -p gpurun_out
timeout 600 python tools/ab_bench.py --batch 32 --secs 10 --rounds 4 --calls 3 --settings "pdl=0;pdl=2;pdl=0,graph=1;pdl=2,graph=1" --json gpurun_out/r02_ab18_b32.json 2>&1 | tail -5
timeout 600 python tools/ab_bench.py --batch 1 --secs 3 --rounds 4 --calls 10 --settings "pdl=0,graph=1;pdl=2,graph=1;pdl=1,graph=1" --json gpurun_out/r02_ab18_b1.json 2>&1 | tail -4
