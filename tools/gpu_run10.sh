mkdir -p gpurun_out
timeout 600 python tools/ab_bench.py --batch 32 --secs 10 --rounds 4 --calls 3 --settings "tune=0,pdl=0;tune=1,pdl=0;tune=2,pdl=0;tune=3,pdl=0;tune=3,pdl=1;tune=0,pdl=1;tune=3,pdl=1,graph=1" --json gpurun_out/r02_ab10_b32.json 2>&1 | tail -12
timeout 600 python tools/ab_bench.py --batch 1 --secs 3 --rounds 6 --calls 10 --settings "tune=3,pdl=0;tune=3,pdl=1;tune=3,pdl=0,graph=1;tune=3,pdl=1,graph=1" --json gpurun_out/r02_ab10_b1.json 2>&1 | tail -8
timeout 900 python tools/tc_check.py 300 > gpurun_out/r02_tc_check10.txt 2>&1; echo "tc_check rc=$?"
grep "tc_check done" gpurun_out/r02_tc_check10.txt | cut -c1-300
