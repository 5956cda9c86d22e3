mkdir -p gpurun_out
timeout 600 python tools/ab_bench.py --batch 1 --secs 3 --rounds 5 --calls 10 --settings "graph=1,branches=0;graph=1,branches=1;graph=0,branches=0;graph=0,branches=1" --json gpurun_out/r02_ab17_b1.json 2>&1 | tail -5
timeout 600 python tools/ab_bench.py --batch 1 --secs 10 --rounds 4 --calls 10 --settings "graph=1,branches=0;graph=1,branches=1" --json gpurun_out/r02_ab17_b1_10s.json 2>&1 | tail -3
timeout 600 python tools/ab_bench.py --batch 4 --secs 3 --rounds 4 --calls 10 --settings "graph=1,branches=0;graph=1,branches=1" --json gpurun_out/r02_ab17_b4.json 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest17.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest17.log
