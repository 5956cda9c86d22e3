mkdir -p gpurun_out
timeout 900 python tools/tc_check.py 300 > gpurun_out/r02_tc_check12.txt 2>&1; echo "tc_check rc=$?"
grep -v "per 128-step" gpurun_out/r02_tc_check12.txt | cut -c1-160 | grep "wide_variant 0 act_tma 1\|dec.pre\|done\|rror" | head -12
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest12.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench12.json 2> gpurun_out/r02_bench12.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench12.err; cut -c1-1200 gpurun_out/r02_bench12.json
timeout 900 python tools/sweep.py --json gpurun_out/r02_sweep12.json 2>&1 | tail -14
