mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench27.json 2> gpurun_out/r02_bench27.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench27.err; cut -c1-400 gpurun_out/r02_bench27.json
timeout 900 python tools/sweep.py --json gpurun_out/r02_sweep27.json 2>&1 | tail -12
