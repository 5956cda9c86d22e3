mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest26.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest26.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cudnn --no-config3 > gpurun_out/r02_bench26.json 2> gpurun_out/r02_bench26.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench26.err; cut -c1-400 gpurun_out/r02_bench26.json
timeout 600 python tools/sweep.py --points "1x1,1x3,1x10,8x3,32x10" --json gpurun_out/r02_sweep26.json 2>&1 | tail -6
