mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench31.json 2> gpurun_out/r02_bench31.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench31.err; cut -c1-300 gpurun_out/r02_bench31.json
