mkdir -p gpurun_out
bash tools/gpu_ncu.sh r02f 3 2>&1 | grep "rc="
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench23.json 2> gpurun_out/r02_bench23.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench23.err; cut -c1-600 gpurun_out/r02_bench23.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench23_ref.json 2> gpurun_out/r02_bench23_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_bench23_ref.json
