mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest8.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench8.err; cat gpurun_out/r02_bench8.json
ONLY_CAPS=1 CAPS="s1_k3_c1:116 s1_k11_c2:129 s2_k3_c2:136 s3_k11_c2:167" bash tools/gpu_ncu.sh r02c 3
