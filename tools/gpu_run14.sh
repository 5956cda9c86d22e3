mkdir -p gpurun_out
bash tools/gpu_ncu.sh r02e 3 2>&1 | grep "rc="
timeout 120 ./build_probe/ffma_bench > gpurun_out/r02_ffma_bench.txt 2>&1; echo "ffma rc=$?"; cat gpurun_out/r02_ffma_bench.txt | tail -6
timeout 300 python tools/determinism_check.py > gpurun_out/r02_determinism_1.txt 2>&1; timeout 300 python tools/determinism_check.py > gpurun_out/r02_determinism_2.txt 2>&1; cat gpurun_out/r02_determinism_1.txt gpurun_out/r02_determinism_2.txt | grep -v Warning
timeout 900 compute-sanitizer --tool racecheck python tools/determinism_check.py 2>&1 | grep -v Warning | tail -6
