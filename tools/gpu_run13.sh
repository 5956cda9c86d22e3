mkdir -p gpurun_out
bash tools/gpu_ncu.sh r02d 3 2>&1 | grep "rc="
timeout 120 ./build_probe/ffma_bench > gpurun_out/r02_ffma_bench.txt 2>&1; echo "ffma rc=$?"; cat gpurun_out/r02_ffma_bench.txt | tail -6
timeout 900 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck_smoke.log
timeout 1200 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck_smoke.log
