# What a GPU session of this repo runs (under gpurun, one B200): the -m gpu tests, the smoke check, the default bench,
# and -- with NCU=1 -- the ncu evidence (launch list, DRAM pass at the benched batch, --set full captures; tools/gpu_ncu.sh).
#   gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'        then, here:  bash tools/export_profiles.sh <tag>
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_n1.json
if [ -n "$NCU" ]; then bash tools/gpu_ncu.sh ${TAG:-r02x} 3 2>&1 | grep "rc="; fi
