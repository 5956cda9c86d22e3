mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "pipelined or graph or api_convert or full_size_batch" > gpurun_out/r02_pytest30.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02_pytest30.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cudnn --no-config3 > gpurun_out/r02_bench30.json 2> gpurun_out/r02_bench30.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench30.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench30.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d.get("e2e_convert_batch"), d["config4"], d["config"]["e2e_api"])
PY
