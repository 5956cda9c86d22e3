mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench15_n2.json 2> gpurun_out/r02_bench15_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r02_bench15_n2.err; cut -c1-900 gpurun_out/r02_bench15_n2.json
