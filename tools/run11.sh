export MP_BENCH_VERBOSE=1
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nvidia-smi topo -m | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; grep "loop A" gpurun_out/bench_n2.err
MP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/trace_n2.err; grep mp_trace gpurun_out/trace_n2.err | tail -40
