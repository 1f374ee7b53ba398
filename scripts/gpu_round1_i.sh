#!/bin/bash
# 8-GPU box: multi-GPU parity at 8 ranks, bench at N=8 and N=4, other configs at N=8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_i_n$n.json 2> gpurun_out/bench_i_n$n.err; tail -2 gpurun_out/bench_i_n$n.err | cut -c1-300; cat gpurun_out/bench_i_n$n.json | cut -c1-2500
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 scripts/bench_configs.py --cases cfg3,cfg4,cfg5 --steps 10 --warmup 3 > gpurun_out/configs_n8.jsonl 2> gpurun_out/configs_n8.err; tail -2 gpurun_out/configs_n8.err | cut -c1-300; cat gpurun_out/configs_n8.jsonl
