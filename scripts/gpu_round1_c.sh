#!/bin/bash
# 2-GPU run: multi-GPU parity (peer access + CUDA IPC), torchrun bench at N=2, N=1 bench re-check
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_c_n1.json 2> gpurun_out/bench_c_n1.err; tail -3 gpurun_out/bench_c_n1.err; cat gpurun_out/bench_c_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_c_n2.json 2> gpurun_out/bench_c_n2.err; tail -5 gpurun_out/bench_c_n2.err; cat gpurun_out/bench_c_n2.json
