#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
echo "== PROCS vmm"; timeout 300 python scripts/probes/mix_probe.py procs 2>&1 | sort | grep -v "^$"
echo "== THREADS vmm"; timeout 300 python scripts/probes/mix_probe.py threads 2>&1 | sort | grep -v "^$"
echo "== PROCS legacy"; DDS_SHARD_ALLOC=legacy timeout 300 python scripts/probes/mix_probe.py procs 2>&1 | sort | grep "both=True"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_d_n2.json 2> gpurun_out/bench_d_n2.err; tail -5 gpurun_out/bench_d_n2.err; cat gpurun_out/bench_d_n2.json
