#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python scripts/bench_configs.py --steps 20 --warmup 3 > gpurun_out/configs_n1.jsonl 2> gpurun_out/configs_n1.err; tail -5 gpurun_out/configs_n1.err; cat gpurun_out/configs_n1.jsonl
