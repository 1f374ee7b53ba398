#!/bin/bash
# first GPU contact: smoke, parity tests, short bench, launch list. Every step under its own timeout.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -3
nproc; free -g | head -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_a.json 2>&1; cat gpurun_out/bench_ref_a.json
