#!/bin/bash
# what the driver runs at round end on one GPU: gpu tests, smoke, reference arm, our arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -1 gpurun_out/final_ref.err | cut -c1-200; cut -c1-400 gpurun_out/final_ref.json
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/final_ours.json 2> gpurun_out/final_ours.err; tail -1 gpurun_out/final_ours.err | cut -c1-200
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/final_ours.json') if l.startswith('{')][0]
print('value',d['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches']); print('roofline',d['roofline']['frac'],d['roofline']['per_launch_ms'],d['roofline']['per_launch_event_pair_ms'],d['roofline']['traffic']); print('e2e',d['e2e']); print('cpu',d['cpu_baseline']); print('clocks',d['clocks'])
PY
timeout 300 python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | cut -c1-200 | tail -2
