#!/bin/bash
# what the driver does at round end: bench at N=1,2,4,8 back to back on one 8-GPU box (+ multi-GPU parity)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; tail -2 gpurun_out/scale_n1.err | cut -c1-300
for n in 2 4 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; tail -2 gpurun_out/scale_n$n.err | cut -c1-300
done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=[json.loads(l) for l in open(f'gpurun_out/scale_n{n}.json') if l.startswith('{')][0]
        print(n, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'clocks', d['clocks'])
    except Exception as e: print(n, 'FAILED', e)
PY
