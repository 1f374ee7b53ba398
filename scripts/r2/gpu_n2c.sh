#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "push" 2>&1 | tail -25 | tee gpurun_out/r2n2c_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs --no-e2e > gpurun_out/r2n2c_bench_n2.json 2> gpurun_out/r2n2c_bench_n2.err
tail -5 gpurun_out/r2n2c_bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2n2c_bench_n2.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("value", d["value"], d["fetch_mode"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"])
        print("pull", d["pull"] and (d["pull"]["value"], d["pull"]["ms_per_step"]))
        print("push", d["push"] and {k:v for k,v in d["push"].items() if k!="path"})
PY
