#!/bin/bash
# round 2, 2-GPU call: multi-GPU parity tests, push-vs-pull probe, NVLink counters, bench.py --gpus 2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8 | tee gpurun_out/r2n2_topo.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_dataset.py tests/test_bindings.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2n2_pytest.txt
timeout 300 scripts/probes/p2p_probe 2>&1 | tee gpurun_out/r2n2_p2p_probe.txt
for MODE in B Br A Ar; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 --no-python \
     bash scripts/r2/nvlink_ncu_wrap.sh $MODE gpurun_out/r2n2_nvlink_$MODE.csv 2>&1 | tail -3
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2_bench_n2.json 2> gpurun_out/r2n2_bench_n2.err
tail -3 gpurun_out/r2n2_bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2n2_bench_n2.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], d["ms_per_step_p10_p50_p90"], "roof", d["roofline"]["frac"], "ver", d["verified_rows"], d["mismatches"], "owners", d["owners_hit"])
        print("e2e", {k:v for k,v in d["e2e"].items() if k!="path"})
        for c in d["configs"]:
            print("  ", c["name"].ljust(28), "%8.1f" % c.get("value",0), "ms %.4f" % c.get("ms_per_step",0), "frac", round(c.get("roofline",{}).get("frac",0),3), "ver", c.get("verified_rows"), c.get("owners_hit"))
PY
