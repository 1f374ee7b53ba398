#!/bin/bash
# round 2, 8-GPU call: multi-GPU parity tests (log kept), bench.py --gpus 8 (and 4), reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -v 2>&1 | tail -25 | tee gpurun_out/r2_gpu_multi_tests_n$NG.txt
for N in $NG 4; do
  [ "$N" -gt "$NG" ] && continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 --push > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
  tail -2 gpurun_out/r2_bench_n$N.err
  python - $N <<'PY'
import json,sys
N=sys.argv[1]
for l in open(f"gpurun_out/r2_bench_n{N}.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("N",N,"value", round(d["value"],1), d["fetch_mode"], "ms", round(d["ms_per_step"],4), "roof", round(d["roofline"]["frac"],3), "ver", d["verified_rows"], d["mismatches"], "owners", d["owners_hit"])
        print("  pull", d["pull"] and round(d["pull"]["value"],1), "push", d["push"] and (round(d["push"]["value"],1), round(d["push"]["ms_per_step"],4)))
        print("  e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k!="path"})
        for c in d["configs"]:
            print("  ", c["name"].ljust(28), "%8.1f" % c.get("value",0), "ms %.4f" % c.get("ms_per_step",0), "frac", round(c.get("roofline",{}).get("frac",0),3), "ver", c.get("verified_rows"), c.get("owners_hit"))
PY
done
