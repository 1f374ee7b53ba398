#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py --steps 20 --warmup 3 > gpurun_out/configs_n1_final.jsonl 2> gpurun_out/configs_n1_final.err; tail -3 gpurun_out/configs_n1_final.err; cat gpurun_out/configs_n1_final.jsonl | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('|', d['case'][:70], '|', d['payload_GBps'], '|', d['ms_per_step'], '|', d.get('hbm_frac_of_measured_copy_peak'), '|')"
