#!/bin/bash
# final check after the cleanup commit: full GPU suite, smoke, default bench line
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_final2_tests.txt 2>&1; tail -3 gpurun_out/r2_final2_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final2_smoke.txt 2>&1; tail -2 gpurun_out/r2_final2_smoke.txt
python bench.py > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final2_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('verified_rows'), len(d.get('configs',[])))
PY
