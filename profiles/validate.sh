#!/bin/bash
# Full single-GPU validation under gpurun: `gpurun --timeout 900 -- "bash profiles/validate.sh <tag>"` runs every GPU test, smoke() and the
# default bench line (outputs in gpurun_out/<tag>_*).  Used for r2ab / r2ac / r2ae.
cd $GRAFT_REPO_ROOT
T=${1:-x}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 300 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k: round(x,3) for k,x in d['roofline']['stage_ms'].items()}, d['roofline']['frac'], d['clocks'], d['exit_agreement']['all_docs'], d['padded_variant']['value'])
PY
