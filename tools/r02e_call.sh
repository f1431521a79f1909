#!/bin/bash
# Round 2, call e (1 GPU): full GPU suite after the padding cut of k_elem_ws + new plot fields, bench line.
tag=${1:-r02e}
o=gpurun_out
mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $o/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err; echo "bench rc=$?"; tail -3 $o/${tag}_bench_c5.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/TAG_bench_c5.json'.replace('TAG','r02e')).read().strip().splitlines()[-1])
print("c5", l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["phase_ms"], l["clocks"])
for k,v in l.get("also",{}).items(): print(k, v["value"], v["ms_per_step"], v["roofline_frac"])
PY
