#!/bin/bash
# Round 2, call an (1 GPU): the C5 line at SURVEY.md 8(d)'s protocol (>= 50 timed steps after 5 warm-up steps)
o=gpurun_out
mkdir -p $o
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-also > $o/r02an_bench_c5_50steps.json 2> $o/r02an_bench.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open('$o/r02an_bench_c5_50steps.json').read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["whole_stage"]["frac"], l["e2e"]["value"], l["e2e"]["step_loop_value"], l["clocks"], l["checksum"]["l2"][0])
PY
exit 0
