#!/bin/bash
# Round 2, call o (1 GPU): ring-depth sweep of kernel 5 after the round-2 consumer changes
tag=${1:-r02o}
o=gpurun_out
mkdir -p $o
for st in 0 2 3 4; do
  DFR2D_WS_STAGES=$st timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_bench_c5_s$st.json 2> $o/${tag}_bench.err
  python -c "
import json
l=json.loads(open('$o/${tag}_bench_c5_s$st.json').read().strip().splitlines()[-1]); print('stages $st', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['phase_ms']['element kernel'])"
done
