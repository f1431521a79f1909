#!/bin/bash
# Round 2, call s (1 GPU): k_grad_pipe accumulation groups after the m-tile cut (MT = 4 at N=4: {3,1} (variant 10) vs {2,2} (11))
tag=${1:-r02s}
o=gpurun_out
mkdir -p $o
for rep in 1 2; do
for n in 4 3; do
  timeout 200 python tools/grad_kernel_ab.py --order $n --variants 10,11 > $o/${tag}_ab_N${n}_rep$rep.json 2>> $o/${tag}_ab.err
  python -c "
import json,sys
d=json.load(open('$o/${tag}_ab_N${n}_rep$rep.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
"
done
done
