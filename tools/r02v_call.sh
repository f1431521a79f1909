#!/bin/bash
# Round 2, call v (1 GPU): (1) tools/exp/pipemix: DMMA mixed with LDS / DFMA on the shared FP64 pipe; (2) kernel 5 with parts of
# the consumer work knocked out (tools/elem_knockout.py); (3) k_grad_ws accumulation groups at N=2.
tag=${1:-r02v}
o=gpurun_out
mkdir -p $o
( cd tools/exp && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipemix pipemix.cu && /tmp/pipemix ) > $o/${tag}_pipemix.txt 2>&1
cat $o/${tag}_pipemix.txt
timeout 300 python tools/elem_knockout.py --nx 1000 > $o/${tag}_knockout_N4.json 2> $o/${tag}_knockout.err
grep -v '^{"K"' $o/${tag}_knockout_N4.json
timeout 200 python tools/grad_kernel_ab.py --order 2 --variants 1,12,13 > $o/${tag}_ab_N2.json 2>> $o/${tag}_ab.err
python -c "
import json
d=json.load(open('$o/${tag}_ab_N2.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('N=2',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
    elif k.startswith('rel_l2'): print(k,v)
"
tail -n 3 $o/${tag}_knockout.err
tail -n 3 $o/${tag}_ab.err
exit 0
