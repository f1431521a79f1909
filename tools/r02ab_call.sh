#!/bin/bash
# Round 2, call ab (1 GPU): k_grad_ws with the block loops of grad_mgroup unrolled (compile-time variants, 104 registers)
tag=${1:-r02ab}
o=gpurun_out
mkdir -p $o
for v in gd ge; do
  L=gocfd_b200/csrc/ko/libdfr2d_$v.so
  timeout 40 python tools/grad_kernel_ab.py --nx 400 --ny 100 --order 4 --steps 2 --variants 13 --lib $L > /dev/null 2>&1 || { echo "$v: small run failed or hung"; continue; }
  for n in 4 3; do
    timeout 90 python tools/grad_kernel_ab.py --order $n --variants 1,13 --lib $L > $o/${tag}_ab_${v}_N$n.json 2>> $o/${tag}_ab.err
    python -c "
import json
d=json.load(open('$o/${tag}_ab_${v}_N$n.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('$v N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
    elif k.startswith('rel_l2'): print(k,v)
" 2>/dev/null
  done
done
tail -n 3 $o/${tag}_ab.err
exit 0
