#!/bin/bash
# Round 2, call aa (1 GPU): GPU parity suite with the interpolation warps as default; k_grad_ws unroll / register variants
tag=${1:-r02aa}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 $o/${tag}_pytest.log
for v in main ga gb gc; do
  L=gocfd_b200/csrc/ko/libdfr2d_$v.so; [ $v = main ] && L=gocfd_b200/csrc/libdfr2d.so
  for n in 4 3; do
    timeout 200 python tools/grad_kernel_ab.py --order $n --variants 12,13 --lib $L > $o/${tag}_ab_${v}_N$n.json 2>> $o/${tag}_ab.err
    python -c "
import json
d=json.load(open('$o/${tag}_ab_${v}_N$n.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('$v N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
"
  done
done
tail -n 3 $o/${tag}_ab.err
exit 0
