#!/bin/bash
# Round 2, call t (1 GPU): warp-specialised gradient kernel k_grad_ws (DFR2D_GRAD_KERNEL=4, variant 12) vs k_grad_pipe (10)
tag=${1:-r02t}
o=gpurun_out
mkdir -p $o
DFR2D_GRAD_KERNEL=4 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_device_properties.py tests/test_configs_as_specified.py -m gpu -q -x --timeout 600 -k "diss or dissipation or naca or c3 or c4 or scattered or multi_step" > $o/${tag}_pytest_gk4.log 2>&1; echo "pytest rc=$?"
tail -4 $o/${tag}_pytest_gk4.log
for n in 4 3 2; do
  timeout 200 python tools/grad_kernel_ab.py --order $n --variants 1,10,12 > $o/${tag}_ab_N$n.json 2>> $o/${tag}_ab.err
  python -c "
import json,sys
d=json.load(open('$o/${tag}_ab_N$n.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
    elif k.startswith('rel_l2'): print(k,v)
"
done
tail -3 $o/${tag}_ab.err
