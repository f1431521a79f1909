#!/bin/bash
# Round 2, call l (1 GPU): the PerssonC0 element kernel on the warp-specialised ring (k_elem_ws<N,8,true>, DFR2D_DISS_ELEM_KERNEL=5)
tag=${1:-r02l}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_device_properties.py -m gpu -q -x --timeout 600 -k "diss or dissipation or naca_front" > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $o/${tag}_pytest.log
for n in 4 3 2; do
  timeout 200 python tools/grad_kernel_ab.py --order $n --variants 3,9,10 > $o/${tag}_ab_N$n.json 2>> $o/${tag}_ab.err
  python -c "
import json,sys
d=json.load(open('$o/${tag}_ab_N$n.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
    elif k.startswith('rel_l2'): print(k,v)
"
done
tail -3 $o/${tag}_ab.err
for st in 2 3; do
DFR2D_WS_STAGES=$st DFR2D_DISS_ELEM_KERNEL=5 timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_bench_c3_k5_s$st.json 2> $o/${tag}_bench_c3.err; echo "bench c3 rc=$?"
python -c "
import json
l=json.loads(open('$o/${tag}_bench_c3_k5_s$st.json').read().strip().splitlines()[-1]); print('c3 k5 stages $st', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['phase_ms'])"
done
