#!/bin/bash
# Round 2, call i (1 GPU): the edge-indexed viscous-normal arrays (vn) of the PerssonC0 path -- full GPU suite, then the A/B
# of the stage phases and the c3 bench line.
tag=${1:-r02i}
o=gpurun_out
mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $o/${tag}_pytest.log
for n in 4 3 2; do
  timeout 200 python tools/grad_kernel_ab.py --order $n --variants 1,3,9 > $o/${tag}_ab_N$n.json 2>> $o/${tag}_ab.err
  python -c "
import json,sys
d=json.load(open('$o/${tag}_ab_N$n.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
    elif k.startswith('rel_l2'): print(k,v)
"
done
DFR2D_PREFETCH_TILES=0 timeout 200 python tools/grad_kernel_ab.py --order 4 --variants 9 > $o/${tag}_ab_N4_noprefetch.json 2>> $o/${tag}_ab.err
python -c "
import json
d=json.load(open('$o/${tag}_ab_N4_noprefetch.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('N=4 no prefetch',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
"
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err; echo "bench c3 rc=$?"
python -c "
import json
l=json.loads(open('$o/${tag}_bench_c3.json').read().strip().splitlines()[-1]); print('c3', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['phase_ms'])"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_elem_mma_diss|k_grad_pipe|k_visc_edge' \
    -s 15 -c 3 -f -o $o/${tag}_diss_kernels python tools/grad_kernel_ab.py --nx 1000 --ny 250 --steps 1 --variants 9 \
    > $o/${tag}_ncu_diss.log 2>&1
tail -1 $o/${tag}_ncu_diss.log
