#!/bin/bash
# Round 2, call n (1 GPU): viscous edge flux fused into the edge kernels (default) vs k_visc_edge (DFR2D_EDGE_VISC_FUSED=0)
tag=${1:-r02n}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $o/${tag}_pytest.log
for fused in 1 0; do
for n in 4 3 2; do
  DFR2D_EDGE_VISC_FUSED=$fused timeout 200 python tools/grad_kernel_ab.py --order $n --variants 10 > $o/${tag}_ab_N${n}_fused$fused.json 2>> $o/${tag}_ab.err
  python -c "
import json,sys
d=json.load(open('$o/${tag}_ab_N${n}_fused$fused.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('fused=$fused N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
"
done
done
tail -3 $o/${tag}_ab.err
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err; echo "bench c3 rc=$?"
python -c "
import json
l=json.loads(open('$o/${tag}_bench_c3.json').read().strip().splitlines()[-1]); print('c3', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['phase_ms'])"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
