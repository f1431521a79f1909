#!/bin/bash
# Round 2, call m (1 GPU): k_grad_pipe with the ragged gradient rows by DFMA (MT 5 -> 4 at N=4); full GPU suite; c3 line
tag=${1:-r02m}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $o/${tag}_pytest.log
for n in 4 3 2; do
  timeout 200 python tools/grad_kernel_ab.py --order $n --variants 1,10 > $o/${tag}_ab_N$n.json 2>> $o/${tag}_ab.err
  python -c "
import json,sys
d=json.load(open('$o/${tag}_ab_N$n.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
    elif k.startswith('rel_l2'): print(k,v)
"
done
tail -3 $o/${tag}_ab.err
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err; echo "bench c3 rc=$?"
python -c "
import json
l=json.loads(open('$o/${tag}_bench_c3.json').read().strip().splitlines()[-1]); print('c3', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['phase_ms'])"
