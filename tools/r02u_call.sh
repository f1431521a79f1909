#!/bin/bash
# Round 2, call u (1 GPU): k_grad_ws is the default gradient kernel.  ncu --set full with source attribution of k_grad_ws
# (500K triangles), accumulation groups {3,1} vs {2,2} for it, and the C3 bench line.
tag=${1:-r02u}
o=gpurun_out
mkdir -p $o
for n in 4 3; do
  timeout 200 python tools/grad_kernel_ab.py --order $n --variants 1,10,12,13 > $o/${tag}_ab_N$n.json 2>> $o/${tag}_ab.err
  python -c "
import json
d=json.load(open('$o/${tag}_ab_N$n.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'phase_ms_mean' in v: print('N=$n',k,round(v['ms_per_stage'],3),{a:round(b,3) for a,b in v['phase_ms_mean'].items()})
    elif k.startswith('rel_l2'): print(k,v)
"
done
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-also --no-cpu-baseline > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err; echo "bench c3 rc=$?"
python -c "
import json
l=json.loads(open('$o/${tag}_bench_c3.json').read().strip().splitlines()[-1])
print('c3', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['phase_ms'])
"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_grad_ws' -s 10 -c 2 -f -o /tmp/${tag}_grad \
    python tools/grad_kernel_ab.py --nx 1000 --ny 250 --steps 1 --variants 12 > $o/${tag}_ncu_grad.log 2>&1
python tools/ncu_summary.py /tmp/${tag}_grad.ncu-rep > $o/${tag}_ncu_full_k_grad_ws_500K.txt 2>&1
python tools/ncu_hot_lines.py /tmp/${tag}_grad.ncu-rep 1 45 > $o/${tag}_hot_lines_k_grad_ws.txt 2>&1
tail -3 $o/${tag}_ab.err $o/${tag}_bench_c3.err
