#!/bin/bash
# Round profile on ONE B200: GPU parity tests, the default bench line and the reference arm, the ncu launch list of the
# same command and ncu --set full captures of the stage kernels (C5 at full size for the DRAM traffic of kernel 5, the
# PerssonC0 kernels at 500K).  usage (under gpurun): bash tools/run_round_profile.sh <tag>
tag=${1:-r02z}
o=gpurun_out
mkdir -p $o
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
  tail -3 $o/${tag}_pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err; echo "bench c5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_c5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_elem_ws|k_edge_int' -s 30 -c 10 -f -o /tmp/${tag}_c5_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_elem_ws|k_grad_ws|k_edge_int|k_diss_prepare' \
    -s 24 -c 4 -f -o /tmp/${tag}_diss_full python tools/grad_kernel_ab.py --nx 1000 --ny 250 --steps 1 --variants 13 > $o/${tag}_ncu_diss.log 2>&1
# the reports stay on the box (an 8M-element capture with source is > 64 MiB); only text summaries travel
python tools/ncu_summary.py /tmp/${tag}_c5_full.ncu-rep > $o/${tag}_ncu_full_c5.txt 2>&1
python tools/ncu_summary.py /tmp/${tag}_diss_full.ncu-rep > $o/${tag}_ncu_full_diss_500K.txt 2>&1
python tools/ncu_hot_lines.py /tmp/${tag}_c5_full.ncu-rep 3 30 > $o/${tag}_hot_lines_k_elem_ws.txt 2>&1
python - <<PY2
import json, re
t=open('$o/${tag}_ncu_full_c5.txt').read()
out={}
blocks=t.split('=== ')[1:]
el=[b for b in blocks if 'k_elem_ws' in b]
def gb(b,key):
    m=re.search(key+r'\\s+([0-9.]+) (G|M)byte', b)
    return float(m.group(1))*(1e9 if m.group(2)=='G' else 1e6) if m else 0.0
tr=[gb(b,'dram__bytes_read.sum')+gb(b,'dram__bytes_write.sum') for b in el]
out['k_elem_N4']=sum(tr)/len(tr) if tr else None
out['per_launch']=tr
out['source']='profiles/${tag}_ncu_full_c5.txt (mean over the captured k_elem_ws<4,8,false> launches, 8M elements)'
json.dump(out, open('$o/${tag}_traffic.json','w'))
print(out)
PY2
python - <<PY
import json
l=json.loads(open('$o/${tag}_bench_c5.json').read().strip().splitlines()[-1])
print("c5", l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["whole_stage"]["frac"], l["e2e"], l["clocks"])
for k,v in l.get("also",{}).items(): print(k, v["value"], v["ms_per_step"], v["roofline_frac"])
print(l.get("cpu_baseline",{}).get("value"))
PY
