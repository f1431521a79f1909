#!/bin/bash
# Round profile on ONE B200: GPU parity tests, the default bench line and the reference arm, the ncu launch list of the
# same command and ncu --set full captures of the stage kernels (C5 at full size for the DRAM traffic of kernel 5, the
# PerssonC0 kernels at 500K).  usage (under gpurun): bash tools/run_round_profile.sh <tag>
tag=${1:-r02z}
o=gpurun_out
mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $o/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err; echo "bench c5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_c5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_elem_ws|k_edge_int' -s 30 -c 10 -f -o $o/${tag}_c5_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_elem_mma_diss|k_grad_pipe|k_visc_edge' \
    -s 15 -c 3 -f -o $o/${tag}_diss_full python tools/grad_kernel_ab.py --nx 1000 --ny 250 --steps 1 --variants 9 > $o/${tag}_ncu_diss.log 2>&1
cuobjdump -sass -fun '_ZN5dfr2d9k_elem_wsILi4ELi8EEEvNS_10ElemWsArgsE' gocfd_b200/csrc/libdfr2d.so 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40 > $o/${tag}_sass_census_k_elem_ws.txt
python - <<PY
import json
l=json.loads(open('$o/${tag}_bench_c5.json').read().strip().splitlines()[-1])
print("c5", l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["whole_stage"]["frac"], l["e2e"], l["clocks"])
for k,v in l.get("also",{}).items(): print(k, v["value"], v["ms_per_step"], v["roofline_frac"])
print(l.get("cpu_baseline",{}).get("value"))
PY
