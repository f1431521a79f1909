#!/bin/bash
# Round 2, call p (1 GPU): kernel 5 with the extra RK-register slabs fetched by the TMA engine (DFR2D_WS_TMA=1) vs cp.async
tag=${1:-r02p}
o=gpurun_out
mkdir -p $o
DFR2D_WS_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "vortex or naca or rhs_parity or final_time or multi_partition" > $o/${tag}_pytest_tma.log 2>&1; echo "pytest rc=$?"
tail -3 $o/${tag}_pytest_tma.log
for tma in 0 1 0 1; do
  DFR2D_WS_TMA=$tma timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_bench_c5_tma$tma.json 2> $o/${tag}_bench.err
  python -c "
import json
l=json.loads(open('$o/${tag}_bench_c5_tma$tma.json').read().strip().splitlines()[-1]); print('tma $tma', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['phase_ms']['element kernel'], l['checksum']['l2'][0])"
done
cuobjdump -sass gocfd_b200/csrc/libdfr2d.so 2>/dev/null | awk '/Function : _ZN5dfr2d9k_elem_wsILi4ELi8ELb0E/{f=1} f&&/Function : /&&!/k_elem_wsILi4ELi8ELb0E/{f=0} f' | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40 > $o/${tag}_sass_census_k_elem_ws.txt
grep -E "UTMALDG|LDGSTS|DMMA|SYNCS|UBLKCP" $o/${tag}_sass_census_k_elem_ws.txt
