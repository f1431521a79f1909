#!/bin/bash
# Round 2, call x (1 GPU): kernel 5 with the interpolation warps (DFR2D_WS_SPLIT=1): smoke under a short timeout first (a
# broken barrier protocol would spin), then the GPU parity suite, then C5 with and without.
tag=${1:-r02x}
o=gpurun_out
mkdir -p $o
DFR2D_WS_SPLIT=1 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke_split.log 2>&1; rc=$?
echo "smoke split rc=$rc"; tail -n 4 $o/${tag}_smoke_split.log
if [ $rc -ne 0 ]; then exit 0; fi
DFR2D_WS_SPLIT=1 timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $o/${tag}_pytest_split.log 2>&1; echo "pytest split rc=$?"
tail -n 4 $o/${tag}_pytest_split.log
for sp in 0 1 0 1; do
  DFR2D_WS_SPLIT=$sp timeout 300 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $o/${tag}_bench_c5_split$sp.json 2> $o/${tag}_bench_split$sp.err
  python -c "
import json
l=json.loads(open('$o/${tag}_bench_c5_split$sp.json').read().strip().splitlines()[-1])
print('split=$sp', l['value'], l['ms_per_step'], 'elem', l['roofline']['avg_launch_ms'], l['roofline']['frac'], 'stage', l['roofline']['whole_stage']['frac'], l['checksum']['l2'][0], l['clocks']['sm_mhz'])
"
done
for sp in 0 1; do
  DFR2D_WS_SPLIT=$sp timeout 200 python tools/elem_knockout.py --nx 1000 --libs gocfd_b200/csrc/libdfr2d.so 2>/dev/null | head -1
  DFR2D_WS_SPLIT=$sp timeout 200 python tools/elem_knockout.py --nx 1000 --order 2 --libs gocfd_b200/csrc/libdfr2d.so 2>/dev/null | head -1
done
exit 0
