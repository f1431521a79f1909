#!/bin/bash
# Round 2, call ag (1 GPU): warp-specialised interior-edge kernel k_edge_ws (DFR2D_EDGE_WS=1)
tag=${1:-r02ag}
o=gpurun_out
mkdir -p $o
DFR2D_EDGE_WS=1 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; rc=$?
echo "smoke rc=$rc"; tail -n 3 $o/${tag}_smoke.log
if [ $rc -ne 0 ]; then exit 0; fi
DFR2D_EDGE_WS=1 timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 $o/${tag}_pytest.log
for cfg in "0 0" "1 0" "1 2" "1 3" "1 4"; do
  set -- $cfg
  echo -n "edge_ws=$1 stages=$2 N=4 2M: "; DFR2D_EDGE_WS=$1 DFR2D_EDGE_WS_STAGES=$2 timeout 100 python tools/elem_knockout.py --nx 1000 --libs gocfd_b200/csrc/libdfr2d.so 2>/dev/null | head -1
done
for cfg in "0 0" "1 0" "1 5"; do
  set -- $cfg
  echo -n "edge_ws=$1 stages=$2 N=2 2M: "; DFR2D_EDGE_WS=$1 DFR2D_EDGE_WS_STAGES=$2 timeout 100 python tools/elem_knockout.py --nx 1000 --order 2 --libs gocfd_b200/csrc/libdfr2d.so 2>/dev/null | head -1
done
for ws in 0 1; do
  DFR2D_EDGE_WS=$ws timeout 300 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $o/${tag}_bench_c5_ws$ws.json 2> $o/${tag}_bench_c5_ws$ws.err
  python -c "
import json
l=json.loads(open('$o/${tag}_bench_c5_ws$ws.json').read().strip().splitlines()[-1])
print('c5 edge_ws=$ws', l['value'], l['ms_per_step'], 'edge', l['roofline']['edge_kernel']['avg_launch_ms'], 'elem', l['roofline']['avg_launch_ms'], 'stage', l['roofline']['whole_stage']['frac'], l['checksum']['l2'][0], l['clocks']['sm_mhz'])
"
done
exit 0
