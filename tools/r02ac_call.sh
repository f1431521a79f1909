#!/bin/bash
# Round 2, call ac (1 GPU): residual maxima reduced inside the rk 4 launch of kernel 5 (no residual register write) and the
# fresh register stored by the interpolation warps; smoke first, then the parity suite, then A/B against the previous build
tag=${1:-r02ac}
o=gpurun_out
mkdir -p $o
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; rc=$?
echo "smoke rc=$rc"; tail -n 4 $o/${tag}_smoke.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 $o/${tag}_pytest.log
for v in prev new prev new; do
  L=""; [ $v = prev ] && L=$PWD/gocfd_b200/csrc/ko/libdfr2d_prev.so
  DFR2D_LIB_PATH=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $o/${tag}_bench_c5_$v.json 2> $o/${tag}_bench_$v.err
  python -c "
import json
l=json.loads(open('$o/${tag}_bench_c5_$v.json').read().strip().splitlines()[-1])
print('$v', l['value'], l['ms_per_step'], 'elem', l['roofline']['avg_launch_ms'], l['roofline']['frac'], 'stage', l['roofline']['whole_stage']['frac'], l['checksum']['l2'][0], l['clocks']['sm_mhz'])
"
done
for v in prev new; do
  L=gocfd_b200/csrc/libdfr2d.so; [ $v = prev ] && L=gocfd_b200/csrc/ko/libdfr2d_prev.so
  echo -n "$v N=4 2M: "; timeout 200 python tools/elem_knockout.py --nx 1000 --libs $L 2>/dev/null | head -1
done
exit 0
