#!/bin/bash
# Round 2, call ae (1 GPU): boundary-list edge kernel on a second stream beside the interior-edge kernel (single partition)
tag=${1:-r02ae}
o=gpurun_out
mkdir -p $o
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; rc=$?
echo "smoke rc=$rc"; tail -n 3 $o/${tag}_smoke.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 $o/${tag}_pytest.log
for ov in 0 1 0 1; do
  DFR2D_EDGE_OVERLAP=$ov timeout 200 python bench.py --workload c2 --steps 200 --warmup 5 --no-also --no-cpu-baseline > $o/${tag}_bench_c2_ov$ov.json 2> $o/${tag}_bench_c2_ov$ov.err
  python -c "
import json
l=json.loads(open('$o/${tag}_bench_c2_ov$ov.json').read().strip().splitlines()[-1])
print('c2 overlap=$ov', l['value'], l['ms_per_step'], l['e2e']['value'], l['checksum']['l2'][0])
"
done
DFR2D_EDGE_OVERLAP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err
python -c "
import json
l=json.loads(open('$o/${tag}_bench_c5.json').read().strip().splitlines()[-1])
print('c5', l['value'], l['ms_per_step'], 'elem', l['roofline']['avg_launch_ms'], l['roofline']['frac'], 'stage', l['roofline']['whole_stage']['frac'], l['checksum']['l2'][0], l['clocks']['sm_mhz'])
"
exit 0
