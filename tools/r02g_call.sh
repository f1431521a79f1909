#!/bin/bash
# Round 2, call g (2 GPUs): window-mode partitions (dfr2d_create_window) in the tests and in the bench line; boundary-list
# kernel with one point per thread.
tag=${1:-r02g}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $o/${tag}_pytest.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > $o/${tag}_bench_n2.json 2> $o/${tag}_bench_n2.err; echo "bench n2 rc=$?"
tail -3 $o/${tag}_bench_n2.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02g_bench_n2.json').read().strip().splitlines()[-1])
print("peer", l["value"], l["ms_per_step"], "e2e", l["e2e"]["value"], l["driver"], l["checksum"])
print("phases", l["roofline"]["phase_ms"]); print("setup", l["config"]["setup_s"], "wall", l["wall_s"])
print("nccl", l.get("nccl_driver")); ms=l["multi_step"]; print("multi", {k:v for k,v in ms.items() if k not in("timeline_ms","what")})
for k,v in l["also"].items(): print(k, v["value"], v["ms_per_step"], v["e2e"], v["roofline_frac"], v["checksum"])
PY
