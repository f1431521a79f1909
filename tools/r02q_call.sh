#!/bin/bash
# Round 2, call q (N GPUs): final multi-GPU validation -- GPU suite on the N devices (spread multi_step, IPC processes), bench line
n=${1:-2}
tag=${2:-r02q}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $o/${tag}_pytest_n$n.log 2>&1; echo "pytest rc=$?"
tail -4 $o/${tag}_pytest_n$n.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 3 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err; echo "bench n$n rc=$?"
tail -3 $o/${tag}_bench_n$n.err
python - <<PY
import json
l=json.loads(open('$o/${tag}_bench_n$n.json').read().strip().splitlines()[-1])
print("peer", l["value"], l["ms_per_step"], "e2e", l["e2e"]["value"], l["e2e"].get("step_loop_value"), l["driver"], l["checksum"])
print("nccl", l.get("nccl_driver")); ms=l["multi_step"]; print("multi", {k:v for k,v in ms.items() if k not in("timeline_ms","what")})
for k,v in l["also"].items(): print(k, v["value"], v["ms_per_step"], v["e2e"], v["roofline_frac"], v["checksum"])
print(l["config"]["setup_s"], l["wall_s"])
PY
