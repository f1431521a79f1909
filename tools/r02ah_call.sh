#!/bin/bash
# Round 2, call ah (N GPUs): bench line of the final build (one process per GPU over IPC mailboxes + comparison drivers)
n=${1:-8}
tag=${2:-r02ah}
o=gpurun_out
mkdir -p $o
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 3 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err; echo "bench n$n rc=$?"
tail -n 3 $o/${tag}_bench_n$n.err
python - <<PY
import json
l=json.loads(open('$o/${tag}_bench_n$n.json').read().strip().splitlines()[-1])
print("peer", l["value"], l["ms_per_step"], "e2e", l["e2e"]["value"], l["e2e"].get("step_loop_value"), l["driver"], l["checksum"])
print("roofline", l["roofline"]["frac"], l["roofline"]["per_gpu_frac_min_max"], l["roofline"]["whole_stage"])
print("nccl", l.get("nccl_driver")); ms=l["multi_step"]; print("multi", {k:v for k,v in ms.items() if k not in("what",)})
for k,v in l["also"].items(): print(k, v["value"], v["ms_per_step"], v["e2e"], v["roofline_frac"], v["checksum"])
print(l["config"]["setup_s"], l["wall_s"])
PY
exit 0
