#!/bin/bash
# Round 2, call d (N GPUs): the scaling bench line with all three drivers (peer = default, nccl comparison, multi_step).
n=${1:-8}
tag=${2:-r02d}
o=gpurun_out
mkdir -p $o
nvidia-smi topo -m > $o/${tag}_topo_n$n.txt 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 3 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err; echo "bench n$n rc=$?"
tail -5 $o/${tag}_bench_n$n.err
cat $o/${tag}_bench_n$n.json
