#!/bin/bash
# Round 2, call b (2 GPUs): the mailbox protocol on real devices -- GPU suite (multi-device multi_step, multi-process IPC),
# then the 2-GPU bench line with all three drivers.
tag=${1:-r02b}
o=gpurun_out
mkdir -p $o
nvidia-smi topo -m > $o/${tag}_topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $o/${tag}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > $o/${tag}_bench_n2.json 2> $o/${tag}_bench_n2.err; echo "bench n2 rc=$?"
tail -5 $o/${tag}_bench_n2.err
cat $o/${tag}_bench_n2.json
