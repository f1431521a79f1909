#!/bin/bash
# First GPU call of round 2 (one B200, ~6 min): everything round 1 left unmeasured or gated.
#   1 full GPU suite incl. the gated cases (k_elem_mma_diss at N=1,3 / local dt, dfr2d_multi_step on one device)
#   2 A/B of the PerssonC0 kernels: DFMA gradient (1), k_grad_pipe (3), k_grad_pipe + k_elem_mma_diss (9) at N=4,3,2
#   3 bench lines c5 / c3 with the defaults
#   4 ncu --set full of k_elem_mma_diss and k_grad_pipe (500K elements)
# usage (under gpurun): bash tools/round2_first_call.sh <tag>
tag=${1:-r02a}
o=gpurun_out
mkdir -p $o
{ echo "go: $(command -v go || echo absent)"; go version 2>&1; ls -d /usr/local/go /usr/lib/go* 2>&1; echo "nsys: $(command -v nsys || echo absent)"; nproc; grep -m1 'model name' /proc/cpuinfo; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv; } > $o/${tag}_box_probe.txt 2>&1
cat $o/${tag}_box_probe.txt
# fragment layouts of mma.sync f64 (m8n8k4 as used everywhere, m16n8k8 as assumed for the next gradient kernel)
(cd tools/exp && { [ -x dmma_layout ] || nvcc -gencode arch=compute_100a,code=sm_100a -o dmma_layout dmma_layout.cu; } && ./dmma_layout) | tee $o/${tag}_dmma_layout.txt
DFR2D_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q > $o/${tag}_pytest_all.log 2>&1; echo "pytest rc=$?"
tail -5 $o/${tag}_pytest_all.log
for n in 4 3 2; do
  timeout 120 python tools/grad_kernel_ab.py --order $n --variants 1,3,9 > $o/${tag}_ab_N$n.json 2>> $o/${tag}_ab.err
  cat $o/${tag}_ab_N$n.json
done
timeout 300 python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err; cat $o/${tag}_bench_c5.json
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err; cat $o/${tag}_bench_c3.json
DFR2D_DISS_ELEM_KERNEL=3 timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_elem_mma_diss|k_grad_pipe' \
    -s 10 -c 2 -f -o $o/${tag}_diss_kernels python tools/grad_kernel_ab.py --nx 1000 --ny 250 --steps 1 --variants 9 \
    > $o/${tag}_ncu_diss.log 2>&1
tail -2 $o/${tag}_ncu_diss.log
