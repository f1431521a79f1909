#!/bin/bash
# Round profile on ONE B200: GPU parity tests, bench lines, ncu launch list + full capture of the two stage kernels.
# usage (under gpurun): bash tools/run_round_profile.sh <tag>
tag=${1:-r01e}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $o/${tag}_pytest.log
python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err; echo "bench c5 rc=$?"
python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline > $o/${tag}_bench_c2.json 2> $o/${tag}_bench_c2.err; echo "bench c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_c5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_elem_ws|k_edge_int' -s 30 -c 6 -f -o $o/${tag}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_full.log 2>&1
cat $o/${tag}_bench_c5.json
cat $o/${tag}_bench_c2.json
