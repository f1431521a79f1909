#!/bin/bash
# Round 2, call c (1 GPU): config tests as specified, the N=1 bench line (with C2/C3 under "also"), the reference arm on the
# full mesh, and an ncu --set full capture of k_elem_ws WITH source attribution (bank conflicts per SASS line).
tag=${1:-r02c}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests/test_configs_as_specified.py tests/test_zz_device_properties.py -m gpu -q -x --timeout 600 > $o/${tag}_pytest_configs.log 2>&1; echo "pytest rc=$?"
tail -8 $o/${tag}_pytest_configs.log
timeout 600 python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err; echo "bench rc=$?"; tail -3 $o/${tag}_bench_c5.err
cat $o/${tag}_bench_c5.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err; echo "ref rc=$?"; tail -3 $o/${tag}_bench_reference.err
cat $o/${tag}_bench_reference.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_elem_ws' -s 6 -c 2 -f -o $o/${tag}_elem_tma \
    python bench.py --nx 1000 --steps 2 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_ncu_elem.log 2>&1
tail -2 $o/${tag}_ncu_elem.log
