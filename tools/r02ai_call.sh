#!/bin/bash
# Round 2, call ai (1 GPU): the compiled C host (tests/test_c_host.py) first, then the whole GPU suite, smoke() and the
# default bench line of the same build.
o=gpurun_out
mkdir -p $o
timeout 300 python -m pytest tests/test_c_host.py -q -x > $o/r02ai_pytest_c_host.log 2>&1; echo "c_host rc=$?"; tail -n 4 $o/r02ai_pytest_c_host.log
timeout 500 python -m pytest tests -m gpu -q -x > $o/r02ai_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $o/r02ai_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $o/r02ai_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --steps 20 --warmup 3 > $o/r02ai_bench_c5.json 2> $o/r02ai_bench_c5.err; echo "bench rc=$?"
cut -c1-900 $o/r02ai_bench_c5.json
exit 0
