#!/bin/bash
# Round 2, call aq (1 GPU, last seconds of the budget): smoke() and the C-host tests on the final binary
o=gpurun_out
mkdir -p $o
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" > $o/r02aq_smoke.log 2>&1; echo "smoke rc=$?"
timeout 25 python -m pytest tests/test_c_host.py -q > $o/r02aq_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $o/r02aq_pytest.log
exit 0
