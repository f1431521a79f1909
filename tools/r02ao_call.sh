#!/bin/bash
# Round 2, call ao (1 GPU): gradient plot kernel evaluating one direction per call (half the scratch): plot, window, C host and smoke
o=gpurun_out
mkdir -p $o
timeout 200 python -m pytest tests/test_plot_field.py tests/test_window.py tests/test_c_host.py tests/test_c4_golden.py -m gpu -q > $o/r02ao_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $o/r02ao_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $o/r02ao_smoke.log 2>&1; echo "smoke rc=$?"
exit 0
