#!/bin/bash
# Round 2, call aj (1 GPU): gradient plot fields (dfr2d_capture_edge_values / dfr2d_gradient_field) and the C host's
# thread-migration mode
o=gpurun_out
mkdir -p $o
timeout 400 python -m pytest tests/test_plot_field.py tests/test_c_host.py -m gpu -q > $o/r02aj_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -n 30 $o/r02aj_pytest_new.log
exit 0
