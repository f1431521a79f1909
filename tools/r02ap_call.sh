#!/bin/bash
# Round 2, call ap (1 GPU): dfr2d_multi_create / multi_residual / multi_destroy through the C host; smoke
o=gpurun_out
mkdir -p $o
timeout 150 python -m pytest tests/test_c_host.py tests/test_abi_and_partition.py -q > $o/r02ap_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $o/r02ap_pytest.log
exit 0
