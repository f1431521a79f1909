#!/bin/bash
# Round 2, call am (1 GPU): config C4 (NACA 0012 transonic, P=2, PerssonC0) towards its steady state on the device, residual
# history every 100 iterations, compared with the committed C-oracle history where both exist; the golden-history test
o=gpurun_out
mkdir -p $o
timeout 120 python tools/c4_convergence.py --impl device --iterations 60000 --every 100 --out $o/r02am_c4_device.json; echo "c4 rc=$?"
timeout 120 python -m pytest tests/test_c4_golden.py -q > $o/r02am_pytest_c4.log 2>&1; echo "pytest rc=$?"; tail -n 3 $o/r02am_pytest_c4.log
exit 0
