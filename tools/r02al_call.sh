#!/bin/bash
# Round 2, call al (2 GPUs): the whole GPU suite of the final build on two devices (un-skips the peer-GPU tests; the C host
# spreads its partitions over both)
o=gpurun_out
mkdir -p $o
timeout 500 python -m pytest tests -m gpu -q > $o/r02al_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 $o/r02al_pytest_2gpu.log
exit 0
