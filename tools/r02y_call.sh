#!/bin/bash
# Round 2, call y (1 GPU): ring depth of kernel 5 with the interpolation warps (DFR2D_WS_SPLIT=1), 2M triangles
tag=${1:-r02y}
o=gpurun_out
mkdir -p $o
for st in 0 2 3 4 5; do
  for n in 4 2; do
    echo -n "split=1 stages=$st N=$n: "
    DFR2D_WS_SPLIT=1 DFR2D_WS_STAGES=$st timeout 200 python tools/elem_knockout.py --nx 1000 --order $n --libs gocfd_b200/csrc/libdfr2d.so 2>/dev/null | head -1
  done
done | tee $o/${tag}_ring_depth_split.txt
exit 0
