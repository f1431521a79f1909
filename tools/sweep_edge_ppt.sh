timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -1
DFR2D_EDGE_PPT=1 timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -1
echo "split=0"; DFR2D_EDGE_SPLIT=0 python bench.py --steps 6 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['edge_kernel'])"
for ppt in 6 3 2 1; do echo "split=1 ppt=$ppt"; DFR2D_EDGE_PPT=$ppt python bench.py --steps 6 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['edge_kernel'])"; done
for ppt in 4 2 1; do echo "c2 split=1 ppt=$ppt"; DFR2D_EDGE_PPT=$ppt python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['edge_kernel'])"; done
