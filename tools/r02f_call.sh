#!/bin/bash
# Round 2, call f (1 GPU): conflict-free epilogue row order (now default) and the 12-consumer-warp variant (DFR2D_WS_CW=12)
tag=${1:-r02f}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_plot_field.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "plot or vortex or rhs_parity or naca" > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $o/${tag}_pytest.log
for cw in 8 12; do
  DFR2D_WS_CW=$cw timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_bench_c5_cw$cw.json 2> $o/${tag}_bench_c5_cw$cw.err; echo "bench cw=$cw rc=$?"; tail -2 $o/${tag}_bench_c5_cw$cw.err
  DFR2D_WS_CW=$cw timeout 300 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline --no-also > $o/${tag}_bench_c2_cw$cw.json 2> $o/${tag}_bench_c2_cw$cw.err
  DFR2D_WS_CW=$cw timeout 300 pytest -q -x tests/test_gpu_parity.py -m gpu -k "vortex_steps" > $o/${tag}_pytest_cw$cw.log 2>&1; tail -1 $o/${tag}_pytest_cw$cw.log
done
python - <<'PY'
import json
for cw in (8,12):
    for wl in ("c5","c2"):
        l=json.loads(open('gpurun_out/r02f_bench_%s_cw%d.json'%(wl,cw)).read().strip().splitlines()[-1])
        print(wl, "cw", cw, l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["phase_ms"]["element kernel"], l["clocks"])
PY
DFR2D_WS_CW=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_elem_ws' -s 6 -c 2 -f -o $o/${tag}_elem_tma \
    python bench.py --nx 1000 --steps 2 --warmup 3 --no-cpu-baseline --no-also > $o/${tag}_ncu_elem.log 2>&1
tail -1 $o/${tag}_ncu_elem.log
