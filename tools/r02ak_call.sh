#!/bin/bash
# Round 2, call ak (1 GPU): round profile of the final build (tests, bench, launch list, ncu full) + capture cost
bash tools/run_round_profile.sh r02ak
timeout 300 python tools/capture_cost.py > gpurun_out/r02ak_capture_cost.json 2> gpurun_out/r02ak_capture_cost.err; echo "capture rc=$?"; cat gpurun_out/r02ak_capture_cost.json
exit 0
