#!/bin/bash
# One GPU-box visit: pair-stream tests in their own process, the rest of the GPU suite, then the default bench.
# Usage (under gpurun): bash scripts/gpu_round.sh [tag] [bench args...]
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -x -q -m gpu > gpurun_out/${tag}_pair_tests.log 2>&1
echo "pair_tests rc=$?"; tail -4 gpurun_out/${tag}_pair_tests.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_pair.py -s > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu_tests rc=$?"; grep -E "config[23] max rel|passed|failed|Error|FAILED" gpurun_out/${tag}_gpu_tests.log | tail -25
UPSNET_LAYER_TABLE=gpurun_out/${tag}_layers.md timeout 900 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
