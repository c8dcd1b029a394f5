#!/bin/bash
# One GPU-box visit: pair-stream tests in their own process, the rest of the GPU suite, then a short bench per precision.
# Usage (under gpurun): bash scripts/gpu_round.sh [tag]
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -x -q -m gpu > gpurun_out/${tag}_pair_tests.log 2>&1
echo "pair_tests rc=$?" | tee -a gpurun_out/${tag}_pair_tests.log
tail -25 gpurun_out/${tag}_pair_tests.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_pair.py > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu_tests rc=$?" | tee -a gpurun_out/${tag}_gpu_tests.log
tail -8 gpurun_out/${tag}_gpu_tests.log
UPSNET_LAYER_TABLE=gpurun_out/${tag}_layers_x3.md timeout 600 python bench.py --precision bf16x3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_x3.json 2> gpurun_out/${tag}_bench_x3.err
echo "bench x3 rc=$?"; tail -c 1500 gpurun_out/${tag}_bench_x3.json; tail -5 gpurun_out/${tag}_bench_x3.err
