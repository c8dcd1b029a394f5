#!/bin/bash
run() { echo "== $1"; env $2 python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
run "default" "X=1"
run "no sampler" "UPSNET_BENCH_NO_SAMPLER=1"
run "no numa" "UPSNET_BENCH_NO_NUMA=1"
run "neither" "UPSNET_BENCH_NO_SAMPLER=1 UPSNET_BENCH_NO_NUMA=1"
