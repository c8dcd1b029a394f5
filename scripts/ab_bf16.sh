#!/bin/bash
# same-box A/B of the bf16 configuration: round-1 tree (_r1, worktree of c35a7b2) vs the current tree
mkdir -p gpurun_out
( cd _r1 && UPSNET_LAYER_TABLE=../gpurun_out/ab_r1_layers_bf16.md python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > ../gpurun_out/ab_r1_bench_bf16.json 2>/dev/null )
UPSNET_LAYER_TABLE=gpurun_out/ab_r2_layers_bf16.md python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_r2_bench_bf16.json 2>/dev/null
python - <<'PY'
import json
for t in ("r1","r2"):
    d=json.loads(open("gpurun_out/ab_%s_bench_bf16.json"%t).read().strip().splitlines()[-1])
    print(t, d["value"], d["ms_per_step"], d["roofline"]["families_ms_per_step"])
PY
