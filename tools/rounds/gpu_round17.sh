#!/bin/bash
set -u
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
EQF_PROFILE_GEMM_SHAPES=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_shapes.json 2> $OUT/bench_shapes.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_shapes.json"))
k=d["kernels"]; tot=0
for n,v in sorted(k.items(), key=lambda x:-x[1]["ms_per_step"]):
    if "gemm" in n: print(f"{n:52s} {v['launches_per_step']:5.0f} {v['ms_per_step']:7.3f} ms  {1e3*v['ms_per_step']/v['launches_per_step']:6.1f} us each")
PY
