#!/bin/bash
# quick validation: GPU tests, smoke, one bench line without the CPU leg
set -u
TAG=${1:-r3d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-300 $OUT/bench.json
python - <<PY
import json
d = json.load(open("$OUT/bench.json"))
print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if k.startswith("attn")})
PY
