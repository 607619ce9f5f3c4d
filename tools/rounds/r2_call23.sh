#!/bin/bash
# radial first-layer hoist: model-level parity + A/B on one box
set -u
TAG=${1:-r2c23}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== model-level tests"
timeout -k 10 1500 python -m pytest tests/test_reference_golden.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -5
for h in 1 0 1 0; do
  EQF_RAD_HOIST=$h timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9_hoist$h.json 2> $OUT/bench_h$h.err; echo "hoist=$h rc=$?"
  python -c "
import json; d=json.load(open('$OUT/bench_qm9_hoist$h.json')); print('hoist=$h', round(d['ms_per_step'],3), round(d['value']), d['gpu_launches'])"
done
for h in 1 0; do
  EQF_RAD_HOIST=$h timeout -k 10 600 python bench.py --workload oc20_l1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_oc20_hoist$h.json 2> $OUT/bench_oh$h.err
  EQF_RAD_HOIST=$h timeout -k 10 600 python bench.py --workload md17_l3 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_md17_hoist$h.json 2> $OUT/bench_mh$h.err
  python -c "
import json
for w in ('oc20','md17'):
    d=json.load(open('$OUT/bench_%s_hoist$h.json' % w)); print(w, 'hoist=$h', round(d['ms_per_step'],3), round(d['value']), d['gpu_launches'])"
done
