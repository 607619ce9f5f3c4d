#!/bin/bash
# first call of the next round: everything that was written after the round-1 GPU budget ended
#   - the full GPU suite (includes test_cuda_headline_model_matches_reference_model_file, test_cuda_block_matches_reference_at_oc20_sizes and
#     test_cuda_dot_product_attention_matches_reference_model_file of tests/test_reference_golden.py, never run on a B200)
#   - the 128-column tensor-memory weight gradient (EQF_TF32X3_WGRAD_TS=2): parity + timing against the shared-memory kernel
#   - one bench line with and without it
set -u
TAG=${1:-r4a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== wgrad through tensor memory, 65 ... 128 columns"
EQF_WGRAD_TS_LEVEL=2 timeout 400 python tools/tf32x3_wgrad_ts_check.py $OUT/wgrad_ts_level2.jsonl | tail -40
echo "== bench (default)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cut -c1-260 $OUT/bench.json
echo "== bench (EQF_TF32X3_WGRAD_TS=2)"; EQF_TF32X3_WGRAD_TS=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_ts2.json 2> $OUT/bench_ts2.err; echo "rc=$?"; cut -c1-260 $OUT/bench_ts2.json
