#!/bin/bash
# K1 v5 (independent rings, table helper warps, prefetching dense producer): parity, A/B + skip modes, timeline; GEMM defaults
set -u
TAG=${1:-r2c15}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== fused tests"
timeout -k 10 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_dtp_linear" 2>&1 | tail -5
rc=$?
echo "== A/B fused"
timeout -k 10 400 python tools/v4_ab.py fused | tee $OUT/v5_ab_fused.jsonl
echo "== timeline l=2"
timeout -k 10 120 python tools/fused_timeline.py qm9_l2 32560 2 32 1 > $OUT/fused_fwd_v5_timeline_l2_dtp1.txt; tail -12 $OUT/fused_fwd_v5_timeline_l2_dtp1.txt | cut -c1-600
echo "== gemm tests + A/B (hint and fast rounding are the defaults now; 8 / 16 turn them off)"
timeout -k 10 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tf32x3_tcgen05 or gemm_autograd" 2>&1 | tail -3
timeout -k 10 400 python tools/v4_ab.py gemm | tee $OUT/v5_ab_gemm.jsonl
