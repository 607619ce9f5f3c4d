#!/bin/bash
# v4 switches: parity with every switch on, then the in-process A/B (fused kernel, GEMM / weight-gradient kernels)
set -u
TAG=${1:-r2c13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== fused tests, padded layout (switches off)"
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_dtp_linear" 2>&1 | tail -3
echo "== fused tests, all v4 switches"
EQF_FUSED_DBG_SKIP=248 timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_dtp_linear" 2>&1 | tail -3
echo "== gemm tests, switches 24"
EQF_TF32X3_DBG_SKIP=24 timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tf32x3_tcgen05 or gemm_autograd" 2>&1 | tail -3
echo "== A/B fused"
timeout 600 python tools/v4_ab.py fused | tee $OUT/v4_ab_fused.jsonl
echo "== A/B gemm"
timeout 600 python tools/v4_ab.py gemm | tee $OUT/v4_ab_gemm.jsonl
