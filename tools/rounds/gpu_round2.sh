#!/bin/bash
# kernel-level round: gemm microbench, dtp microbench (variants / tiles), ncu full captures of the two DTP kernels
set -u
TAG=${1:-r1d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== gemm tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -5
echo "== gemm microbench"; timeout 300 python tools/gemm_microbench.py 32560 2>&1 | tee $OUT/gemm_microbench.jsonl | tail -12
echo "== dtp microbench"
for cfg in "tma 8" "tma 4" "vec 8" "vec 4" "scalar 8"; do
  set -- $cfg
  EQF_DTP_VARIANT=$1 EQF_TILE_EDGES=$2 timeout 120 python tools/dtp_microbench.py qm9_l2 32560 20 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
done
EQF_DTP_VARIANT=tma timeout 120 python tools/dtp_microbench.py qm9_l2 500000 5 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
EQF_DTP_VARIANT=tma timeout 120 python tools/dtp_microbench.py md17_l3 20000 10 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
EQF_DTP_VARIANT=tma timeout 120 python tools/dtp_microbench.py oc20_l1 58000 10 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
echo "== ncu full: dtp_forward_vec"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtp_forward_vec -s 3 -c 1 -o $OUT/prof_dtp_forward python tools/dtp_microbench.py qm9_l2 32560 1 > $OUT/ncu_fwd.log 2>&1; echo "rc=$?"
echo "== ncu full: dtp_grad_x_vec"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtp_grad_x_vec -s 3 -c 1 -o $OUT/prof_dtp_grad_xw python tools/dtp_microbench.py qm9_l2 32560 1 > $OUT/ncu_bwd.log 2>&1; echo "rc=$?"
ls -la $OUT
