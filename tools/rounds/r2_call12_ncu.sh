#!/bin/bash
# ncu --set full captures: the fused K1 kernel (l = 2 group), the K2 kernel, and the new DeNS / CUDA tests
set -u
TAG=${1:-r2c12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== dens + fused-vs-unfused model tests"
timeout -k 10 600 python -m pytest tests/test_reference_golden.py tests/test_gpu_model.py -m gpu -q -x -k "dens or fused_forward_model" 2>&1 | tail -5
echo "== K2 microbench"; timeout 120 python tools/attn_microbench.py | tee $OUT/attn_microbench.jsonl
echo "== ncu K1 (dtp_gemm_fwd, l=2 N=32 gather)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtp_gemm_fwd -s 3 -c 1 -o $OUT/prof_k1_l2 python tools/fused_timeline.py qm9_l2 32560 2 32 1 > $OUT/ncu_k1.log 2>&1; echo rc=$?
ncu -i $OUT/prof_k1_l2.ncu-rep --page raw --csv > $OUT/prof_k1_l2_raw.csv 2>/dev/null; wc -l $OUT/prof_k1_l2_raw.csv
echo "== ncu K2 (softmax_aggregate)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softmax_aggregate -s 3 -c 1 -o $OUT/prof_k2 python tools/attn_microbench.py 32560 2324 5 > $OUT/ncu_k2.log 2>&1; echo rc=$?
ncu -i $OUT/prof_k2.ncu-rep --page raw --csv > $OUT/prof_k2_raw.csv 2>/dev/null; wc -l $OUT/prof_k2_raw.csv
python - <<PY
import csv
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "smsp__inst_executed.avg.per_cycle_active"]
for n in ["prof_k1_l2", "prof_k2"]:
    try:
        rows = list(csv.reader(open("$OUT/%s_raw.csv" % n)))
        hdr, vals = rows[0], rows[-1]
        print("==", n, vals[hdr.index("Kernel Name")][:60])
        for k in keys:
            if k in hdr: print("   ", k, "=", vals[hdr.index(k)], rows[1][hdr.index(k)])
        for i, h in enumerate(hdr):
            if "tensor" in h and "pct" in h or "issue_stalled" in h and "pct" not in h: pass
    except Exception as e: print(n, "failed", e)
PY
