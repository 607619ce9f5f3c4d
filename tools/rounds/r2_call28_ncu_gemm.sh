#!/bin/bash
# ncu --set full of the dominant GEMM kernels in their final form (wait hint + two-instruction rounding): the stacked
# tensor-memory kernel on [162 800, 352] -> 32 and the wide shared-memory kernel on [162 800, 32] -> 352 (data gradient)
set -u
TAG=${1:-r2c28}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cat > /tmp/one_gemm.py <<PY
import sys, torch
sys.path.insert(0, ".")
from equiformer_b200 import ops
M, K, N = [int(v) for v in sys.argv[1:4]]
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g); Bt = torch.randn(N, K, device="cuda", generator=g)
for _ in range(4): ops.gemm_tf32x3_raw(A, Bt)
torch.cuda.synchronize()
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_ts -s 2 -c 1 -o $OUT/prof_gemm_ts32 python /tmp/one_gemm.py 162800 352 32 > $OUT/ncu_ts.log 2>&1; echo rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 2 -c 1 -o $OUT/prof_gemm_ss256 python /tmp/one_gemm.py 162800 32 352 > $OUT/ncu_ss.log 2>&1; echo rc=$?
for n in prof_gemm_ts32 prof_gemm_ss256; do ncu -i $OUT/$n.ncu-rep --page raw --csv > $OUT/${n}_raw.csv 2>/dev/null; done
python - <<PY
import csv
for n in ["prof_gemm_ts32", "prof_gemm_ss256"]:
    rows = list(csv.reader(open("$OUT/%s_raw.csv" % n)))
    d = dict(zip(rows[0], rows[-1]))
    print("==", n, d["Kernel Name"][:70])
    for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_tc.sum", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]:
        if k in d: print("   ", k, "=", d[k])
PY
