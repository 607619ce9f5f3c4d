#!/bin/bash
# final evidence: launch list of the headline region from the final code; ncu --set full of the grouped warp-MMA kernel
set -u
TAG=${1:-r2c27}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu launch list (headline region only)"
EQF_BENCH_CUDA_PROFILER=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu rc=$?"; wc -l $OUT/launches.csv
python tools/summarise_launches.py $OUT/launches.csv 40 | tee $OUT/launches_summary.txt | head -44
echo "== ncu --set full: grouped_gemm_mma_kernel (backward launch of a node-level linear: 3 data + 3 weight gradients)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm_mma -s 8 -c 1 -o $OUT/prof_grouped_mma python tools/grouped_microbench.py > $OUT/ncu_grouped.log 2>&1; echo rc=$?
ncu -i $OUT/prof_grouped_mma.ncu-rep --page raw --csv > $OUT/prof_grouped_mma_raw.csv 2>/dev/null; wc -l $OUT/prof_grouped_mma_raw.csv
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/prof_grouped_mma_raw.csv")))
hdr, vals = rows[0], rows[-1]
d = dict(zip(hdr, vals))
for k in ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
          "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]:
    if k in d: print(k, "=", d[k][:90])
PY
