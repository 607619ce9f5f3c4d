#!/bin/bash
# MD17: flop-based tcgen05 threshold; generated vs table-driven DTP kernels at E = 2 100 edges
set -u
TAG=${1:-r2c21}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in gen tma v3 vec; do
  echo "== dtp microbench md17_l3 E=2100 variant $v"
  EQF_DTP_VARIANT=$v timeout 200 python tools/dtp_microbench.py md17_l3 2100 50 2>&1 | tail -8 | cut -c1-400 | tee $OUT/dtp_microbench_md17_E2100_$v.txt
done
for v in gen tma v3; do
  EQF_DTP_VARIANT=$v timeout -k 10 900 python bench.py --workload md17_l3 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_md17_$v.json 2> $OUT/bench_md17_$v.err; echo "rc=$?"
done
EQF_SMALL_GEMM=cublas timeout -k 10 900 python bench.py --workload md17_l3 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_md17_gen_cublas_small.json 2> $OUT/bench_md17_c.err; echo "rc=$?"
EQF_GEMM_MIN_FLOP=1e30 timeout -k 10 900 python bench.py --workload md17_l3 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_md17_gen_rows_only.json 2> $OUT/bench_md17_r.err; echo "rc=$?"
python - <<PY
import json
for n in ["bench_md17_gen", "bench_md17_tma", "bench_md17_v3", "bench_md17_gen_cublas_small", "bench_md17_gen_rows_only"]:
    try:
        d = json.load(open("$OUT/%s.json" % n))
        print(n, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "launches", d["gpu_launches"])
        ks = d["kernels"]
        print("   ", [(k, round(v["ms_per_step"], 2)) for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms_per_step"])[:7]])
    except Exception as e:
        print(n, "failed", e)
PY
