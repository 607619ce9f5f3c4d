#!/bin/bash
# single small products on the own kernel (A/B vs cuBLAS), attention-dropout line, refreshed launch list of the headline region
set -u
TAG=${1:-r2c18}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== model-level + gemm tests"
timeout -k 10 1200 python -m pytest tests/test_reference_golden.py tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -x -k "not fused_dtp" 2>&1 | tail -5
echo "== bench md17 / qm9: own small kernel vs cuBLAS"
for w in md17_l3 qm9; do
  timeout -k 10 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${w}.json 2> $OUT/bench_${w}.err; echo "rc=$?"
  EQF_SMALL_GEMM=cublas timeout -k 10 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${w}_cublas_small.json 2> $OUT/bench_${w}_c.err; echo "rc=$?"
done
echo "== attention dropout 0.2"
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --alpha-drop 0.2 > $OUT/bench_qm9_alpha_drop.json 2> $OUT/bench_ad.err; echo "rc=$?"; tail -2 $OUT/bench_ad.err
python - <<PY
import json
for n in ["bench_md17_l3", "bench_md17_l3_cublas_small", "bench_qm9", "bench_qm9_cublas_small", "bench_qm9_alpha_drop"]:
    try:
        d = json.load(open("$OUT/%s.json" % n))
        print(n, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "roofline", d["roofline"].get("frac"), d["roofline"].get("kernel"))
    except Exception as e:
        print(n, "failed", e)
PY
echo "== ncu launch list (headline region only)"
EQF_BENCH_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu rc=$?"; wc -l $OUT/launches.csv
python tools/summarise_launches.py $OUT/launches.csv 60 | tee $OUT/launches_summary.txt | head -64
