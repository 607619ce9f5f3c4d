#!/bin/bash
# bench with the tuned GEMM mainloop, launch list of the headline region, ncu full capture of one GEMM launch
set -u
TAG=${1:-r1m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu gemm+model"; timeout 900 python -m pytest tests -m gpu -q -x -k "gemm or model or golden" 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-300 $OUT/bench.json
echo "== ncu launch list (headline region only)"
EQF_BENCH_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu rc=$?"; wc -l $OUT/launches.csv
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: t=float(r[vi].replace(",",""))
    except: continue
    a=agg[r[ki][:90]]; a[0]+=1; a[1]+=t
tot=sum(v[1] for v in agg.values())
print("launches",sum(v[0] for v in agg.values()),"total_ms",tot/1e6)
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:60]:
    print(f"{v[1]/1e6:8.3f} ms {v[0]:5d} {100*v[1]/tot:5.1f}%  {k}")
PY
echo "== ncu full: one fast-fp32 GEMM launch (val1_l0 forward)"
EQF_GEMM_FORCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:device_kernel -s 6 -c 1 -o $OUT/prof_gemm_fwd python tools/gemm_microbench.py > $OUT/ncu_gemm.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_gemm_fwd.ncu-rep --page raw --csv > $OUT/prof_gemm_fwd_raw.csv 2>/dev/null; wc -l $OUT/prof_gemm_fwd_raw.csv
