#!/bin/bash
# per-family DRAM traffic of the GEMM kernels inside the headline region + 2-GPU bench
set -u
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu dram bytes, GEMM kernels of the headline region"
EQF_BENCH_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off -k regex:tf32x3 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $OUT/gemm_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_traffic.log 2>&1 ; echo "ncu rc=$?"; wc -l $OUT/gemm_traffic.csv
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$OUT/gemm_traffic.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.defaultdict(lambda: collections.Counter())
cnt=collections.Counter()
scale={"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9,"ns":1,"us":1e3,"ms":1e6,"usecond":1e3,"nsecond":1,"msecond":1e6}
for r in rows[1:]:
    fam = "gemm_tf32x3_wgrad" if "wgrad" in r[ki] else ("tf32x3_split" if "split" in r[ki] else "gemm_tf32x3")
    v=float(r[vi].replace(",",""))*scale.get(r[ui],1)
    agg[fam][r[mi]]+=v
    if r[mi]=="gpu__time_duration.sum": cnt[fam]+=1
for fam,m in agg.items():
    tot=m["dram__bytes_read.sum"]+m["dram__bytes_write.sum"]
    print(fam, "launches", cnt[fam], "dram bytes/launch %.0f" % (tot/cnt[fam]), "time us/launch %.1f" % (m["gpu__time_duration.sum"]/cnt[fam]/1e3), "GB/s of dram traffic %.0f" % (tot/m["gpu__time_duration.sum"]))
PY
