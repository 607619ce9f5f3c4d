"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total time, share."""
import collections
import csv
import sys


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            t = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg[r[ki][:100]]
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"launches {sum(v[0] for v in agg.values())} total_ms {tot / 1e6:.3f}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{v[1] / 1e6:8.3f} ms {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main()
