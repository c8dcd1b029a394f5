"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table
(markdown).  Usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/x.md"""
import collections
import csv
import sys


def main(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = r["Kernel Name"]
        n = n[n.index("ups::"):].split("(")[0] if "ups::" in n else "torch: " + n.split("(")[0][-70:]
        agg[n][0] += 1
        agg[n][1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if k.startswith("ups::"))
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print("| `%s` | %d | %.3f | %.1f%% |" % (n, c, t / 1e6, 100 * t / tot))
    print("\n%d launches, %.3f ms of kernel time (ncu: cold-cache, serialised); "
          "upsnet_b200 kernels = %.1f%% of it." % (len(rows), tot / 1e6, 100 * mine / tot))


if __name__ == "__main__":
    main(sys.argv[1])
