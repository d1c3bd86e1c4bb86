"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel.

    python benchmarks/summarize_launches.py gpurun_out/launches.csv [--steps 5] > profiles/rN_launches.txt

Per-launch times under ncu are cold-cache and serialised: use the SHARES, not the absolute numbers.
"""
import csv
import collections
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "")
    return name[:110]


def main():
    path = sys.argv[1]
    steps = None
    if "--steps" in sys.argv:
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    tot = sum(ns for _, ns in rows)
    agg = collections.OrderedDict()
    for k, ns in rows:
        a = agg.setdefault(short(k), [0, 0.0])
        a[0] += 1
        a[1] += ns
    ours = sum(v[1] for k, v in agg.items() if k.startswith("vf::"))
    print(f"# {path}: {len(rows)} launches, {tot/1e6:.2f} ms total kernel time under ncu"
          + (f" ({steps} steps -> {tot/1e6/steps:.2f} ms/step)" if steps else ""))
    print(f"# vface_b200 kernels (vf::*): {ours/tot*100:.1f} % of kernel time, "
          f"{sum(v[0] for k, v in agg.items() if k.startswith('vf::'))} launches")
    print(f"{'share%':>7} {'ms':>9} {'calls':>6} {'avg_us':>9}  kernel")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns/tot*100:7.2f} {ns/1e6:9.3f} {n:6d} {ns/n/1e3:9.1f}  {k}")


if __name__ == "__main__":
    main()
