"""Stall-sample hot spots of one launch in an .ncu-rep (SASS view), grouped by warp-role region when markers are given.

    python benchmarks/ncu_sass_top.py x.ncu-rep [launch_index] [top_n]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
num = lambda v: int(float(v)) if v not in ("", None) else 0
tot = sum(num(r[isamp]) for r in data)
print(rows[0][1] if rows and len(rows[0]) > 1 else "", "| total samples", tot, "| instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -num(data[i][isamp]))[:topn]
for i in order:
    r = data[i]
    st = {hdr[c][6:]: num(r[c]) for c in stall_cols if num(r[c]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{num(r[isamp]):6d} {100 * num(r[isamp]) / max(tot, 1):5.1f}%  #{i:5d} ex {num(r[iex]):>9d}  {r[isrc][:60]:60s} {st}")
