"""Key metrics of the first kernel in an .ncu-rep (read here, no GPU needed):

    python benchmarks/ncu_summary.py gpurun_out/x.ncu-rep [--json out.json] > profiles/rN_x_ncu.txt
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {}
    for ri, vals in enumerate(rows[2:]):
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print(f"== launch {ri}: {d.get('Kernel Name', ('?',))[0][:120]}")
        for k in KEYS:
            if k in d:
                print(f"  {k:82s} {d[k][0]:>16s} {d[k][1]}")
        stalls = sorted(((float(v[0].replace(',', '')), k[len(STALL):].replace('_per_issue_active.ratio', ''))
                         for k, v in d.items() if k.startswith(STALL) and v[0]), reverse=True)
        print("  warp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n}={x:.2f}" for x, n in stalls[:8]))
        if ri == 0:
            def num(k):
                v, u = d[k]
                x = float(v.replace(",", ""))
                return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
            out = dict(kernel=d["Kernel Name"][0], dram_bytes_per_launch=num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                       duration_ms_under_ncu=float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[d["gpu__time_duration.sum"][1]],
                       source=rep)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
