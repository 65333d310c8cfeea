#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into the few numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls N] > profiles/rNN_<what>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "sm__cycles_elapsed.avg",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, rows = raw(path)
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    for r in rows:
        print(f"\n## {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for k in KEYS:
            if k in col and r[col[k]] != "":
                print(f"{k:75s} {r[col[k]]:>18s} {units[col[k]]}")
        stalls = [(float(r[i].replace(',', '')), h) for h, i in col.items()
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
        if not stalls:
            stalls = [(float(r[i].replace(',', '')), h) for h, i in col.items()
                      if h.startswith("smsp__average_warp_latency_issue_stalled") and r[i]]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {h:90s} {v:10.3f}")
    if "--stalls" in sys.argv:
        n = int(sys.argv[sys.argv.index("--stalls") + 1])
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        # first kernel only: header row, then one row per SASS/source line
        try:
            h = rows[0]
            ci = {name: i for i, name in enumerate(h)}
            samp = next(k for k in ci if k.startswith("# Samples") or k == "Warp Stall Sampling (All Samples)")
            body = [r for r in rows[1:] if len(r) == len(h) and r[ci[samp]].replace(',', '').isdigit()]
            body.sort(key=lambda r: -int(r[ci[samp]].replace(',', '')))
            print(f"\n## top {n} lines by stall samples ({samp})")
            for r in body[:n]:
                print(r[ci[samp]].rjust(8), r[ci.get('Source', 1)][:150])
        except Exception as e:       # noqa: BLE001
            print("source page not parsed:", e)


if __name__ == "__main__":
    main()
