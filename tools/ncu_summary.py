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
        # blocks: ["Kernel Name", name] / header / one row per SASS instruction
        k = 0
        while k < len(rows):
            if rows[k] and rows[k][0] == "Kernel Name":
                name, h = rows[k][1], rows[k + 1]
                ci = {c: i for i, c in enumerate(h)}
                body = []
                k += 2
                while k < len(rows) and rows[k] and rows[k][0] != "Kernel Name":
                    if len(rows[k]) >= len(h) - 2:
                        body.append(rows[k])
                    k += 1
                samp = ci["# Samples"]
                tot = sum(int(r[samp]) for r in body) or 1
                stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
                print(f"\n## SASS hot spots of {name[:90]} ({tot} samples, {len(body)} instructions)")
                agg = {c: sum(int(r[ci[c]] or 0) for r in body) for c in stall_cols}
                print("   by reason: " + ", ".join(f"{c[6:]} {100 * v / tot:.1f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
                op = {}
                for r in body:
                    o = r[ci["Source"]].split()[0] if r[ci["Source"]].split() else "?"
                    if o.startswith("@"):
                        o = r[ci["Source"]].split()[1]
                    op[o] = op.get(o, [0, 0])
                    op[o][0] += int(r[samp]); op[o][1] += int(r[ci["Instructions Executed"]] or 0)
                itot = sum(v[1] for v in op.values()) or 1
                print("   by opcode (samples%, executed%): " + ", ".join(f"{o} {100 * v[0] / tot:.1f}/{100 * v[1] / itot:.1f}" for o, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:14]))
                for r in sorted(body, key=lambda r: -int(r[samp]))[:n]:
                    top = max(stall_cols, key=lambda c: int(r[ci[c]] or 0))
                    print(f"{100 * int(r[samp]) / tot:6.2f}%  {top[6:]:14s} {r[ci['Source']].strip()[:110]}")
            else:
                k += 1


if __name__ == "__main__":
    main()
