#!/usr/bin/env python
"""Condenses `ncu --page raw --csv` (+ optionally `--page source --csv`) of ONE kernel launch into the text kept under
profiles/:  python tools/ncu_summary.py raw.csv [source.csv] [--cells N] > profiles/<name>_ncu_summary.txt"""
import csv
import sys

csv.field_size_limit(1 << 30)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
cells = float(sys.argv[sys.argv.index("--cells") + 1]) if "--cells" in sys.argv else None
rows = list(csv.reader(open(args[0])))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
print(f"# ncu --set full, one launch; source: {args[0]}")
for k in KEYS:
    if k in d:
        print(f"{k:75s} {d[k][0]:>22s} {d[k][1]}")
print("# warp stall reasons, cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active)")
st = []
for k, v in d.items():
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            st.append((float(v[0]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
for val, name in sorted(st, reverse=True):
    if val >= 0.05:
        print(f"  {name:28s} {val:6.2f}")
if cells and "smsp__inst_executed.sum" in d:
    wi = float(d["smsp__inst_executed.sum"][0])
    tpi = float(d.get("smsp__thread_inst_executed_per_inst_executed.ratio", ("32", ""))[0])
    print(f"# {cells:.0f} cells in this launch: {wi * 32 / cells:.2f} issued lane-slots per cell ({wi * tpi / cells:.2f} executed thread-instructions per cell)")
if len(args) > 1:
    src = list(csv.reader(open(args[1])))
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    hdr = src[hi]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in src[hi + 1:] if len(r) == len(hdr)]
    tot = sum(float(r[col["# Samples"]] or 0) for r in body) or 1.0
    ops = {}
    for r in body:
        op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[col["Source"]].split()[1]
        op = op.split(".")[0]
        e = ops.setdefault(op, [0.0, 0.0])
        e[0] += float(r[col["Instructions Executed"]] or 0)
        e[1] += float(r[col["# Samples"]] or 0)
    tot_i = sum(e[0] for e in ops.values()) or 1.0
    print("# SASS opcodes: share of executed warp-instructions / share of stall samples")
    for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:16]:
        extra = f"  {n * 32 / cells:6.3f} per cell" if cells else ""
        print(f"  {op:14s} {100 * n / tot_i:6.2f} %   {100 * s / tot:6.2f} %{extra}")
