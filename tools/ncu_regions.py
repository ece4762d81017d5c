#!/usr/bin/env python3
"""Stall samples of k_admm by code region: joins an `ncu --page source --csv` export (SASS level) with the line table of the
same binary (`nvdisasm -g -c` of its cubin) and groups the instructions by the drc_qp.h line they were generated from.

    python tools/ncu_regions.py <src.csv> <sass with line info> [out.md]
"""
import csv
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REGIONS = [("load", "// ---------------------------------------------------------------- load"),
           ("Ruiz equilibration (10 passes)", "Ruiz equilibration (OSQP scale_data)"),
           ("bounds / classes", "scaled bounds, constraint classes"),
           ("factorisation", "factorisation (setup and rho updates)"),
           ("hot loop (phase A + B)", "hot loop phases (branch-free)"),
           ("checked iteration (phase B keep)", "phase B on a checked iteration"),
           ("loop control", "ADMM iterations (osqp_solve)"),
           ("termination check + rho adaptation", "OSQP update_info + check_termination")]


def main():
    src_csv, sass = sys.argv[1], sys.argv[2]
    qp = (ROOT / "dyros_robot_controller_b200" / "csrc" / "drc_qp.h").read_text().split("\n")
    marks = []
    for name, needle in REGIONS:
        marks.append((name, next(i + 1 for i, l in enumerate(qp) if needle in l)))
    marks.append(("end", len(qp) + 1))
    # instruction index -> region, from the line table
    region_of, cur, infn = [], None, False
    for ln in open(sass):
        if ln.startswith(".text.") or ".section" in ln:
            infn = "k_admm" in ln
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m and m.group(1).endswith("drc_qp.h"):
            cur = int(m.group(2))
        if infn and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
            reg = "prologue / emit"
            if cur is not None:
                for (n, a), (_, b) in zip(marks, marks[1:]):
                    if a <= cur < b:
                        reg = n
            region_of.append(reg)
    rows = list(csv.reader(open(src_csv)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    body = [dict(zip(hdr, r)) for r in rows[h + 1:] if len(r) == len(hdr)]
    assert len(body) == len(region_of), (len(body), len(region_of))
    keys = ["stall_wait", "stall_no_inst", "stall_selected", "stall_branch_resolving", "stall_short_sb", "stall_math", "stall_not_selected", "stall_dispatch"]
    agg = {}
    for reg, r in zip(region_of, body):
        a = agg.setdefault(reg, dict(n=0, samples=0, exe=0, **{k: 0 for k in keys}))
        a["n"] += 1
        a["samples"] += int(r["# Samples"] or 0)
        a["exe"] += int(r["Instructions Executed"] or 0)
        for k in keys:
            a[k] += int(r.get(k) or 0)
    tot = sum(a["samples"] for a in agg.values())
    out = ["| region | SASS instr. | samples | share | wait | no_inst | selected | branch | short_sb | math | warp instr. executed |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    order = ["prologue / emit"] + [n for n, _ in REGIONS]
    for n in order:
        if n not in agg:
            continue
        a = agg[n]
        out.append(f"| {n} | {a['n']} | {a['samples']} | {100.0 * a['samples'] / tot:.1f} % | {a['stall_wait']} | {a['stall_no_inst']} | {a['stall_selected']} | "
                   f"{a['stall_branch_resolving']} | {a['stall_short_sb']} | {a['stall_math']} | {a['exe']} |")
    tk = {k: sum(a[k] for a in agg.values()) for k in keys}
    out.append("")
    out.append("Totals: " + ", ".join(f"{k[6:]} {100.0 * v / tot:.1f} %" for k, v in sorted(tk.items(), key=lambda kv: -kv[1])) +
               f"; {sum(a['n'] for a in agg.values())} SASS instructions, {sum(a['exe'] for a in agg.values())} warp instructions executed.")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 3:
        Path(sys.argv[3]).write_text(text + "\n")


if __name__ == "__main__":
    main()
