#!/usr/bin/env python3
"""Stall samples / executed instructions of one kernel by SOURCE FUNCTION (or by line): joins an `ncu --page source --csv`
export (SASS level) with the line table of the same binary (`nvdisasm -g -c <cubin>`).

    python tools/ncu_lines.py <src.csv> <sass-with-line-info> <kernel substring> [--lines FILE] [--inline]

Each SASS instruction is attributed to the innermost source line (default) -- i.e. helper code inlined into the kernel counts for
the helper's own file/function.  Functions are recognised by their DRC_HD / template / static definitions."""
import bisect
import collections
import csv
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "dyros_robot_controller_b200" / "csrc"


def functions(path):
    out = []
    for i, l in enumerate(path.read_text().split("\n")):
        m = re.match(r"^(?:static |inline |DRC_HD(?:_NOINLINE)? |__device__ |__forceinline__ |__global__ )+[\w:<>&\*, ]*?\b(\w+)\(", l)
        if m and not l.startswith(" "):
            out.append((i + 1, m.group(1)))
    return out


def main():
    src_csv, sass, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    per_line = "--lines" in sys.argv
    want_file = sys.argv[sys.argv.index("--lines") + 1] if per_line else None
    loc, cur, infn = [], ("?", 0), False
    for ln in open(sass):
        if ".section" in ln and ".text." in ln:
            infn = kernel in ln
        elif ".section" in ln:
            infn = False
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
            loc.append(cur)
    rows = list(csv.reader(open(src_csv)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    body = [dict(zip(hdr, r)) for r in rows[h + 1:] if len(r) == len(hdr)]
    assert len(body) == len(loc), (len(body), len(loc))
    fn_tables = {p.name: functions(p) for p in CSRC.glob("*.h")} | {p.name: functions(p) for p in CSRC.glob("*.cu*")}
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot_s = tot_e = 0
    for (f, l), r in zip(loc, body):
        s, e = int(r["# Samples"] or 0), int(r["Instructions Executed"] or 0)
        tot_s += s; tot_e += e
        if per_line:
            key = f"{f}:{l}" if f == want_file else f
        else:
            t = fn_tables.get(f)
            if t:
                k = bisect.bisect_right([a for a, _ in t], l) - 1
                key = f"{f}::{t[k][1]}" if k >= 0 else f
            else:
                key = f
        a = agg[key]
        a[0] += s; a[1] += e; a[2] += 1
    print(f"kernel {kernel}: {len(loc)} SASS instructions, {tot_e} warp instructions executed, {tot_s} samples")
    print("| source function | SASS | samples | share | warp instr. executed | share |\n|---|---|---|---|---|---|")
    for k, (s, e, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"| {k} | {n} | {s} | {100.0 * s / max(tot_s, 1):.1f} % | {e} | {100.0 * e / max(tot_e, 1):.1f} % |")


if __name__ == "__main__":
    main()
