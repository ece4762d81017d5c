#!/usr/bin/env python3
"""Summarise `ncu --page raw --csv` exports (gpurun_out/<tag>_<kernel>_raw.csv) into profiles/<name>.md:
the launch configuration, occupancy limits, issue / FP64-pipe utilisation, DRAM traffic and the stall mix of each kernel.

    python tools/ncu_summary.py <tag> <out.md>
"""
import csv
import glob
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (blocks/SM)"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps)"),
    ("smsp__cycles_active.avg", "SMSP active cycles (avg)"), ("sm__cycles_elapsed.max", "elapsed cycles (max)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (%)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads"), ("smsp__inst_executed_op_local_st.sum", "local stores"),
]


def main():
    tag, out = sys.argv[1], Path(sys.argv[2])
    lines = [f"# ncu --set full summaries ({tag})", "",
             "Captured with `ncu --set full --clock-control none --import-source on -k regex:^<kernel>$ -c 1 python tools/prof_cycle.py 65536 1`",
             "(FR3 updateState+QPIKStep, batch 65536, one launch of each kernel; cold-cache, serialised: shares, not absolutes).", ""]
    for f in sorted(glob.glob(str(ROOT / "gpurun_out" / f"{tag}_k_*_raw.csv"))):
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            lines += [f"## {d.get('Kernel Name', '?')}", "", "| metric | value |", "|---|---|"]
            for k, name in KEYS:
                if k in d and d[k] not in ("", None):
                    lines.append(f"| {name} (`{k}`) | {d[k]} {u.get(k, '')} |")
            stalls = sorted(((float(d[k]), k) for k in hdr if "average_warp_latency_issue_stalled" in k and k.endswith(".ratio") and d[k] not in ("", "n/a")),
                            reverse=True)[:6]
            if stalls:
                lines += ["", "top stall reasons (warp latency per issued instruction): " +
                          ", ".join(f"{k.split('issue_stalled_')[1].split('.')[0]} {v:.2f}" for v, k in stalls)]
            lines.append("")
    out.write_text("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
