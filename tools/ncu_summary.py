#!/usr/bin/env python
"""Condense `ncu --page raw --csv` exports into the table kept under profiles/ (one row per launch)."""
import csv, sys, json

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pct"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit"), ("lts__t_sector_hit_rate.pct", "l2_hit"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pct"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier")]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def load(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    out = []
    for r in data:
        if len(r) != len(hdr):
            continue
        rec = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        row = {"kernel": rec["Kernel Name"].split("(")[0].replace("void ", "")}
        for key, short in KEYS:
            if key in rec and rec[key] != "":
                try:
                    v = float(rec[key].replace(",", ""))
                except ValueError:
                    continue
                row[short] = v * UNIT_SCALE.get(u[key], 1.0) if short in ("us", "dram_rd", "dram_wr", "dsmem") else v
        if "us" in row:
            row["dram_gbs"] = (row.get("dram_rd", 0) + row.get("dram_wr", 0)) / row["us"] * 1e-3
        out.append(row)
    return out


if __name__ == "__main__":
    for path in sys.argv[1:]:
        print("#", path)
        for row in load(path):
            print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()}))
