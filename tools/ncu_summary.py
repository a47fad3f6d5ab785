"""Condense `ncu --page raw --csv` exports (tools/ncu_round2.sh) into the per-launch table committed under profiles/."""
import csv
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"), ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"), ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"), ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    print("kernel | " + " | ".join(n for c, n in COLS if c in idx))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")[:40]
        vals = []
        for c, n in COLS:
            if c not in idx:
                continue
            v = r[idx[c]].replace(",", "")
            try:
                f = float(v)
                u = units[idx[c]]
                if n == "time":
                    f = f / 1000 if u == "ns" else f
                    vals.append(f"{f:.1f}us" if u in ("ns", "us") else f"{f:.3f}{u}")
                elif n in ("dram_rd", "dram_wr"):
                    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                    vals.append(f"{f * mult / 1e6:.1f}MB")
                elif f >= 1000:
                    vals.append(f"{f:.0f}")
                else:
                    vals.append(f"{f:.2f}")
            except ValueError:
                vals.append(v[:12])
        print(f"{name} | " + " | ".join(vals))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
        print()
