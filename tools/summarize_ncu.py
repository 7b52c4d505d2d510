"""Turns the ncu outputs of one profiled proof (tools/profile_step.py) into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/rNN_launches.csv profiles/rNN_launch_summary.csv
  python tools/summarize_ncu.py full gpurun_out/full.ncu-rep profiles/rNN_ncu_full_summary.csv
The launch list comes from `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv`;
the full summary from `ncu --set full --clock-control none --import-source on` read back with `ncu -i ... --page raw --csv`.
"""
import collections
import csv
import re
import subprocess
import sys


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def launches(src, dst_list, dst_summary):
    lines = [l for l in open(src) if l.startswith('"')]
    open(dst_list, "w").writelines(lines)
    rows = list(csv.DictReader(lines))
    tot = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        n, ms = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, ms + float(r["Metric Value"]) / 1e6)
    total = sum(ms for _, ms in tot.values())
    with open(dst_summary, "w") as f:
        f.write("# ncu launch list of ONE proof (poseidon-1000 shapes, m=21, m_0=20), `--metrics gpu__time_duration.sum --clock-control none`\n")
        f.write("# command: ncu --profile-from-start off ... python tools/profile_step.py   (cold-cache, serialised: compare SHARES, not absolutes)\n")
        f.write(f"# total {total:.3f} ms over {len(rows)} launches\n")
        f.write("kernel,launches,ms,share_pct\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{ms:.4f},{100 * ms / total:.2f}\n")
    print(f"{len(rows)} launches, {total:.3f} ms")


COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def full(rep, dst):
    # rep: the .ncu-rep, or its `ncu -i rep --page raw --csv` dump (made on the GPU box when the report is too big to bring back)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [c for c in COLS if c in hdr]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one proof (tools/profile_step.py), first launches of the listed kernels\n")
        f.write("id,kernel," + ",".join(f"{c} [{units[hdr.index(c)]}]" for c in cols) + "\n")
        for i, r in enumerate(rows[2:]):
            f.write(f"{i},{short(r[hdr.index('Kernel Name')])}," + ",".join(r[hdr.index(c)] for c in cols) + "\n")
    print(len(rows) - 2, "kernels")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    else:
        full(*sys.argv[2:4])
