"""Summarise ncu output for profiles/: launch list shares + key raw metrics of full captures.

  python tools/ncu_summary.py <launches.csv> [<capture.ncu-rep> ...] > profiles/rNN_summary.md

Runs here (no GPU): reads the CSV launch list written by
  ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file <launches.csv> <cmd>
and `ncu -i <rep> --page raw --csv` of each full capture.
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path):
    tot = collections.OrderedDict()
    for r in csv.reader(l for l in open(path) if l.startswith('"')):
        if r[0] == "ID" or len(r) < 15:
            continue
        name = r[4].split("(")[0]
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + float(r[14]) * 1e-6)
    s = sum(t for _, t in tot.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f%% |" % (k, n, t, 100 * t / s))
    print("\nsum %.1f ms over %d launches\n" % (s, sum(n for n, _ in tot.values())))


def capture(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for v in rows[2:]:
        print("### %s  (grid %s, block %s)\n" % (v[ki].split("(")[0], v[h.index("Grid Size")], v[h.index("Block Size")]))
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print("| %s | %s | %s |" % (k, v[i], u[i]))
        print()


if __name__ == "__main__":
    print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold cache, serialised)\n")
    launches(sys.argv[1])
    for rep in sys.argv[2:]:
        print("## %s (`ncu --set full --clock-control none --import-source on`)\n" % rep.split("/")[-1])
        capture(rep)
