"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv profiles/r01_launches_td.txt
    python profiles/summarize.py raw gpurun_out/prof_td.ncu-rep profiles/r01_ncu_td.txt
"""
import collections
import csv
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void <unnamed>::", "").replace("void ", "")


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        a = agg[short(r["Kernel Name"])]
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ({len(rows)} launches; cold-cache, "
                "serialised: compare SHARES)\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>10s} {'avg_us':>8s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:60s} {v[0]:8d} {v[1] / 1e3:10.1f} {v[1] / v[0] / 1e3:8.2f} {v[1] / tot * 100:6.1f}%\n")
    print(open(dst).read())


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, {src}\n")
        for r in rows[2:]:
            f.write(f"\n== {short(r[hdr.index('Kernel Name')])}  (id {r[0]})\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"  {w:80s} {r[i]:>16s} {units[i]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
