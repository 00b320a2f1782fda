"""Turns ncu exports (raw page CSV + launch-list CSV) into the short text summaries under profiles/."""
import collections
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("lts__t_sectors.sum", "L2 sectors"),
]


def raw_summary(rep: str, title: str) -> str:
    """rep: one or more (comma-separated) .ncu-rep files or `ncu --page raw --csv` exports of them"""
    hdr, units, body = None, None, []
    for one in rep.split(","):
        if one.endswith(".csv"):
            out = open(one).read()
        else:
            out = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = [r for r in csv.reader(out.splitlines()) if len(r) > 20]
        hdr, units = rows[0], rows[1]
        body += [(one.split("/")[-1], r, rows[1]) for r in rows[2:]]
    ki = hdr.index("Kernel Name")
    lines = [f"# {title}", f"# source: ncu --set full --clock-control none --import-source on, {', '.join(x.split('/')[-1] for x in rep.split(','))}", ""]
    for n, (src, r, units) in enumerate(body):
        name = r[ki].replace("peb::<unnamed>::", "").replace("void ", "")
        lines.append(f"## launch {n} ({src}): {name[:110]}")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                lines.append(f"  {label:32s} {r[i]:>16s} {units[i]}")
        lines.append("")
    return "\n".join(lines)


def launch_summary(path: str, title: str) -> str:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].replace("peb::<unnamed>::", "").replace("void ", "")[:95]
        a = agg.setdefault(k, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {title}", f"# source: ncu --metrics gpu__time_duration.sum --clock-control none, {path.split('/')[-1]} "
             "(cold-cache, serialised: compare shares, not absolutes)", "",
             f"{'launches':>8s} {'total ms':>10s} {'share':>7s}  kernel (grid, block of the first launch)"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{a[0]:8d} {a[1] / 1e6:10.3f} {100 * a[1] / tot:6.1f}%  {k}  {a[2]} {a[3]}")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    kind, src, dst, title = sys.argv[1:5]
    text = raw_summary(src, title) if kind == "raw" else launch_summary(src, title)
    open(dst, "w").write(text)
    print(text[:3000])
