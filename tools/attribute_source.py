"""Joins `ncu --page source --csv` (per-SASS-instruction counters of one kernel) with `nvdisasm --print-line-info` of the
library's cubin: warp instructions, thread instructions and stall samples per source line.
usage: attribute_source.py <report.ncu-rep | source-page.csv> <kernel regex> <object.o> <mangled-name substring> [launch index] [--functions]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile


def line_table(obj: str, needle: str):
    tmp = tempfile.mkdtemp()
    import os
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = subprocess.run("ls *.cubin", cwd=tmp, shell=True, capture_output=True, text=True).stdout.split()[0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, capture_output=True, text=True).stdout
    table, cur, on = {}, None, False
    for ln in sass.splitlines():
        if ln.startswith("\t.section\t.text."):
            on = needle in ln
            cur = None
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, kregex, obj, needle = sys.argv[1:5]
    launch = sys.argv[5] if len(sys.argv) > 5 and not sys.argv[5].startswith("--") else "0"
    if rep.endswith(".csv"):  # an `ncu --page source --csv` export made on the GPU box (reports over 64 MiB do not travel)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}"], capture_output=True, text=True).stdout
    allrows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
    rows = allrows[starts[int(launch)]:starts[int(launch) + 1]]
    hdr = rows[1]
    ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    table = line_table(obj, needle)
    base = int(rows[2][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[2:]:
        off = int(r[ia], 16) - base
        key = table.get(off, (("?", 0), ""))[0] or ("?", 0)
        v = [int(r[ii]), int(r[it]), int(r[isamp])]
        for k in range(3):
            agg[key][k] += v[k]
            tot[k] += v[k]
    print(f"# {rows[0][1]}\n# total warp instructions {tot[0]}, thread instructions {tot[1]} ({tot[1] / max(tot[0], 1):.1f} per warp instruction), stall samples {tot[2]}")
    if "--functions" in sys.argv:  # the same per enclosing function of the library's sources
        import bisect
        import pathlib
        src = pathlib.Path(__file__).resolve().parent.parent / "pose_estimation_b200" / "csrc"
        heads = {}
        for f in src.glob("*.cu*"):
            hs = []
            for n, l in enumerate(f.read_text().splitlines(), 1):
                m = re.match(r"^(?:PEB_HD|__device__|__global__|static|inline)[^;]*?\b([A-Za-z_0-9]+)\s*\(", l)
                if m:
                    hs.append((n, m.group(1)))
            heads[f.name] = hs
        fagg = collections.defaultdict(lambda: [0, 0, 0])
        for (f, ln), v in agg.items():
            name = f
            if f in heads and heads[f]:
                k = bisect.bisect_right([x[0] for x in heads[f]], ln) - 1
                name = f"{f}:{heads[f][k][1] if k >= 0 else '?'}"
            for k in range(3):
                fagg[name][k] += v[k]
        print("# function  warp-inst %  lanes  samples %")
        for name, v in sorted(fagg.items(), key=lambda kv: -kv[1][0])[:30]:
            print(f"{name:48s} {100 * v[0] / tot[0]:6.2f} {v[1] / max(v[0], 1):5.1f} {100 * v[2] / max(tot[2], 1):6.2f}")
        return
    print("# file:line  warp-inst %  thread-inst %  lanes  samples %")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
        print(f"{key[0]}:{key[1]:<5d} {100 * v[0] / tot[0]:6.2f} {100 * v[1] / tot[1]:6.2f} {v[1] / max(v[0], 1):5.1f} {100 * v[2] / max(tot[2], 1):6.2f}")


if __name__ == "__main__":
    main()
