#!/bin/bash
# usage: tools/collect_profiles.sh <tag>      (here, after `gpurun ... tools/capture_icp.sh <tag>` and the bench runs of the same call)
# Turns what the GPU call left under gpurun_out/<tag>_* into the tracked summaries under profiles/: bench lines, test and smoke
# logs, the launch list per kernel, the `ncu --set full` reading of iterations 0 / 1 / 14 / 28, the per-function attribution
# and the measured DRAM traffic per launch (profiles/roofline_traffic.json, read by bench.py).
set -e
tag=$1
cd "$(dirname "$0")/.."
for f in bench_n1.json bench_reference.json pytest_gpu.log smoke.log; do
  [ -f gpurun_out/${tag}_$f ] && cp gpurun_out/${tag}_$f profiles/
done
raws=$(for it in it0 it1 it14 it28; do printf "gpurun_out/${tag}_icp_${it}_raw.csv,"; done | sed 's/,$//')
python tools/summarize_ncu.py raw "$raws" profiles/${tag}_icp_batch_kernels_full.txt \
  "ncu --set full of the batched icp_iteration_kernel: iterations 0, 1, 14 and 28 of one chain (256 hypotheses) of a warmed-up C4 step (tools/capture_icp.sh)" > /dev/null
python tools/summarize_ncu.py launches gpurun_out/${tag}_launches.csv profiles/${tag}_launches_bench_c4.txt \
  "every launch of bench.py --steps 1 --warmup 1 --no-single --no-cpu-baseline --no-inproc (warm-up step, timed step, profile-2 detail step)" > /dev/null
{
  echo "# per-function attribution of the batched icp_iteration_kernel (ncu --page source joined with the cubin's line table,"
  echo "# tools/attribute_source.py --functions); one chain (256 hypotheses) of a warmed-up C4 step"
  echo
  echo "## iteration 0 (candidates by greedy descent on the graph, cooperative verification)"
  python tools/attribute_source.py gpurun_out/${tag}_icp_it0_source.csv icp_iteration_kernel pose_estimation_b200/csrc/build/icp.o ILi1ELi0ELi8ELb0ELb1ELi0E --functions
  for it in it1 it14 it28; do
    echo
    echo "## iteration ${it#it} (search over the k-NN graph, grid walk where the row cannot certify)"
    python tools/attribute_source.py gpurun_out/${tag}_icp_${it}_source.csv icp_iteration_kernel pose_estimation_b200/csrc/build/icp.o ILi1ELi0ELi8ELb0ELb0ELi4E --functions
  done
} > profiles/${tag}_icp_source_attribution.txt
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
def dram(it):
    rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_icp_{it}_raw.csv")) if len(r) > 20]
    hdr, units, r = rows[0], rows[1], rows[2]
    def mb(key):
        i = hdr.index(key)
        v = float(r[i].replace(",", ""))
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[units[i]]
    return mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
it0 = dram("it0")
warm = [dram(x) for x in ("it1", "it14", "it28")]
w = sum(warm) / len(warm)
mean = (it0 + 29 * w) / 30 * 4 * 1e6
d = {"icp_iteration_kernel_batch_bytes_per_launch": int(round(mean, -6)),
     "how": f"ncu --set full (profiles/{tag}_icp_batch_kernels_full.txt, one chain = 256 hypotheses, x 4 for the rank's 1024): "
            f"dram__bytes_read.sum + dram__bytes_write.sum = {it0:.0f} MB x 4 for launch 0 (cooperative first iteration) and "
            f"{w:.0f} MB x 4 for a warm launch over the k-NN graph (iterations 1, 14, 28: {', '.join('%.0f' % x for x in warm)} MB); "
            "mean over the 30 launches of a step",
     "algorithmic_bytes_per_launch": 2048000000, "hypotheses_per_launch": 1024, "n_source": 50000}
open("profiles/roofline_traffic.json", "w").write(json.dumps(d, indent=1))
print(d["icp_iteration_kernel_batch_bytes_per_launch"])
PY
[ -f profiles/${tag}_bench_n1.json ] && python tools/fill_baseline_table.py profiles/${tag}_bench_n1.json --write > /dev/null
ls profiles/${tag}_*
