#!/bin/bash
# usage: tools/capture_icp.sh <tag>      (on the GPU box, from the repository root)
# Launch list of one bench step and `ncu --set full` captures of the batched icp_iteration_kernel at iterations 0, 1, 14
# and 28 of a warmed-up step.  The bench issues its launches iteration-major over S = 4 chains of 256 hypotheses, so the
# k-th icp_iteration_kernel launch of a step is iteration k / 4 of chain k % 4; one warm-up step = 120 launches.
# gpurun brings back at most 64 MiB: the raw and source pages are exported as CSV here and only the report of the
# iteration-14 launch (one launch, ~16 MB) travels.
tag=$1
B="python bench.py --steps 1 --warmup 1 --no-single --no-cpu-baseline --no-inproc"
$B > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $B \
  > gpurun_out/${tag}_launches.log 2>&1
for spec in "it0:120" "it1:124" "it14:176" "it28:232"; do
  IFS=: read name skip <<< "$spec"
  rep=gpurun_out/${tag}_icp_${name}
  ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel --launch-skip $skip \
    --launch-count 1 -f -o $rep $B > ${rep}.log 2>&1
  ncu -i ${rep}.ncu-rep --page raw --csv > ${rep}_raw.csv 2>> ${rep}.log
  ncu -i ${rep}.ncu-rep --page source --csv > ${rep}_source.csv 2>> ${rep}.log
  [ "$name" != it14 ] && rm -f ${rep}.ncu-rep
done
ls -la gpurun_out/${tag}_*
