"""Measurement probe for the in-process multi-GPU path (peb_multi_*): the C4 workload
(1024 poses x 50k-point model against the voxel-down-sampled 1944x1200 scene, 30 point-to-point iterations) through
peb_multi_icp_align_batch from HOST buffers on 1, 2, 4, ... devices of this box, ONE process, no process group.

    python tools/bench_multi_inproc.py [--devices 1,2,4,8] [--hypotheses 1024] [--reps 5] [--same-device]

Wall clock around the blocking call (that IS the call the node makes: guesses up, records back); one JSON line per
device count.  --same-device puts all contexts on device 0 (a functional check on a one-GPU box, not a speed-up).
Not part of bench.py's contract (the driver launches one rank per GPU); written at the end of round 1 without GPU
time left, so its numbers are for round 2 to take.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from pose_estimation_b200 import pcl  # noqa: E402
from pose_estimation_b200.testing import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="1,2,4,8")
    ap.add_argument("--hypotheses", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--same-device", action="store_true")
    args = ap.parse_args()

    ctx0 = pcl.Context(0)

    def ds(points, leaf):
        vg = pcl.VoxelGrid(ctx0)
        vg.setInputCloud(points)
        vg.setLeafSize(leaf)
        return vg.filter()

    prob = synth.make_c4(scale=args.scale, n_guesses=args.hypotheses, downsample=ds)
    guesses = prob.guess[: args.hypotheses]
    base = None
    reference = None
    for n in [int(v) for v in args.devices.split(",")]:
        devices = [0] * n if args.same_device else list(range(n))
        try:
            many = pcl.MultiContext(devices)
        except pcl.PebError as e:
            print(json.dumps({"n_devices": n, "unavailable": str(e)}), flush=True)
            continue
        icp = pcl.IterativeClosestPoint(many)
        icp.setMaximumIterations(30)
        icp.getConvergeCriteria().setAbsoluteMSE(-1.0)
        icp.setMaxCorrespondenceDistance(0.02)
        t0 = time.perf_counter()
        icp.setInputTarget(prob.target)
        icp.setInputSource(prob.source)
        t_set = time.perf_counter() - t0
        times = []
        res = None
        for r in range(args.warmup + args.reps):
            t0 = time.perf_counter()
            res = icp.alignBatch(guesses)
            if r >= args.warmup:
                times.append(time.perf_counter() - t0)
        records = [bytes(x) for x in res]
        if reference is None:
            reference = records
        med = float(np.median(times))
        if base is None:
            base = med
        print(json.dumps({"n_devices": n, "devices": devices, "hypotheses": len(guesses), "n_source": int(len(prob.source)),
                          "n_target": int(len(prob.target)), "ms_per_batch_median": 1e3 * med, "ms_per_batch_min": 1e3 * min(times),
                          "hypotheses_per_s": len(guesses) / med, "speedup_vs_first": base / med,
                          "scene_and_model_setup_ms": 1e3 * t_set, "records_identical_to_first": records == reference,
                          "kernel_launches": many.launch_count, "timing": "wall clock around the blocking host-buffer call"}),
              flush=True)
        many.close()
    ctx0.close()


if __name__ == "__main__":
    main()
