#!/usr/bin/env bash
# The first GPU call of the next round, in one gpurun (about 4 minutes of box time on one B200):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/next_gpu_call.sh'
# 1. the GPU tests whose code was written after round 1's GPU budget was spent (bit-identity of the warm_upfront kernels;
#    the multi-device layer ran already),
# 2. the bench with and without the experimental warm search (same box, back to back; no CPU legs),
# 3. the in-process multi-device probe with every context on device 0 (functional; real scaling needs --gpus N boxes).
# Everything lands in gpurun_out/.  Numbers printed under ncu are never bench values: no profiler here.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_zz_experimental_gpu.py tests/test_multi_device.py -q -m gpu -rxXs > gpurun_out/next_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/next_pytest.log
for rep in 1 2; do
  python bench.py --steps 5 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/next_bench_default_$rep.json 2> gpurun_out/next_bench_default_$rep.err
  PEB_OPTS=warm_upfront=1 python bench.py --steps 5 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/next_bench_upfront2x2_$rep.json 2> gpurun_out/next_bench_upfront2x2_$rep.err
  PEB_OPTS=warm_upfront=3 python bench.py --steps 5 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/next_bench_upfront3x3_$rep.json 2> gpurun_out/next_bench_upfront3x3_$rep.err
  PEB_OPTS=warm_upfront=3,warm_upfront_from=1 python bench.py --steps 5 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/next_bench_upfront3x3_from1_$rep.json 2> gpurun_out/next_bench_upfront3x3_from1_$rep.err
  PEB_OPTS=warm_upfront=3,warm_upfront_from=6 python bench.py --steps 5 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/next_bench_upfront3x3_from6_$rep.json 2> gpurun_out/next_bench_upfront3x3_from6_$rep.err
done
python tools/bench_multi_inproc.py --devices 1,2 --same-device --reps 3 > gpurun_out/next_multi_inproc.jsonl 2> gpurun_out/next_multi_inproc.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/next_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), "roofline", round(d["roofline"]["frac"], 4))
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
PY
