#!/bin/bash
# One GPU-box pass that produces everything profiles/ cites for a round:
#   bench (N=1), reference arm, ncu launch list of one serialised call, ncu --set full of the probe kernel.
# Usage (from the repo root, on the GPU box): tools/round_profiles.sh <tag>      -> gpurun_out/<tag>_*
set -x
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/${TAG}_bench_N1.json 2> $OUT/${TAG}_bench_N1.err
python bench.py --impl reference --steps 6 --warmup 3 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err
CMX_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
  python tools/profile_run.py --calls 2 > $OUT/${TAG}_launch_run.log 2>&1
CMX_LANES=1 ncu --set full --clock-control none --import-source on -k regex:probe_kernel --launch-skip 1 --launch-count 1 -f \
  -o $OUT/${TAG}_probe python tools/profile_run.py --calls 2 > $OUT/${TAG}_probe_run.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/${TAG}_nvidia_smi.csv
tail -c 600 $OUT/${TAG}_bench_N1.json; echo; tail -c 600 $OUT/${TAG}_bench_reference_arm.json
