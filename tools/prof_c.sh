# parity + ncu launch list + ncu --set full with per-CUDA-line metrics of the named kernels -> gpurun_out/<tag>_*
# usage: prof_c.sh <tag> <kernel regex> <launch-skip> <launch-count> [notest]
set -x
TAG=${1:-r2d}
KERN=${2:-pair_candidates_cta_kernel|seed_front_kernel}
SKIP=${3:-3}
COUNT=${4:-3}
mkdir -p gpurun_out
if [ "$5" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
CMX_LANES=1 timeout 600 ncu -k regex:'seed_front_kernel|cluster_kernel|pair_candidates|verify_|pairing_|collect_overflow|prep_kernel|seed_cta|select_kernel|emit_|compact_|barcode_' --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_run.py --calls 2 > gpurun_out/${TAG}_launch_run.log 2>&1
CMX_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KERN" --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/cap python tools/profile_run.py --calls 2 > gpurun_out/${TAG}_cap.log 2>&1
ncu -i /tmp/cap.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_cap.raw.csv.gz
ncu -i /tmp/cap.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | grep -v '^"",""' | gzip > gpurun_out/${TAG}_cap.lines.csv.gz
ls -la gpurun_out/${TAG}*; du -sh gpurun_out
