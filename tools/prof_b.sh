# parity + bench + ncu launch list + ncu --set full (CUDA-source view) of the named kernels -> gpurun_out/<tag>_*
set -x
TAG=${1:-r2c}
KERN=${2:-pair_candidates_cta_kernel|seed_front_kernel}
SKIP=${3:-3}
COUNT=${4:-3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1200 gpurun_out/${TAG}_bench.json
CMX_LANES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_run.py --calls 2 > gpurun_out/${TAG}_launch_run.log 2>&1
CMX_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$KERN" --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/cap python tools/profile_run.py --calls 2 > gpurun_out/${TAG}_cap.log 2>&1
ncu -i /tmp/cap.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_cap.raw.csv.gz
ncu -i /tmp/cap.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip > gpurun_out/${TAG}_cap.cuda.csv.gz
ncu -i /tmp/cap.ncu-rep --page details --csv 2>/dev/null | gzip > gpurun_out/${TAG}_cap.details.csv.gz
ls -la gpurun_out/${TAG}*; du -sh gpurun_out
