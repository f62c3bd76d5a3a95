# parity tests + ncu launch list (per-kernel durations of one serialised call) -> gpurun_out/<tag>_*
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
CMX_LANES=1 timeout 600 ncu -k regex:'seed_front_kernel|cluster_kernel|pair_candidates|verify_|pairing_|collect_overflow|prep_kernel|seed_cta|select_kernel|emit_|compact_|barcode_' --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_run.py --calls 2 > gpurun_out/${TAG}_launch_run.log 2>&1
tail -2 gpurun_out/${TAG}_launch_run.log | cut -c1-600
