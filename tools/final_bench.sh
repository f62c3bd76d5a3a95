# final round-2 measurement on one GPU: parity tests, launch list, ncu --set full of the front-end kernel, the bench lines of
# BASELINE configs 2-5, the reference arm  -> gpurun_out/<tag>_*
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
CMX_LANES=1 timeout 600 ncu -k regex:'seed_front_kernel|cluster_kernel|pair_candidates|verify_|pairing_|collect_overflow|prep_kernel|seed_cta|select_kernel|emit_|compact_|barcode_' --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_run.py --calls 2 > gpurun_out/${TAG}_launch_run.log 2>&1
CMX_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:seed_front_kernel" --launch-skip 1 --launch-count 1 -f -o /tmp/capf python tools/profile_run.py --calls 2 > gpurun_out/${TAG}_front_cap.log 2>&1
ncu -i /tmp/capf.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_front_cap.raw.csv.gz
ncu -i /tmp/capf.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | grep -v '^"",""' | gzip > gpurun_out/${TAG}_front_cap.lines.csv.gz
bash tools/bench_configs.sh $TAG
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
echo "reference arm rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_reference_arm.json
du -sh gpurun_out
