# (only the pipeline's own kernels are profiled: with the data generator's thousands of torch kernels included, two metrics
# per launch took more than ten minutes per run and the first attempt of this script ran out of budget)
# A/B of the L2 fetch-granularity hint (cudaLimitMaxL2FetchGranularity) on serialised calls: per-kernel ncu durations and DRAM
# read bytes at the driver default (0) / 32 / 64 / 128 bytes -> gpurun_out/<tag>_l2fetch.txt (the CSVs stay on the box)
TAG=${1:-ab}
mkdir -p gpurun_out
for g in 0 32 64 128; do
  CMX_L2_FETCH=$g CMX_LANES=1 timeout 600 ncu -k regex:'seed_front_kernel|cluster_kernel|pair_candidates|verify_|pairing_|collect_overflow|prep_kernel|seed_cta|select_kernel|emit_|compact_|barcode_' --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file /tmp/l2f${g}.csv python tools/profile_run.py --calls 2 > /tmp/l2f${g}.log 2>&1
  echo "gran $g rc=$?"
done
python tools/ab_l2fetch_sum.py > gpurun_out/${TAG}_l2fetch.txt; cat gpurun_out/${TAG}_l2fetch.txt
