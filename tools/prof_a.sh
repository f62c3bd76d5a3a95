# ncu --set full captures of the named kernels during the second cmx_map_batch_pe call; exports CSV pages on the box
# (the .ncu-rep files are too large to bring back) -> gpurun_out/<tag>_<name>.{raw,source}.csv.gz
set -x
TAG=${1:-r2a}
mkdir -p gpurun_out
cap() {  # name regex skip count
  CMX_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o /tmp/$1 python tools/profile_run.py --calls 2 > gpurun_out/${TAG}_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${TAG}_$1.raw.csv.gz
  ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_$1.source.csv.gz
}
cap tiers 'pair_candidates_cta_kernel|seed_cta_kernel|emit_kernel|verify_cta_kernel|pairing_cta_kernel' 11 11
cap t0 '^minimizer_kernel|^cluster_kernel|^pair_candidates_kernel|^verify_kernel|^pairing_kernel|^prep_kernel' 10 10
ls -la gpurun_out/${TAG}*
du -sh gpurun_out
