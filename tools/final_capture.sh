# final round-2 capture: parity tests, launch list, ncu --set full of the main kernels of the second call, scale parity cases
TAG=${1:-r2f}
CASES=${2:-chip,hic}
bash tools/prof_c.sh $TAG 'seed_front_kernel|pair_candidates|verify_kernel|verify_cta|cluster_kernel|emit_kernel' 12 12
timeout 900 python tools/scale_parity.py --cases $CASES > gpurun_out/${TAG}_scale_parity.jsonl 2> gpurun_out/${TAG}_scale_parity.err
echo "scale_parity rc=$?"; cat gpurun_out/${TAG}_scale_parity.jsonl | cut -c1-1800; tail -3 gpurun_out/${TAG}_scale_parity.err
