# parity tests + ncu launch list + one scale-parity case -> gpurun_out/<tag>_*
TAG=${1:-q}
CASES=${2:-chip}
bash tools/gpu_quick.sh $TAG
timeout 900 python tools/scale_parity.py --cases $CASES > gpurun_out/${TAG}_scale_parity.jsonl 2> gpurun_out/${TAG}_scale_parity.err
echo "scale_parity rc=$?"; cat gpurun_out/${TAG}_scale_parity.jsonl | cut -c1-1200; tail -3 gpurun_out/${TAG}_scale_parity.err
