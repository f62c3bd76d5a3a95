TAG=${1:-r2}
CASES=${2:-chip,atac,hic,scatac}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 1500 python tools/scale_parity.py --cases $CASES > gpurun_out/${TAG}_scale_parity.jsonl 2> gpurun_out/${TAG}_scale_parity.err
echo "scale_parity rc=$?"; cat gpurun_out/${TAG}_scale_parity.jsonl | cut -c1-900; tail -5 gpurun_out/${TAG}_scale_parity.err
