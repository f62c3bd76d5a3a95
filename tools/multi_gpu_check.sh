# N-GPU checks (run under gpurun --gpus N): end-to-end BED identity through the native exchange, then the bench line at N
N=${1:-2}
TAG=${2:-r2}
PRESETS=${3:-chip atac}
STEPS=${4:-20}
mkdir -p gpurun_out
for preset in $PRESETS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_map.py --dir tests/golden/synth_small --preset $preset --batch 500 > gpurun_out/${TAG}_multi_gpu_map_N${N}_${preset}.log 2>&1
  echo "multi_gpu_map $preset rc=$?"; grep "multi_gpu_map" gpurun_out/${TAG}_multi_gpu_map_N${N}_${preset}.log | cut -c1-400
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/${TAG}_bench_N${N}.json 2> gpurun_out/${TAG}_bench_N${N}.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_N${N}.json; tail -3 gpurun_out/${TAG}_bench_N${N}.err
