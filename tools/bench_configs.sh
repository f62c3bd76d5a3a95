# bench lines of BASELINE configs 2-5 on one GPU -> gpurun_out/<tag>_bench_*.json (committed under profiles/ afterwards)
TAG=${1:-r2}
mkdir -p gpurun_out
run() {  # name args...
  name=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  echo "== $name rc=$?"; tail -c 700 gpurun_out/${TAG}_bench_${name}.json; echo; tail -2 gpurun_out/${TAG}_bench_${name}.err
}
run chip --steps 20 --warmup 3
run atac --preset atac --steps 20 --warmup 3 --no-cpu-baseline
run scatac --preset atac --barcodes --steps 20 --warmup 3 --no-cpu-baseline
run hic --preset hic --read-len 150 --pairs-per-step 1000000 --steps 10 --warmup 3 --no-cpu-baseline
