export TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_chip.json 2> gpurun_out/${TAG}_bench_chip.err; echo "chip rc=$?"
timeout 900 python bench.py --preset hic --read-len 150 --pairs-per-step 1000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_hic.json 2> gpurun_out/${TAG}_bench_hic.err; echo "hic rc=$?"
python - <<'P'
import json
for n in ['chip','hic']:
    d=json.loads(open('gpurun_out/%s_bench_%s.json'%(__import__('os').environ.get('TAG','r2'),n)).read().strip().splitlines()[-1])
    print(n, 'value %.1f M  %.2f ms | e2e %.1f M %.2f ms'%(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']), d['kernel_ms_per_step'])
P
