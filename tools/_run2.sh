python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --config 4 --no-cpu-baseline > gpurun_out/r2w_bench_config4_8gpu.json 2> gpurun_out/r2w_bench_config4_8gpu.err; tail -c 300 gpurun_out/r2w_bench_config4_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --config 4 --no-cpu-baseline > gpurun_out/r2w_bench_config4_4gpu.json 2> gpurun_out/r2w_bench_config4_4gpu.err
python - <<'PY'
import json
for f in ('gpurun_out/r2w_bench_config4_8gpu.json','gpurun_out/r2w_bench_config4_4gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','n_gpus','scaling','ms_per_step','grow_ms_per_pass')}, d.get('e2e',{}).get('value'), d.get('per_rank'))
    except Exception as e: print(f, 'ERR', e)
PY
