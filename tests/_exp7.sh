cd /root/repo
nvidia-smi -L
echo "== 2-GPU tests"
timeout 900 python -m pytest tests/test_render_nccl_gpu.py tests/test_trainer_gpu.py -q -m gpu 2>&1 | tail -5
echo "== bench N=2 weak"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 300 gpurun_out/bench_r2_n2.err
echo "== bench N=2 strong"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong > gpurun_out/bench_r2_n2_strong.json 2> gpurun_out/bench_r2_n2_strong.err; tail -c 300 gpurun_out/bench_r2_n2_strong.err
python - <<'PY'
import json
for f in ('bench_r2_n2','bench_r2_n2_strong'):
    try:
        d=json.loads(open('/root/repo/gpurun_out/%s.json'%f).read())
        print(f, d['n_gpus'], d['scaling'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['frac'], d['config']['rays_per_gpu'], d['train']['ms_per_step'], round(d['train']['value']), d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
