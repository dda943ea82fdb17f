cd /root/repo
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
run() {  # name, extra args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 50 --warmup 5 "$@" > gpurun_out/bench_r2_n${N}_$name.json 2> gpurun_out/bench_r2_n${N}_$name.err || tail -c 800 gpurun_out/bench_r2_n${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('/root/repo/gpurun_out/bench_r2_n${N}_$name.json').read())
    print('$name', 'N', d['n_gpus'], d['scaling'], 'rays/gpu', d['config']['rays_per_gpu'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'train ms', round(d['train']['ms_per_step'],3), 'train rays/s', round(d['train']['value']), d['clocks'].get('reasons'))
except Exception as e: print('$name', 'ERR', e)
PY
}
run weak
run strong --scaling strong
if [ "$N" = "8" ]; then run c5 --config c5; fi
if [ "$N" = "4" ]; then run c4 --config c4; fi
