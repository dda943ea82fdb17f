cd /root/repo
echo "== full gpu suite"
timeout 1500 python -m pytest tests/ -x -q -m gpu -s 2>&1 | grep -E "passed|failed|unmodified trainer|PSNR oracle|Error" | tail -8
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench N=1 (driver's flags)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 300 gpurun_out/bench_r2c.err; python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_r2c.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'train', d['train']['ms_per_step'], d['train']['roofline']['frac'], d['train']['kernels_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['clocks'])
print(d['parity']['per_ray_rel']['rgb'], d['parity']['per_ray_rel']['depth'])
PY
echo "== bench default flags (200 steps)"
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['steps'], round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['clocks'])"
