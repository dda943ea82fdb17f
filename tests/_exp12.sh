cd /root/repo
echo "== full gpu suite"
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== small-batch overlap A/B"
for r in 512 1024 2048; do for o in 0 1; do
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-train --rays-per-gpu $r --overlap $o 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rays', $r, 'overlap', $o, 'ms', round(d['ms_per_step'],4), 'rays/s', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
echo "== bench N=1 default"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 300 gpurun_out/bench_r2b.err; python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_r2b.json').read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['train']['ms_per_step'], d['train']['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['clocks'])
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
