cd /root/repo
echo "== new gpu tests"
timeout 1500 python -m pytest tests/test_convergence_gpu.py tests/test_trainer_gpu.py tests/test_render_nccl_gpu.py -x -q -m gpu -s 2>&1 | grep -v "Warning\|warn" | tail -40
echo "== full backward test verbose"
timeout 300 python -m pytest tests/test_backward_gpu.py -q -m gpu -s -k "full_backward" 2>&1 | grep -E "worst|passed|failed"
echo "== bench (short)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 600 gpurun_out/bench_r2a.err; python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_r2a.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d.get('train'), d.get('cpu_baseline'), d.get('parity',{}).get('per_ray_rel'))
PY
