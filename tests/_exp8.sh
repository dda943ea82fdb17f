cd /root/repo
echo "== tests (parity + backward) with overlap"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_backward_gpu.py -q -m gpu -x 2>&1 | tail -3
for f in "" "--no-overlap" "" "--no-overlap"; do
  echo "-- bench $f"
  timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), 'train', round(d['train']['ms_per_step'],3), d['clocks'].get('reasons'))"
done
