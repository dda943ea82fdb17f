cd /root/repo
echo "== new gpu tests"
timeout 1500 python -m pytest tests/test_convergence_gpu.py tests/test_trainer_gpu.py tests/test_render_nccl_gpu.py -q -m gpu -s 2>&1 | grep -v "Warning\|warn" | tail -60
echo "== bench train leg only (short)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['train'])"
