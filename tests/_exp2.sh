cd /root/repo
echo "== fused vs split tests"
timeout 600 python -m pytest tests/test_backward_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== train step timing: fused (default consumers), then sweep"
for c in 480 400 440 520 560; do
  echo "-- consumers permille $c"
  BWD_CONS=$c timeout 120 python tests/diag_train_step.py 2>&1 | tail -1
done
echo "-- split"
BWD_MODE=1 timeout 120 python tests/diag_train_step.py 2>&1 | tail -1
echo "== precision attribution"
timeout 400 python tests/diag_precision.py 2>&1 | tail -14
