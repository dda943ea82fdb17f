cd /root/repo
timeout 900 python -m pytest tests/test_backward_gpu.py -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do timeout 120 python tests/diag_train_step.py 2>&1 | tail -1; done
