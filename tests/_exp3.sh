cd /root/repo
echo "== backward tests"
timeout 900 python -m pytest tests/test_backward_gpu.py -x -q -m gpu 2>&1 | tail -8
for m in 2 1 2 1; do
  echo "-- bwd mode $m"
  BWD_MODE=$m timeout 120 python tests/diag_train_step.py 2>&1 | tail -1
done
echo "== convergence (two loss scales)"
timeout 600 python tests/diag_convergence.py --steps 600 --rays 1024 > gpurun_out/r2_convergence2.log 2>&1; grep -E "delta_psnr|psnr_mean|min_cos" gpurun_out/r2_convergence2.log | head; grep -E "^(oracle|ours) +step (100|200|300|400|500) " gpurun_out/r2_convergence2.log
