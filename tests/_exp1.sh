cd /root/repo
for cfg in "0 1" "8 1" "1024 1" "2048 1" "1032 1" "0 2"; do
  set -- $cfg
  echo "== FLAGS=$1 CLUSTER=$2"
  NERFPP_TC_FLAGS=$1 NERFPP_TC_CLUSTER=$2 timeout 120 python tests/diag_train_step.py 2>&1 | tail -1
done
