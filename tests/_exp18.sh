cd /root/repo
for i in 1 2 3; do for f in 1 0; do echo -n "folded=$f "; HEADS_FOLDED=$f timeout 120 python tests/diag_train_step.py 2>&1 | tail -1; done; done
