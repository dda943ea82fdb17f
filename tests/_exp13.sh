cd /root/repo
for i in 1 2 3; do
  NERFPP_B200_LIB=/root/repo/tests/ref_lib/libnerfpp_b200_prev.so python tests/diag_field_ab.py 2>&1 | tail -1
  python tests/diag_field_ab.py 2>&1 | tail -1
done
