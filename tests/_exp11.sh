cd /root/repo
mkdir -p /tmp/prof
echo "== launch list: bench step + train step"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/prof/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_r2.log 2>&1; tail -2 gpurun_out/b_ncu_r2.log | cut -c1-200
python profiles/summarize_ncu.py launches /tmp/prof/launches_r2.csv > gpurun_out/r2_launches.txt
echo "== full: inference field kernels (3rd cascade pass = launches 8..11) + split"
ncu --set full --clock-control none --import-source on -k regex:"field_tc_kernel" -s 8 -c 5 -o /tmp/prof/prof_r2c_field python tests/diag_field_once.py > gpurun_out/r2c_field_ncu.log 2>&1; tail -2 gpurun_out/r2c_field_ncu.log
python profiles/summarize_ncu.py full /tmp/prof/prof_r2c_field.ncu-rep > gpurun_out/r2c_field_tc_ncu.txt
echo "== full: one training step"
ncu --set full --clock-control none --import-source on -k regex:"field_tc_kernel|bwd_fused_kernel|wgrad_tc_kernel|wgrad_small_kernel" -s 48 -c 16 -o /tmp/prof/prof_r2c_train python tests/diag_train_step.py > gpurun_out/r2c_train_ncu.log 2>&1; tail -2 gpurun_out/r2c_train_ncu.log
python profiles/summarize_ncu.py full /tmp/prof/prof_r2c_train.ncu-rep > gpurun_out/r2c_train_ncu.txt
python profiles/summarize_ncu.py traffic /tmp/prof/prof_r2c_field.ncu-rep /tmp/prof/prof_r2c_train.ncu-rep > gpurun_out/traffic.json
ls -la /tmp/prof gpurun_out | head -30
cat gpurun_out/traffic.json | head -30
