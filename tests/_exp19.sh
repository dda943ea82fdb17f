cd /root/repo
mkdir -p /tmp/prof
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file /tmp/prof/launches_train.csv python tests/diag_train_step.py > /dev/null 2>&1
python profiles/summarize_ncu.py launches /tmp/prof/launches_train.csv > gpurun_out/r2_train_launches.txt
head -40 gpurun_out/r2_train_launches.txt | cut -c1-130
