cd /root/repo
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_ll.csv python tools/train_profile.py 8 2048 > gpurun_out/train_ll.log 2>&1; tail -2 gpurun_out/train_ll.log
python tools/summarize_launches.py gpurun_out/train_ll.csv | head -40 | cut -c1-170
