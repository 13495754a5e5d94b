cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gputest_all.log 2>&1; tail -4 gpurun_out/gputest_all.log | cut -c1-400
