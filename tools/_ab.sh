cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/gputest_par.log 2>&1; tail -6 gpurun_out/gputest_par.log | cut -c1-400
