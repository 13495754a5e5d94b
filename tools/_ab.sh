cd /root/repo
timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q > gpurun_out/gputest_grad.log 2>&1; tail -2 gpurun_out/gputest_grad.log
timeout 600 python tools/train_bench.py 32 512 5 2>&1 | grep -v Warning | tail -4
timeout 600 python tools/train_bench.py 8 2048 3 2>&1 | grep -v Warning | tail -4
