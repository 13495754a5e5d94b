cd /root/repo
timeout 900 python -m pytest tests/test_gpu_grad.py tests/test_gpu_loss.py -m gpu -q > gpurun_out/gputest_grad.log 2>&1; tail -40 gpurun_out/gputest_grad.log
