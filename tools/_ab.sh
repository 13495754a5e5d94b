cd /root/repo
timeout 900 python -m pytest tests/test_gpu_grad.py tests/test_gpu_loss.py tests/test_gpu_x3.py tests/test_abi.py -q -x > gpurun_out/gputest_bwd.log 2>&1; tail -3 gpurun_out/gputest_bwd.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-200
