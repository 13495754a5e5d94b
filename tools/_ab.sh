cd /root/repo
(timeout 200 python tools/attn_bwd_accuracy.py 2 2048; timeout 200 python tools/attn_bwd_accuracy.py 1 8192) 2>&1 | grep -v Warn | tee gpurun_out/attn_bwd_accuracy2.log
timeout 900 python -m pytest tests/test_gpu_grad.py -q -x > gpurun_out/gputest_bwd.log 2>&1; tail -8 gpurun_out/gputest_bwd.log | cut -c1-300
(timeout 120 python tools/attn_bwd_bench.py 16 2048 5) 2>&1 | grep -v Warn | tail -1
