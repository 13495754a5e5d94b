cd /root/repo
timeout 900 python -m pytest tests/test_gpu_grad.py -q -x > gpurun_out/gputest_bwd.log 2>&1; tail -4 gpurun_out/gputest_bwd.log | cut -c1-400
(timeout 600 python tools/train_bench.py 32 512 5; timeout 600 python tools/train_bench.py 8 2048 3) 2>&1 | grep -v Warn | grep '"impl"' | grep glue_factory | tee gpurun_out/train_tcbwd7.log | cut -c1-230
