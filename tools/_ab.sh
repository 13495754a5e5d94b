cd /root/repo
timeout 900 python -m pytest tests/test_gpu_grad.py tests/test_gpu_loss.py tests/test_gpu_x3.py -q -x > gpurun_out/gputest_bwd.log 2>&1; tail -3 gpurun_out/gputest_bwd.log | cut -c1-300
(timeout 600 python tools/train_bench.py 32 512 5; timeout 600 python tools/train_bench.py 8 2048 3) 2>&1 | grep -v Warn | grep '"impl"' | tee gpurun_out/train_tcbwd4.log | cut -c1-230
timeout 300 ncu --set full --clock-control none --import-source on -k regex:x3_attn_bwd --launch-count 2 -o gpurun_out/r2_attn_bwd_ts_full python tools/attn_bwd_bench.py 16 2048 1 > gpurun_out/ncu_bwd.log 2>&1; tail -1 gpurun_out/ncu_bwd.log
