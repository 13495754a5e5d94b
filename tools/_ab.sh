cd /root/repo
timeout 900 python -m pytest tests/test_gpu_grad.py -q -x > gpurun_out/gputest_bwd.log 2>&1; tail -15 gpurun_out/gputest_bwd.log | cut -c1-300
(timeout 600 python tools/train_bench.py 32 512 5; timeout 600 python tools/train_bench.py 8 2048 3) 2>&1 | grep -v Warn | tee gpurun_out/train_tc.log | cut -c1-260
(LGB200_ATTN_BWD_SIMT=1 timeout 600 python tools/train_bench.py 8 2048 3) 2>&1 | grep -v Warn | head -2 | tee gpurun_out/train_simt.log | cut -c1-260
