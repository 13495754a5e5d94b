cd /root/repo
timeout 300 python -m pytest tests/test_gpu_grad.py -q -x -k "attention_bwd" > gpurun_out/gputest_bwd.log 2>&1; tail -15 gpurun_out/gputest_bwd.log | cut -c1-300
(timeout 120 python tools/attn_bwd_bench.py 16 2048 5; LGB200_ATTN_BWD_MMASYNC=1 timeout 120 python tools/attn_bwd_bench.py 16 2048 5) 2>&1 | grep -v Warn | tail -4 | tee gpurun_out/attn_bwd_bench.log | cut -c1-300
