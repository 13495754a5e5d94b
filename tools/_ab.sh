cd /root/repo
timeout 600 python -m pytest tests/test_gpu_grad.py -q -x -k "attention_bwd" > gpurun_out/gputest_bwd.log 2>&1; tail -3 gpurun_out/gputest_bwd.log | cut -c1-300
for cfg in 44 22 44; do echo "np=$cfg"; LGB200_X3_BWD_NP=$cfg timeout 120 python tools/attn_bwd_bench.py 16 2048 10 2>&1 | grep -v Warn | tail -1; done | tee gpurun_out/attn_bwd_np_ab.log
timeout 200 python tools/attn_bwd_accuracy.py 1 8192 2>&1 | grep -v Warn
