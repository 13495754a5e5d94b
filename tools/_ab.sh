cd /root/repo
timeout 900 python -m pytest tests/test_gpu_x3.py tests/test_gpu_parity.py tests/test_gpu_grad.py tests/test_gpu_baseline_shapes.py -q -x > gpurun_out/gputest_x3.log 2>&1; tail -3 gpurun_out/gputest_x3.log | cut -c1-300
for ss in 1 0 1 0; do echo "ss=$ss"; LGB200_X3_ATTN_SS=$ss timeout 200 python tools/sweep_fp32.py 2>&1 | grep -v Warn | tail -2; done | tee gpurun_out/x3_attn_ts_ab.log
