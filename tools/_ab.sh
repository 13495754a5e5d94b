cd /root/repo
timeout 900 python -m pytest tests/test_gpu_grad.py -q -x > gpurun_out/gputest_bwd.log 2>&1; tail -2 gpurun_out/gputest_bwd.log | cut -c1-300
python tools/attn_bwd_bench.py 16 2048 10 2>&1 | tail -1
python tools/attn_bwd_bench.py 64 512 10 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_bwd -c 9 --csv --log-file gpurun_out/bwd_ll.csv python tools/attn_bwd_bench.py 16 2048 1 > /dev/null 2>&1; tail -3 gpurun_out/bwd_ll.csv | cut -c60-100,380-400
