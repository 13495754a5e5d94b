cd /root/repo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:x3_attn_bwd --launch-count 2 -o gpurun_out/r2_attn_bwd_full python tools/attn_bwd_bench.py 16 2048 1 > gpurun_out/ncu_bwd.log 2>&1; tail -3 gpurun_out/ncu_bwd.log
ls -la gpurun_out/r2_attn_bwd_full.ncu-rep
