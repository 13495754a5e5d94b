cd /root/repo
python tools/attn_bwd_bench.py 16 2048 10 2>&1 | tail -1
LGB200_ATTN_BWD_SIMT=1 python tools/attn_bwd_bench.py 16 2048 5 2>&1 | tail -1
ncu --set full --import-source on --clock-control none -k regex:attn_bwd_tc_kernel -c 2 -o gpurun_out/bwd_tc python tools/attn_bwd_bench.py 16 2048 1 > gpurun_out/ncu_bwd.log 2>&1; tail -2 gpurun_out/ncu_bwd.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_bwd -c 9 --csv --log-file gpurun_out/bwd_ll.csv python tools/attn_bwd_bench.py 16 2048 1 > /dev/null 2>&1; tail -3 gpurun_out/bwd_ll.csv | cut -c1-40,60-120,200-400
