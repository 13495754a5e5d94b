cd /root/repo
V=glue_factory_colon_b200/lib/var
(LGB200_LIB=$V/attn_p8s1.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -k attention 2>&1 | tail -3) > gpurun_out/ab1.log 2>&1
for r in 1 2; do
for v in old4 p4s0 p4s1 p6s0 p6s1 p8s0 p8s1 p10s1; do
  echo -n "$v: " >> gpurun_out/ab1.log; LGB200_LIB=$V/attn_$v.so timeout 120 python tools/attn_bench.py >> gpurun_out/ab1.log 2>&1
done; done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; tail -15 gpurun_out/gputest.log >> gpurun_out/ab1.log
cat gpurun_out/ab1.log
