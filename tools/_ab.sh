cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x -k "linear or golden or full_size or ragged or c2_bf16 or c3" > gpurun_out/gputest_ffn1.log 2>&1; tail -3 gpurun_out/gputest_ffn1.log | cut -c1-200
for i in 1 2; do
  echo "32-col staging, 3 stages: $(LGB200_LIB=glue_factory_colon_b200/lib/var/ffn1_full.so timeout 300 python tools/profile_step.py --iters 30 2>&1 | tail -1)"
  echo "16-col staging, 4 stages: $(timeout 300 python tools/profile_step.py --iters 30 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/ffn1_ab2.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tc_pair_ln -c 6 --csv --log-file gpurun_out/ffn1_ncu.csv python tools/profile_step.py --pairs 64 > /dev/null 2>&1; tail -3 gpurun_out/ffn1_ncu.csv | cut -c1-60,200-400
