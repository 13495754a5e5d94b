cd /root/repo
timeout 300 python -m pytest tests -m gpu -q -x -k "linear" > gpurun_out/gputest_2sm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/gputest_2sm.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q -x -k "golden or full_size or ragged or c2_bf16 or c3" >> gpurun_out/gputest_2sm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/gputest_2sm.log | cut -c1-300
for i in 1 2; do
  echo "relay: $(LGB200_LIB=glue_factory_colon_b200/lib/var/pair_relay.so timeout 300 python tools/profile_step.py --iters 30 2>&1 | tail -1)"
  echo "2sm-tma: $(timeout 300 python tools/profile_step.py --iters 30 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/pair2sm_ab.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_pair" -c 12 --csv --log-file gpurun_out/pair2sm_ncu.csv python tools/profile_step.py --pairs 64 > /dev/null 2>&1; tail -8 gpurun_out/pair2sm_ncu.csv | cut -c60-110,330-400
