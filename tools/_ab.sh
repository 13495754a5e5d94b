cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x -k "linear or golden or full_size or ragged" > gpurun_out/gputest_rev.log 2>&1; tail -3 gpurun_out/gputest_rev.log | cut -c1-200
for r in 0 3 0 3 1 2 4 5 7; do
  echo "REVERSE=$r $(LGB200_GEMM_REVERSE=$r timeout 300 python tools/profile_step.py --iters 30 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/rev_ab.log
for p in 0 2 3; do
  echo "PREFETCH=$p $(LGB200_GEMM_PREFETCH=$p timeout 300 python tools/profile_step.py --iters 30 2>&1 | tail -1)"
done 2>&1 | tee -a gpurun_out/rev_ab.log
