cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; cut -c1-300 gpurun_out/bench_r2_final.json
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; cut -c1-400 gpurun_out/bench_r2_reference.json
