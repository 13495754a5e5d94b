cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log | cut -c1-200
timeout 600 python tools/fuzz_parity.py 60 777 > gpurun_out/fuzz4.log 2>&1; tail -4 gpurun_out/fuzz4.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; cut -c1-300 gpurun_out/bench_r2_final.json
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; cut -c1-200 gpurun_out/bench_r2_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bf16.csv python tools/profile_step.py --iters 1 > gpurun_out/ncu_ll.log 2>&1; tail -1 gpurun_out/ncu_ll.log
