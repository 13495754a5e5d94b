cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; wc -l gpurun_out/bench_r2_final.json; cut -c1-200 gpurun_out/bench_r2_final.json
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; cut -c1-200 gpurun_out/bench_r2_reference.json
