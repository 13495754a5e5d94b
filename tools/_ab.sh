cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
timeout 600 python tools/sweep.py > gpurun_out/sweep_r2.jsonl 2> gpurun_out/sweep.err; cat gpurun_out/sweep_r2.jsonl | cut -c1-200
