cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gputest_all.log 2>&1; tail -3 gpurun_out/gputest_all.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "
import json; d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.json
