cd /root/repo
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; grep -E "^FAILED|passed|failed|AssertionError: pair" gpurun_out/gputest.log | head -30
python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline --no-gpu-library > gpurun_out/bench_fp32.json 2>/dev/null; cut -c1-220 gpurun_out/bench_fp32.json
