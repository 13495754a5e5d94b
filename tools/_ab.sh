cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -k "x3 or golden" > gpurun_out/gputest.log 2>&1; grep -E "^FAILED|passed|failed|AssertionError" gpurun_out/gputest.log | head
for np in 4 2; do LGB200_X3_ATTN_NP=$np python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline --no-gpu-library 2>/dev/null | cut -c1-140; done
