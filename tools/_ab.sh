cd /root/repo
timeout 400 python tools/adaptive_bench.py 2>&1 | grep -v Warn > gpurun_out/adaptive.log; cat gpurun_out/adaptive.log | cut -c1-330
