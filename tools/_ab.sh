cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -k "x3 or golden or c1 or c2 or c3" > gpurun_out/gputest_x3.log 2>&1; tail -3 gpurun_out/gputest_x3.log | cut -c1-300
for lib in "" glue_factory_colon_b200/lib/var/x3cl2.so glue_factory_colon_b200/lib/var/x3cl1.so ""; do
echo "lib=$lib"; LGB200_LIB=$lib python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline --no-gpu-library 2>/dev/null | cut -c1-160
done
